"""Dev-time tool: compute the per-layer gains baked into
``lightweight_openpose_b200/checkpoint.py:SYNTH_GAINS``.

The raw seeded recipe (He-normal kernels etc.) lets activations grow by orders
of magnitude through 57 convs, which would make the absolute parity tolerances
(1e-4 fp32 / 2e-2 bf16 on heatmaps and PAFs) meaningless.  This tool runs the
CPU oracle layer by layer on a few seeded synthetic frames (synth.synth_image) and picks, for each
layer and output channel, the gain and bias shift that make its pre-activation
zero-mean / unit-std (hidden layers, like a trained BN network -- without the
centering a random ReLU network collapses onto its DC path and the spatial
signal drowns in bf16 rounding) or heat ~ N(-0.12, 0.1^2), paf ~ N(0, 0.15^2)
for the heads, so that a synthetic frame yields a crowded-scene-like number of
peaks above thre1 = 0.1.

Run:  python tools/calibrate_synthetic.py   (writes lightweight_openpose_b200/data/synth_calibration_61236.npz)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lightweight_openpose_b200 import checkpoint as ck          # noqa: E402
from lightweight_openpose_b200.graph import LAYERS              # noqa: E402
from lightweight_openpose_b200.synth import synth_batch         # noqa: E402
from oracle.forward_oracle import ForwardOracle                 # noqa: E402

HEADS = {"heatmaps1": (0.10, -0.12, 2), "heatmaps2": (0.10, -0.12, 0),
         "pafs1": (0.15, 0.0, 3), "pafs2": (0.15, 0.0, 1)}


def main():
    entries, _ = ck.read_index(ck.default_index_prefix() + ".index")
    raw = ck.synthetic_variables(entries, seed=61236, calibrated=False)
    x = synth_batch(range(1000, 1016), 368, 368)
    gains = {}
    shifts = {}
    for layer in LAYERS:
        v = ck.apply_calibration(raw, gains, shifts)
        orc = ForwardOracle(v)
        orc.trace = {}
        orc(x)
        cm, cs = orc.trace[layer.name + ":pre"]
        if layer.note in HEADS:
            # heads: per-channel gain and shift to the target statistics
            # one scalar gain (no per-channel amplification), per-channel shift
            tgt_std, tgt_mean, _ = HEADS[layer.note]
            g = np.full_like(cs, tgt_std / float(np.sqrt(np.mean(cs ** 2))))
            s_ = tgt_mean - cm * g
        else:
            # hidden layers: one scalar gain (output RMS -> 1), no centering: a
            # centred random ReLU stack is chaotic (rounding noise grows ~1.2x per
            # layer relative to the centred signal), an uncentred critical one is not.
            rms = orc.trace[layer.name][0]
            g = np.full_like(cs, 1.0 / rms)
            s_ = np.zeros_like(cs)
        gains[layer.name] = g.astype(np.float32)
        shifts[layer.name] = s_.astype(np.float32)
        print(f"# {layer.index:2d} {layer.name:36s} pre mean={cm.mean():9.4f} std={cs.mean():9.4f}", file=sys.stderr)
    v = ck.apply_calibration(raw, gains, shifts)
    h, p = ForwardOracle(v)(x)
    print("# final heat mean/std", h.mean(), h.std(), "frac>0.1", (h > 0.1).mean(),
          "paf", p.mean(), p.std(), file=sys.stderr)
    out = {}
    for k in gains:
        out["gain:" + k] = gains[k]
        out["shift:" + k] = shifts[k]
    path = os.path.join(os.path.dirname(os.path.abspath(ck.__file__)), "data", "synth_calibration_61236.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
