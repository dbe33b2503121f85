"""Whole-network error of the bf16 path on NON-benign weights: iid He-normal kernels (no spatial stencil, no identity
path -- every 3x3 filter is a random high-pass), BN statistics as in SURVEY 8(d) C1, one scalar gain per layer chosen
live by the CPU oracle so that activations stay O(1) (LSUV-style; heads scaled to the heat / PAF ranges).  Such a
network amplifies perturbations layer by layer, so this is the stress case for bf16 storage.  Reports, against the
exact fp32 oracle: the GPU fp32 check mode, the GPU bf16 path, and the CPU oracle with emulated bf16 storage
(`quantize=True`: what the FORMAT costs, whatever the kernels do).

usage: rough_weights.py [H W batch]     (GPU box; `make_rough_variables` is also used by tests/test_gpu_forward.py)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from lightweight_openpose_b200 import checkpoint as ck
from lightweight_openpose_b200.graph import LAYERS
from lightweight_openpose_b200.synth import synth_batch

HEAD_STD = {"heatmaps1": 0.10, "heatmaps2": 0.10, "pafs1": 0.15, "pafs2": 0.15}


def make_rough_variables(seed: int = 7, calib_hw=(96, 96)):
    """-> variables dict (float32).  One oracle pass per layer on two small frames (O(L^2) layer evaluations)."""
    from oracle.forward_oracle import ForwardOracle
    entries, _ = ck.read_index(ck.default_index_prefix() + ".index")
    rng = np.random.default_rng(seed)
    raw = {}
    for name in sorted(entries):
        if not ck.is_inference_variable(name):
            continue
        shape = entries[name].shape
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("kernel", "pointwise_kernel", "depthwise_kernel"):
            kh, kw, cin, _ = shape
            fan_in = kh * kw * (1 if leaf == "depthwise_kernel" else cin)
            v = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=shape)
        elif leaf == "bias":
            v = rng.normal(0.0, 0.01, size=shape)
        elif leaf == "gamma":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf in ("beta", "moving_mean"):
            v = rng.normal(0.0, 0.1, size=shape)
        else:
            v = rng.uniform(0.5, 1.5, size=shape)
        raw[name] = v.astype(np.float32)
    x = synth_batch([2000, 2001], calib_hw[0], calib_hw[1])
    gains, shifts = {}, {}
    for layer in LAYERS:
        orc = ForwardOracle(ck.apply_calibration(raw, gains, shifts))
        orc.trace = {}
        orc(x)
        cm, cs = orc.trace[layer.name + ":pre"]
        if layer.note in HEAD_STD:
            g = HEAD_STD[layer.note] / float(np.sqrt(np.mean(cs ** 2)))
        else:
            g = 1.0 / max(orc.trace[layer.name][0], 1e-12)          # output RMS -> 1
        gains[layer.name] = np.full_like(cs, g).astype(np.float32)
        shifts[layer.name] = np.zeros_like(cs).astype(np.float32)
    return ck.apply_calibration(raw, gains, shifts)


def errors(a, ref):
    d = np.asarray(a, dtype=np.float64) - np.asarray(ref, dtype=np.float64)
    return float(np.abs(d).max()), float(np.sqrt(np.mean(d * d)))


def main():
    import torch
    from lightweight_openpose_b200 import engine
    from oracle.forward_oracle import ForwardOracle
    H, W, B = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (184, 184, 2)))
    v = make_rough_variables()
    x = synth_batch(range(10, 10 + B), H, W)
    oh, op = ForwardOracle(v)(x)
    o64h, o64p = ForwardOracle(v, dtype=torch.float64)(x)
    qh, qp = ForwardOracle(v, quantize=True)(x)
    eng = engine.Engine(H, W, max_batch=B)
    eng.load_variables(v, "rough")
    xd = torch.from_numpy(x).cuda()
    gh, gp = (t.cpu().numpy() for t in eng.forward(xd, mode="bf16"))
    fh, fp = (t.cpu().numpy() for t in eng.forward(xd, mode="fp32"))
    rec = {"H": H, "W": W, "batch": B, "heat_std": float(oh.std()), "paf_std": float(op.std())}
    for name, (h, p) in {"oracle_fp32_vs_fp64": (oh, op), "gpu_fp32_check": (fh, fp), "gpu_bf16": (gh, gp),
                         "cpu_emulated_bf16_storage": (qh, qp)}.items():
        ref = (o64h, o64p)
        mh, rh = errors(h, ref[0])
        mp, rp = errors(p, ref[1])
        rec[name] = {"heat_maxabs": mh, "heat_rms": rh, "paf_maxabs": mp, "paf_rms": rp}
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
