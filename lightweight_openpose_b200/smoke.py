"""One tiny forward + decode on cuda:0, checked against the CPU oracle (the
oracle is the checker here, never the thing that runs the path)."""
from __future__ import annotations

import os
import sys

import numpy as np


def run(verbose: bool = True) -> None:
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from lightweight_openpose_b200 import checkpoint, engine, synth
    from oracle.forward_oracle import ForwardOracle
    from oracle import decode_oracle

    engine.require_gpu(0)
    H = W = 128
    variables, source = checkpoint.load_or_synthesize()
    eng = engine.Engine(H, W, max_batch=2, device=0)
    eng.load_variables(variables, source)
    x = synth.synth_batch([0, 1], H, W)
    xd = torch.from_numpy(x).cuda()
    heat, paf = (t.clone() for t in eng.forward(xd, mode="bf16"))
    heat32, paf32 = eng.forward(xd, mode="fp32")
    torch.cuda.synchronize()
    oh, op = ForwardOracle(variables)(x)
    e16 = max(float(np.abs(heat.cpu().numpy() - oh).max()), float(np.abs(paf.cpu().numpy() - op).max()))
    e32 = max(float(np.abs(heat32.cpu().numpy() - oh).max()), float(np.abs(paf32.cpu().numpy() - op).max()))
    if verbose:
        print(f"smoke: weights={source} bf16 max-abs {e16:.3e} (tol 2e-2), fp32-check max-abs {e32:.3e} (tol 1e-4)")
    assert e16 <= 2e-2, e16
    assert e32 <= 1e-4, e32
    # decode on label-style scenes (integer outputs exact, floats <= 1e-5)
    sh, sp, _ = synth.synth_scene(3, 3, 368, 368)
    res = eng_decode(sh, sp)
    ref = decode_oracle.decode_detail((368, 368), {"thre1": 0.1, "thre2": 0.0}, sh, sp)
    jl, rjl = res.joint_list, ref["joint_list"]
    assert jl.shape == rjl.shape, (jl.shape, rjl.shape)
    assert np.array_equal(jl[:, [0, 1, 3, 4, 5]], rjl[:, [0, 1, 3, 4, 5]])
    assert np.abs(jl[:, 2] - rjl[:, 2]).max() <= 1e-5
    assert np.array_equal(res.person_to_joint_assoc[:, :14], ref["persons"][:, :14])
    assert np.array_equal(res.joints, ref["joints"])
    if verbose:
        print(f"smoke: decode {jl.shape[0]} joints, {res.person_to_joint_assoc.shape[0]} persons match the oracle; "
              f"kernel launches so far: {eng.launch_count()}")


def eng_decode(heat, paf):
    from lightweight_openpose_b200.src import pose_decode
    return pose_decode.decode_pose_batch((368, 368), {"thre1": 0.1, "thre2": 0.0}, heat[None], paf[None])[0]


if __name__ == "__main__":
    run()
