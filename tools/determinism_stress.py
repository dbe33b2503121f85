"""Bit-wise repeatability of the forward (+ decode) on one input: run it N times (graph replays, optionally with other
lanes running concurrently) and report every run whose maps or decode counts differ from the first.
usage: determinism_stress.py [batch] [runs] [lanes]      (kernel selection through the LWOP_* environment switches)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lightweight_openpose_b200 import checkpoint, engine
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
RUNS = int(sys.argv[2]) if len(sys.argv) > 2 else 40
LANES = int(sys.argv[3]) if len(sys.argv) > 3 else 1
v, src = checkpoint.load_or_synthesize()
frames = bench.make_frames(B)
x = torch.from_numpy(frames).cuda() if os.environ.get("STRESS_U8") else torch.from_numpy((frames / 255.).astype(np.float32)).cuda()
pool = engine.EnginePool(368, 368, max_batch=B, lanes=LANES)
pool.load_variables(v, src)
ref = None
bad = 0
for run in range(RUNS):
    pool.fork()
    for _ in range(LANES):
        pool.step_device(x)
    pool.join()
    torch.cuda.synchronize()
    for k, e in enumerate(pool.engines):
        heat, paf = e._outputs(B, x.device)[:2]
        c = torch.from_numpy(e.fetch(B).counts[:, :2].copy())
        if ref is None:
            ref = (heat.clone(), paf.clone(), c.clone())
            continue
        dh = (heat != ref[0]).reshape(B, -1).any(1)
        dp = (paf != ref[1]).reshape(B, -1).any(1)
        dc = (c != ref[2]).any(1)
        if bool(dh.any()) or bool(dp.any()) or bool(dc.any()):
            bad += 1
            imgs = torch.nonzero(dh.cpu() | dp.cpu() | dc).flatten().tolist()
            md = float((heat - ref[0]).abs().max())
            # where inside the image do the maps differ (rows / cols of the 46 x 46 map)?
            b0 = imgs[0]
            w = torch.nonzero((paf[b0] != ref[1][b0]).any(-1) | (heat[b0] != ref[0][b0]).any(-1))
            box = (w.min(0).values.tolist(), w.max(0).values.tolist()) if len(w) else None
            print(f"run {run} lane {k}: {len(imgs)} images differ {imgs[:8]} max heat diff {md:.3g} counts differ {int(dc.sum())} first image box {box} npix {len(w)}")
print(f"batch {B} runs {RUNS} lanes {LANES}: {bad} differing (run, lane) results", {k: v for k, v in os.environ.items() if k.startswith('LWOP_') or k == 'STRESS_U8'})
