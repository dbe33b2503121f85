"""Steady-state throughput of the pooled host-buffer path (EnginePool.map) against the device-resident path at 32 frames
per step: usage e2e_variants.py <lanes> <depth> [host buffers]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lightweight_openpose_b200 import checkpoint, engine, synth
B, L, D, S = 32, int(sys.argv[1]), int(sys.argv[2]), 200
v, src = checkpoint.load_or_synthesize()
pool = engine.EnginePool(368, 368, max_batch=B, lanes=L, depth=D)
pool.load_variables(v, src)
base = np.stack([synth.synth_image(s, 368, 368) for s in range(32)])
nh = int(sys.argv[3]) if len(sys.argv) > 3 else L * D + 1
hs = [torch.from_numpy(np.ascontiguousarray(np.roll(base, k, 0))).pin_memory() for k in range(nh)]
xs = [torch.from_numpy((np.roll(base, k, 0) / 255.).astype(np.float32)).cuda() for k in range(L)]
if os.environ.get("VALUE_U8"):
    xs = [torch.from_numpy(np.ascontiguousarray(np.roll(base, k, 0))).cuda() for k in range(L)]
def e2e(n):
    for _ in pool.map(hs[i % nh] for i in range(n)):
        pass
def val(n):
    pool.fork()
    for i in range(n):
        pool.step_device(xs[i % L])
    pool.join()
    torch.cuda.synchronize()
order = ((val, "value"), (e2e, "e2e pool.map"), (val, "value again")) if os.environ.get("VALUE_FIRST") else ((e2e, "e2e pool.map"), (val, "value"))
for fn, name in order:
    fn(4 * L * D)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(S)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"lanes {L} depth {D} host buffers {nh}: {name}: {B * S / dt:.0f} images/s")
