"""Decode-only driver for ncu: forward once (B frames), then the decode chain a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lightweight_openpose_b200 import checkpoint, engine
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = engine.Engine(368, 368, max_batch=B)
v, src = checkpoint.load_or_synthesize()
eng.load_variables(v, src)
x = torch.from_numpy((bench.make_frames(B) / 255.).astype(np.float32)).cuda()
heat, paf = eng.forward(x)
torch.cuda.synchronize()
for _ in range(3):
    eng.decode_device(heat, paf)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    eng.decode_device(heat, paf)
ev[1].record()
torch.cuda.synchronize()
print(f"decode batch {B}: {ev[0].elapsed_time(ev[1]) / 10:.3f} ms")
