"""Forward-only driver for ncu captures: a few graph-less forwards at batch B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lightweight_openpose_b200 import checkpoint, engine
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = engine.Engine(368, 368, max_batch=B, use_graph=False)
v, src = checkpoint.load_or_synthesize()
eng.load_variables(v, src)
x = torch.from_numpy((bench.make_frames(B) / 255.).astype(np.float32)).cuda()
for _ in range(n):
    eng.forward(x)
torch.cuda.synchronize()
print("ok")
