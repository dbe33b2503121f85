"""CUDA-event timing of the fused head-pair layers inside the forward: python tools/time_heads.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightweight_openpose_b200 import checkpoint, engine  # noqa: E402

B = 256
eng = engine.Engine(368, 368, max_batch=B, device=0)
variables, source = checkpoint.load_or_synthesize()
eng.load_variables(variables, source)
x = torch.rand((B, 368, 368, 3), device="cuda")
descs = eng.step_descs()
acc = np.zeros(len(descs))
eng.profile_forward(x)
for _ in range(3):
    acc += np.asarray(eng.profile_forward(x))
acc /= 3
for d, ms in zip(descs, acc):
    if d[9] == 4:
        print("head pair layer", d[0], "hidden", d[5], "ms %.4f" % ms)
print("forward sum ms %.3f" % acc.sum())
