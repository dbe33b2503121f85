"""A few launches of the fused head-pair kernel (for ncu): python tools/profile_heads_one.py hidden [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightweight_openpose_b200 import _lib, engine  # noqa: E402

hidden = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
M = B * 46 * 46
eng = engine.Engine(16, 16, max_batch=1, device=0, use_graph=False)
x = torch.randn((M, 128), device="cuda").to(torch.bfloat16)
w = [torch.randn((hidden, 128), device="cuda") / 11, torch.randn((hidden,), device="cuda"),
     torch.randn((14, hidden), device="cuda") / hidden ** 0.5, torch.randn((14,), device="cuda"),
     torch.randn((hidden, 128), device="cuda") / 11, torch.randn((hidden,), device="cuda"),
     torch.randn((26, hidden), device="cuda") / hidden ** 0.5, torch.randn((26,), device="cuda")]
oa = torch.empty((M, 14), device="cuda")
ob = torch.empty((M, 26), device="cuda")
for _ in range(3):
    rc = eng.lib.lwop_head_pair(eng.handle, x.data_ptr(), M, hidden, *[t.data_ptr() for t in w], oa.data_ptr(),
                                ob.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
torch.cuda.synchronize()
print("ok")
