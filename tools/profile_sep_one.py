"""One fused separable layer shape, a few launches (for ncu): python tools/profile_sep_one.py H W cin cout dil [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightweight_openpose_b200 import _lib, engine  # noqa: E402

H, W, cin, cout, dil = (int(v) for v in sys.argv[1:6])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 256
eng = engine.Engine(16, 16, max_batch=1, device=0, use_graph=False)
x = torch.randn((B, H, W, cin), device="cuda").to(torch.bfloat16)
out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device="cuda")
dw = torch.randn((9, cin), device="cuda") / 3
pw = torch.randn((cout, cin), device="cuda") / cin ** 0.5
b = torch.randn((cout,), device="cuda")
for _ in range(3):
    rc = eng.lib.lwop_separable_conv2d(eng.handle, _lib.MODE_BF16, x.data_ptr(), B, H, W, cin, dw.data_ptr(),
                                       pw.data_ptr(), b.data_ptr(), cout, dil, 1, None, out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
torch.cuda.synchronize()
print("ok")
