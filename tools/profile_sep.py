"""Per-role stall cycles of the fused separable kernel at the network's layer shapes (run under gpurun):
    LWOP_SEP_PROF=1 python tools/profile_sep.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightweight_openpose_b200 import _lib, engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SHAPES = [(184, 184, 32, 64, 1), (92, 92, 128, 128, 1), (46, 46, 256, 256, 1), (46, 46, 256, 512, 1),
          (46, 46, 512, 512, 2), (46, 46, 512, 512, 1), (46, 46, 128, 128, 1)]
eng = engine.Engine(16, 16, max_batch=1, device=0, use_graph=False)
for (H, W, cin, cout, dil) in SHAPES:
    x = torch.randn((B, H, W, cin), device="cuda").to(torch.bfloat16)
    out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device="cuda")
    dw = torch.randn((9, cin), device="cuda") / 3
    pw = torch.randn((cout, cin), device="cuda") / cin ** 0.5
    b = torch.randn((cout,), device="cuda")
    for _ in range(2):
        rc = eng.lib.lwop_separable_conv2d(eng.handle, _lib.MODE_BF16, x.data_ptr(), B, H, W, cin, dw.data_ptr(),
                                           pw.data_ptr(), b.data_ptr(), cout, dil, 1, None, out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, eng.handle)
    torch.cuda.synchronize()
