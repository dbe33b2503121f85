"""Stall counters / timing of one fused separable layer under the activation variants the network uses:
    LWOP_SEP_PROF=1 python tools/profile_sep_act.py H W cin cout [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightweight_openpose_b200 import _lib, engine  # noqa: E402

H, W, cin, cout = (int(v) for v in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 256
eng = engine.Engine(16, 16, max_batch=1, device=0, use_graph=False)
x = torch.randn((B, H, W, cin), device="cuda").to(torch.bfloat16)
res = torch.randn((B, H, W, cout), device="cuda").to(torch.bfloat16)
out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device="cuda")
dw = torch.randn((9, cin), device="cuda") / 3
pw = torch.randn((cout, cin), device="cuda") / cin ** 0.5
b = torch.randn((cout,), device="cuda")
for name, act, r in (("relu", 1, None), ("elu", 2, None), ("elu+residual", 2, res)):
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = eng.lib.lwop_separable_conv2d(eng.handle, _lib.MODE_BF16, x.data_ptr(), B, H, W, cin, dw.data_ptr(),
                                           pw.data_ptr(), b.data_ptr(), cout, 1, act, r.data_ptr() if r is not None else None,
                                           out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        e1.record()
        _lib.check(rc, eng.handle)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, "ms (incl. per-call weight packing)", ["%.3f" % t for t in ts])
