"""Repeatability stress of one fused separable layer with a residual (the shape of Cpm layer 15): 60 launches on the same
tensors, every differing launch is reported with the pixels it touched (how the slab write-after-read race was found)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightweight_openpose_b200 import _lib, engine
B, H, W, cin, cout, dil, act = 64, 46, 46, 128, 128, 1, 2
eng = engine.Engine(16, 16, max_batch=1, device=0, use_graph=False)
torch.manual_seed(0)
x = torch.randn((B, H, W, cin), device="cuda").to(torch.bfloat16)
res = torch.randn((B, H, W, cout), device="cuda").to(torch.bfloat16)
dw = torch.randn((9, cin), device="cuda") / 3
pw = torch.randn((cout, cin), device="cuda") / cin ** 0.5
b = torch.randn((cout,), device="cuda")
use_res = int(os.environ.get("USE_RES", "1"))
ref = None
nbad = 0
for run in range(60):
    out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device="cuda")
    rc = eng.lib.lwop_separable_conv2d(eng.handle, _lib.MODE_BF16, x.data_ptr(), B, H, W, cin, dw.data_ptr(), pw.data_ptr(), b.data_ptr(),
                                       cout, dil, act, res.data_ptr() if use_res else None, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
        continue
    d = (out != ref)
    if bool(d.any()):
        nbad += 1
        idx = torch.nonzero(d)
        imgs = sorted(set(idx[:, 0].tolist()))
        i0 = idx[0].tolist()
        sub = idx[idx[:, 0] == idx[0, 0]]
        print(f"run {run}: {len(idx)} elements differ in images {imgs[:10]}; first {i0} got {float(out[tuple(i0)])} ref {float(ref[tuple(i0)])}; "
              f"in that image rows {sub[:,1].min().item()}-{sub[:,1].max().item()} cols {sub[:,2].min().item()}-{sub[:,2].max().item()} "
              f"ch {sub[:,3].min().item()}-{sub[:,3].max().item()} n {len(sub)}")
print("differing runs:", nbad, "use_res", use_res)
# analysis of the last differing run: does diff[:, n] = delta * pw[n, k] for one (or a few) input channels k?
if nbad:
    d = (out.float() - ref.float())
    idx = torch.nonzero(d.abs().sum(-1) > 0)
    for (bi, yi, xi) in idx[:6].tolist():
        dv = d[bi, yi, xi]                      # [cout]
        coef = (pw.t() @ dv) / (pw * pw).sum(0)      # per-k least squares
        resid = [(float((dv - coef[k] * pw[:, k]).norm() / dv.norm()), k) for k in range(cin)]
        resid.sort()
        # two-channel (pair) fit
        k0 = resid[0][1] & ~1
        A = pw[:, k0:k0 + 2]
        sol = torch.linalg.lstsq(A, dv.unsqueeze(1)).solution.squeeze(1)
        r2 = float((dv - A @ sol).norm() / dv.norm())
        print(f"pixel b={bi} y={yi} x={xi}: |diff| {float(dv.norm()):.3f} best single k={resid[0][1]} rel.resid {resid[0][0]:.3f}; pair k={k0},{k0+1} coef {sol.tolist()} rel.resid {r2:.3f}; resid-row corr {float(torch.dot(dv, res[bi,yi,xi].float())/ (dv.norm()*res[bi,yi,xi].float().norm())):.3f}")
