"""Debug aid: print the milestone timeline of the pipelined host-buffer path (LWOP_TRACE=1)."""
import os, sys, time
os.environ["LWOP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lightweight_openpose_b200 import checkpoint, engine
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = engine.Engine(368, 368, max_batch=B)
v, src = checkpoint.load_or_synthesize()
eng.load_variables(v, src)
eng.set_option("chunk", chunk)
frames = bench.make_frames(B)
x = torch.from_numpy((frames / 255.).astype(np.float32)).pin_memory()
for i in range(4):
    print("--- call", i, file=sys.stderr)
    t0 = time.perf_counter()
    eng.infer_host(x)
    print(f"wall {1e3*(time.perf_counter()-t0):.3f} ms", file=sys.stderr)
# decode-only timing on device
xd = x.cuda()
heat, paf = eng.forward(xd)
torch.cuda.synchronize()
for nb in (64, 256):
    h, p = heat[:nb].contiguous(), paf[:nb].contiguous()
    eng.decode_device(h, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.decode_device(h, p)
    e1.record()
    torch.cuda.synchronize()
    print(f"decode batch {nb}: {e0.elapsed_time(e1)/10:.3f} ms", file=sys.stderr)
# H2D alone
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    xd.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"H2D {x.numel()*4/1e6:.0f} MB: {ms:.3f} ms = {x.numel()*4/ms/1e6:.1f} GB/s", file=sys.stderr)
for nb in (64, 256):
    xb = xd[:nb].contiguous()
    eng.forward(xb); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward(xb)
    e1.record(); torch.cuda.synchronize()
    print(f"forward batch {nb}: {e0.elapsed_time(e1)/10:.3f} ms", file=sys.stderr)
