"""Where the peak kernel's time goes: empty maps (load + scan only), real maps without / with refinement.
    python tools/time_peaks.py [batch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightweight_openpose_b200 import _lib, checkpoint, engine  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = engine.Engine(368, 368, max_batch=B)
v, src = checkpoint.load_or_synthesize()
eng.load_variables(v, src)
x = torch.from_numpy((bench.make_frames(B) / 255.).astype(np.float32)).cuda()
heat, paf = (t.clone() for t in eng.forward(x))
empty = torch.zeros_like(heat)


def t(name, hm, mode):
    for _ in range(3):
        eng.nms_device(hm, 8.0, 0.1, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record()
        peaks, pc, counts = eng.nms_device(hm, 8.0, 0.1, mode)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name}: {min(ts) * 1e3:.0f} us (incl. ~3 output allocations + zero kernel); peaks/img {float(pc.sum()) / B:.1f}, max per channel {int(pc.max())}")


t("empty maps", empty, 0)
t("real maps, no refinement", heat, _lib.PEAKS_NO_REFINE)
t("real maps, refinement", heat, 0)
