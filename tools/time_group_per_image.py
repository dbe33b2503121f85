"""Per-image time of the grouping stage (load imbalance check): python tools/time_group_per_image.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightweight_openpose_b200 import checkpoint, engine  # noqa: E402
import bench  # noqa: E402

B = 64
eng = engine.Engine(368, 368, max_batch=B)
v, src = checkpoint.load_or_synthesize()
eng.load_variables(v, src)
x = torch.from_numpy((bench.make_frames(B) / 255.).astype(np.float32)).cuda()
heat, paf = (t.clone() for t in eng.forward(x))
peaks, pc, counts = eng.nms_device(heat, 8.0, 0.1)
limbs = eng.connect_device(paf, False, (46, 46), (368, 368), 0.0, peaks, pc, counts)
torch.cuda.synchronize()
res = []
for i in range(B):
    a = (peaks[i:i + 1].contiguous(), pc[i:i + 1].contiguous(), limbs[i:i + 1].contiguous(), counts[i:i + 1].clone())
    for _ in range(2):
        eng.group_device(*a[:3], a[3].clone())
    c = a[3].clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.group_device(a[0], a[1], a[2], c)
    e1.record()
    torch.cuda.synchronize()
    cc = c.cpu().numpy()[0]
    res.append((e0.elapsed_time(e1) * 1e3, int(cc[0]), int(cc[1]), int(cc[3])))
res.sort()
print("us, joints, persons, connections: fastest", res[:3], "slowest", res[-6:])
print("median us", res[len(res) // 2][0])
