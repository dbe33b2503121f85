"""A/B of kernel-selection switches: per-kernel-class forward times (lwop_profile_forward, live CUDA events, median of
several passes) and the device-resident step time, in THIS process's environment.
usage: [LWOP_...=..] ab_steps.py [batch] [passes]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from lightweight_openpose_b200 import checkpoint, engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = int(sys.argv[2]) if len(sys.argv) > 2 else 7
v, src = checkpoint.load_or_synthesize()
eng = engine.Engine(368, 368, max_batch=B)
eng.load_variables(v, src)
x = torch.from_numpy((bench.make_frames(B) / 255.).astype(np.float32)).cuda()
descs = eng.step_descs()
runs = []
for _ in range(P):
    runs.append(np.asarray(eng.profile_forward(x), dtype=np.float64))
ms = np.median(np.stack(runs), axis=0)
cls = {}
for i, d in enumerate(descs):
    if ms[i] <= 0:
        continue
    name = bench.classify_step(d, descs[i + 1] if i + 1 < len(descs) else None)
    if name is None:
        continue
    cls[name] = cls.get(name, 0.0) + float(ms[i])
# device-resident steps (one graph launch each), 30 back to back
for _ in range(5):
    eng.forward_decode(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(eng.stream)
for _ in range(30):
    eng.forward_decode(x)
e1.record(eng.stream)
torch.cuda.synchronize()
print(json.dumps({"env": {k: os.environ[k] for k in os.environ if k.startswith("LWOP_")}, "batch": B,
                  "forward_ms": round(float(ms.sum()), 4), "step_ms_30_back_to_back": round(e0.elapsed_time(e1) / 30, 4),
                  "classes_ms": {k: round(t, 4) for k, t in cls.items()},
                  "sep_layers_ms": [round(float(ms[i]), 4) for i, d in enumerate(descs) if ms[i] > 0 and d[9] == 2]}))
