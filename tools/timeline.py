"""Kernel timeline of the pooled path (CUPTI through torch.profiler; nsys is not in the image).

usage: timeline.py [batch] [lanes] [mode: value|e2e] [steps]   ->  gpurun_out/timeline_<mode>_b<batch>_l<lanes>.csv
columns: start_us (relative), dur_us, stream, name.  Also prints the union busy time, the time with >= 2 kernels
resident at once and per-kernel-family totals.  Numbers taken under the profiler are for reading the schedule, not
for reporting.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from lightweight_openpose_b200 import checkpoint, engine, synth


def short(name):
    for k in ("gemm_conv2sm", "gemm_conv_kernel", "sep_pair2", "sep_conv", "head_pair", "dw_tma", "conv1", "stem", "peaks_kernel",
              "limbs_kernel", "group_fast", "group_kernel", "pack_offsets", "pack_rows", "zero_counts", "Memcpy", "Memset"):
        if k in name:
            return k
    return name[:40]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    mode = sys.argv[3] if len(sys.argv) > 3 else "value"
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 9
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    H = W = 368
    variables, src = checkpoint.load_or_synthesize()
    pool = engine.EnginePool(H, W, max_batch=B, lanes=lanes, depth=depth)
    pool.load_variables(variables, src)
    base = np.stack([synth.synth_image(s, H, W) for s in range(32)])
    frames = np.concatenate([base] * (-(-B // 32)))[:B]
    xs = [torch.from_numpy(np.roll(frames, k, 0)).cuda() for k in range(lanes)]
    hs = [torch.from_numpy(np.ascontiguousarray(np.roll(frames, k, 0))).pin_memory() for k in range(lanes * depth + 1)]

    def run(n):
        if mode == "value":
            pool.fork()
            for i in range(n):
                pool.step_device(xs[i % lanes])
            pool.join()
        else:
            for _ in pool.map(hs[i % len(hs)] for i in range(n)):
                pass
        torch.cuda.synchronize()

    run(2 * lanes * depth)
    run(2 * lanes * depth)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(steps)
    path = "/tmp/trace.json"
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    out = f"gpurun_out/timeline_{mode}_b{B}_l{lanes}_d{depth}.csv"
    os.makedirs("gpurun_out", exist_ok=True)
    with open(out, "w") as fh:
        fh.write("start_us,dur_us,stream,name\n")
        for e in ev:
            fh.write(f"{e['ts'] - t0:.2f},{e['dur']:.2f},{e['args'].get('stream')},{short(e['name'])}\n")
    # union / overlap
    pts = []
    for e in ev:
        if e.get("cat") != "kernel":
            continue
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    busy = over = 0.0
    depth, last = 0, pts[0][0]
    for t, d in pts:
        if depth >= 1:
            busy += t - last
        if depth >= 2:
            over += t - last
        depth += d
        last = t
    span = ev[-1]["ts"] + ev[-1]["dur"] - t0
    fam = {}
    for e in ev:
        f = fam.setdefault(short(e["name"]), [0, 0.0])
        f[0] += 1
        f[1] += e["dur"]
    print(json.dumps({"batch": B, "lanes": lanes, "mode": mode, "steps": steps, "span_us": round(span, 1),
                      "us_per_step": round(span / steps, 1), "kernel_busy_us": round(busy, 1),
                      "two_or_more_kernels_us": round(over, 1),
                      "family_us_per_step": {k: [v[0] // steps, round(v[1] / steps, 1)] for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])}}))
    pool.close()


if __name__ == "__main__":
    main()
