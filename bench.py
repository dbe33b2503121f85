#!/usr/bin/env python
"""Benchmark of the lightweight_openpose hot path: 368x368 forward + decode images/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one pass of the whole path -- network forward (bf16 tensor-core kernels) + on-device pose decode -- over
ONE batch of 256 synthetic frames (BASELINE.json configs[2]).  With N GPUs (one process per GPU under torchrun) the
batch is sharded 256/N frames per rank (strong scaling); there is no data-path collective, the per-image result
tables are gathered on rank 0 through host shared memory.  NCCL carries only the barrier and the max-reduction of
the elapsed time.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "368x368 forward+decode images/sec"
print_json = print
H = W = 368
PARAM = {"thre1": 0.1, "thre2": 0.0}

# algorithmic work per 368x368 image
GFLOP_TOTAL = 17.639
GFLOP_TENSOR = 17.453


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region: NVML from a thread every ~2 ms (the timed region
    of a sharded run lasts only tens of milliseconds), `nvidia-smi -lms` as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_s: float = 0.002):
        self.period = period_s
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                ent = vis.split(",")[self.gpu].strip()
                if ent.isdigit():
                    idx = int(ent)
                else:
                    uuid = ent
            self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                reasons = int(fn(self.handle))
                self.samples.append((mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            n = self.nvml
            # every bit of NVML's clocks-event-reasons mask seen under load is named (gpu_idle aside: it is what an idle
            # sample between two steps reports), so that clocks below the maximum never come without their reason
            names = {"applications_clocks_setting": 0x2, "sw_power_cap": 0x4, "hw_slowdown": 0x8, "sync_boost": 0x10,
                     "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80,
                     "display_clock_setting": 0x100}
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.samples))
            mask = 0
            for _, r in self.samples:
                mask |= r
            sm = [m for m, _ in self.samples]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml", "reasons_mask_or": hex(mask),
                    "sm_mhz_min_max": [min(sm), max(sm)] if sm else None}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def make_frames(batch: int, distinct: int = 32):
    """[batch, H, W, 3] uint8 synthetic frames: `distinct` seeded frames, tiled."""
    from lightweight_openpose_b200 import synth
    base = np.stack([synth.synth_image(s, H, W) for s in range(distinct)])
    reps = -(-batch // distinct)
    return np.concatenate([base] * reps)[:batch]


# ---------------------------------------------------------------------------
# CPU arm: oracle port of the reference path (torch-CPU forward + numpy/cv2 decode)
# ---------------------------------------------------------------------------
def cpu_path_images_per_s(n_images: int, threads: int):
    import torch
    from lightweight_openpose_b200 import checkpoint, synth
    from oracle.forward_oracle import ForwardOracle
    from oracle import decode_oracle
    torch.set_num_threads(threads)
    variables, _ = checkpoint.load_or_synthesize()
    orc = ForwardOracle(variables)
    frames = [synth.synth_image(s, H, W) for s in range(n_images)]
    img0 = np.zeros((H, W, 3), np.uint8)
    orc((frames[0] / 255.).astype(np.float32)[None])          # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    for f in frames:                                          # serial per-frame loop, like model_test.py:82-99
        x = (f / 255.).astype(np.float32)[None]
        heat, paf = orc(x)
        decode_oracle.decode(img0, PARAM, heat[0], paf[0], draw=False)
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def workload_name(global_batch: int) -> str:
    """config.workload, shared by both arms so the driver compares like with like."""
    return (f"one batch of {global_batch} frames per step, 368x368 forward+decode, batch-sharded over the GPUs of the "
            "run with a host-side result gather on rank 0 (BASELINE.json configs[2]); no collective on the data path")


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(2, min(args.cpu_images, 1600 // max(args.steps + args.warmup, 1)))     # whole run: <= ~90 s of CPU work
    vals = []
    for _ in range(args.warmup):
        cpu_path_images_per_s(sample, threads)
    t_total = 0.0
    for _ in range(args.steps):
        v, dt = cpu_path_images_per_s(sample, threads)
        vals.append(v)
        t_total += dt
    value = sample * args.steps / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "global_batch": args.batch, "height": H,
                   "width": W, "thre1": PARAM["thre1"], "thre2": PARAM["thre2"],
                   "reference_arm": f"bounded sample of that workload: each step is the reference's batch-1 serial "
                                    f"loop (model_test.py:82-99) over {sample} of its frames; CPU port of the reference "
                                    "path (torch-CPU fp32 forward on all host threads + numpy/cv2 decode, one thread "
                                    "as in the reference loop), stand-in for TF-CPU: TensorFlow is not installable here",
                   "why_steps_and_config_differ": "a CPU step of the full 256-frame batch takes ~11 s, so the arm times a "
                                                  f"bounded {sample}-frame sample per step and clamps warm-up to 2 (each "
                                                  "warm-up is a full CPU pass); the metric is a throughput (images/s), "
                                                  "which does not depend on the sample size; it runs on the host cores "
                                                  "only, so its value is the same at every --gpus N"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} frames x {args.steps} steps"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(json.dumps(line))


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def classify_step(desc, next_desc=None):
    """Kernel class of a step = the CUDA kernel (template family) that runs it."""
    layer, is_dw, h, w, cin, cout, k, stride, dil, tc = desc
    if tc == 2:
        # sep_conv_kernel<..., NH=2>: Cout = 512 in one pass over all 512 TMEM columns (layers 6-11, tensor/L2-bound);
        # NH=1 instantiations: the narrower separable layers (HBM-bound)
        return "sep_fused_cout512" if next_desc is not None and next_desc[5] == 512 else "sep_fused_cout64_256"
    if tc == 4:
        return "head_pair_fused_1x1_1x1"
    if tc == 3:
        return None              # pointwise half of a fused separable layer: no launch of its own
    if is_dw:
        return "depthwise3x3"
    if layer == 0:
        return "conv1_3x3s2"
    return "conv3x3_halo" if k == 3 else "gemm_conv1x1"


def step_work(desc, batch):
    """(flops, algorithmic bytes) of one step for `batch` images (bf16 activations)."""
    layer, is_dw, h, w, cin, cout, k, stride, dil, tc = desc
    oh, ow = -(-h // stride), -(-w // stride)
    if is_dw:
        return 2.0 * 9 * oh * ow * cin * batch, 2.0 * (h * w * cin + oh * ow * cin) * batch
    flops = 2.0 * oh * ow * cout * cin * k * k * batch
    in_b = (4 if layer == 0 else 2) * h * w * cin
    out_b = (4 if cout in (14, 26) else 2) * oh * ow * cout
    return flops, float(in_b + out_b) * batch


def roofline_record(eng, x_dev, batch, mode, dump_steps=None):
    """Live per-kernel timing (graph-less forward, CUDA events around every launch) -> (roofline, per-class table,
    forward ms).  `achieved` = algorithmic flops (or bytes) of the dominant kernel class / its measured time."""
    descs = eng.step_descs()
    acc = np.zeros(len(descs))
    n_prof = 5
    eng.profile_forward(x_dev, mode=mode)
    for _ in range(n_prof):
        acc += np.asarray(eng.profile_forward(x_dev, mode=mode))
    acc /= n_prof
    classes = {}
    rows = []
    for i, (d, ms) in enumerate(zip(descs, acc)):
        cls = classify_step(d, descs[i + 1] if i + 1 < len(descs) else None)
        if cls is None:
            continue
        c = classes.setdefault(cls, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        fl, by = step_work(d, batch)
        if d[9] == 2:        # fused: depthwise + pointwise flops; bytes = layer input + layer output only
            fl2, _ = step_work(descs[i + 1], batch)
            oh, ow = -(-d[2] // d[7]), -(-d[3] // d[7])
            fl += fl2
            by = 2.0 * (d[2] * d[3] * d[4] + oh * ow * descs[i + 1][5]) * batch
        if d[9] == 4:        # fused head pair: flops of the four 1x1 convs; bytes = shared input + two narrow outputs
            fl = sum(step_work(descs[i + k], batch)[0] for k in range(4))
            by = (2.0 * d[4] + 4.0 * (descs[i + 1][5] + descs[i + 3][5])) * d[2] * d[3] * batch
        c["ms"] += ms
        c["flops"] += fl
        c["bytes"] += by
        c["launches"] += 1
        rows.append({"layer": d[0], "kind": cls, "in_hw": [d[2], d[3]], "cin": d[4],
                     "cout": descs[i + 1][5] if d[9] == 2 else d[5], "k": d[6], "stride": d[7], "dil": d[8],
                     "ms": round(float(ms), 4), "tflops": round(fl / (ms * 1e-3) / 1e12, 1),
                     "gbs": round(by / (ms * 1e-3) / 1e9, 1)})
    if dump_steps:
        with open(dump_steps, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")
    peaks = measured_peaks()
    dom = max(classes, key=lambda k: classes[k]["ms"])
    dc = classes[dom]
    # a fused separable layer is bounded by whichever of the two rooflines is slower for its shape mix:
    # report it against the tensor roofline when its tensor time exceeds its HBM time
    tensor_bound = dom.startswith("gemm") or dom.startswith("head") or (
        dom.startswith("sep") and dc["flops"] / (peaks["bf16_tflops_sustained"] * 1e12) > dc["bytes"] / (peaks["hbm_gbs"] * 1e9))
    if tensor_bound:
        achieved = dc["flops"] / (dc["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"], "frac_burst": achieved / peaks["bf16_tflops"],
                "peak_burst": peaks["bf16_tflops"], "traffic": None}
    else:
        achieved = dc["bytes"] / (dc["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "frac_burst": achieved / peaks["hbm_gbs"], "traffic": None}
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of the same workload
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tr = json.load(fh)
        if tr.get("batch") == batch and dom in tr["traffic_bytes_per_launch"]:
            roof["traffic"] = tr["traffic_bytes_per_launch"][dom]
            roof["traffic_source"] = tr.get("source", "profiles/ncu_traffic.json (ncu --set full, one launch)")
            roof["algorithmic_bytes_per_launch"] = dc["bytes"] / dc["launches"]
    except (OSError, ValueError, KeyError):
        pass
    # the same figure for every kernel class (which roofline bounds a class: tensor when its tensor time at the sustained
    # peak exceeds its HBM time at the copy bandwidth)
    per_class_roof = {}
    for k, c in classes.items():
        t_tensor = c["flops"] / (peaks["bf16_tflops_sustained"] * 1e12)
        t_hbm = c["bytes"] / (peaks["hbm_gbs"] * 1e9)
        tb = k.startswith(("gemm_conv3", "conv3x3", "head")) or (not k.startswith(("conv1", "depthwise")) and t_tensor > t_hbm)
        if tb:
            a = c["flops"] / (c["ms"] * 1e-3) / 1e12
            per_class_roof[k] = {"bound": "tensor", "achieved_tflops": round(a, 1), "frac": round(a / peaks["bf16_tflops_sustained"], 3),
                                 "frac_burst": round(a / peaks["bf16_tflops"], 3), "ms": round(c["ms"], 4)}
        else:
            a = c["bytes"] / (c["ms"] * 1e-3) / 1e9
            per_class_roof[k] = {"bound": "hbm", "achieved_gbs": round(a, 1), "frac": round(a / peaks["hbm_gbs"], 3), "ms": round(c["ms"], 4)}
    roof["classes"] = per_class_roof
    roof["kernel"] = dom
    roof["launches_per_step"] = dc["launches"]
    roof["avg_launch_ms"] = dc["ms"] / dc["launches"]
    roof["peak_source"] = peaks["source"] + (" (sustained; frac_burst is against the burst figure)" if tensor_bound else "")
    per_class = {}
    for k, c in classes.items():
        per_class[k] = {"ms": round(c["ms"], 4), "launches": c["launches"],
                        "tflops": round(c["flops"] / (c["ms"] * 1e-3) / 1e12, 2),
                        "gbs": round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1)}
    return roof, per_class, float(acc.sum()), peaks


def latency_record(variables, source, device, mode):
    """Batch-1 latency, the reference loop's shape (model_test.py:39 placeholder [1,None,None,3], :68 one sess.run
    per frame + decode_pose): one synchronous lwop_infer_host call per uint8 host frame (upload, forward, decode,
    download), and the device-resident forward+decode graph alone."""
    import torch
    from lightweight_openpose_b200 import engine, synth
    eng = engine.Engine(H, W, max_batch=1, device=device)
    eng.load_variables(variables, source)
    frames = [torch.from_numpy(synth.synth_image(s, H, W)[None]).pin_memory() for s in range(8)]
    x_dev = [f.cuda() for f in frames]
    for i in range(10):
        eng.infer_host(frames[i % 8], mode=mode)
        eng.forward_decode(x_dev[i % 8], mode=mode)
    torch.cuda.synchronize()
    lat = []
    for i in range(200):
        t0 = time.perf_counter()
        eng.infer_host(frames[i % 8], mode=mode)
        lat.append(1e3 * (time.perf_counter() - t0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        eng.forward_decode(x_dev[i % 8], mode=mode)
    e1.record()
    torch.cuda.synchronize()
    eng.close()
    lat.sort()
    return {"host_frame_to_result_ms_median": round(lat[len(lat) // 2], 4), "host_frame_to_result_ms_p99": round(lat[197], 4),
            "device_forward_decode_ms": round(e0.elapsed_time(e1) / 200, 4), "frames": 200,
            "what": "batch 1, 368x368, one synchronous lwop_infer_host per uint8 host frame (wall clock, upload + "
                    "forward + decode + download), and back-to-back device-resident forward+decode graph launches"}


def batch64_forward_record(variables, source, device, check_frames: int):
    """BASELINE.json configs[1]: batch 64, 368x368, forward only, bf16 and the fp32 check mode; the maps of the first
    `check_frames` frames are compared with the CPU oracle (tolerances of BASELINE.json: 2e-2 bf16, 1e-4 fp32 check)."""
    import torch
    from lightweight_openpose_b200 import engine, synth
    B = 64
    eng = engine.Engine(H, W, max_batch=B, device=device)
    eng.load_variables(variables, source)
    frames = np.stack([synth.synth_image(s, H, W) for s in range(B)])
    x = torch.from_numpy((frames / 255.).astype(np.float32)).cuda()
    rec = {"batch": B, "height": H, "width": W}
    outs = {}
    for m, iters in (("bf16", 50), ("fp32", 3)):
        for _ in range(3 if m == "bf16" else 1):
            eng.forward(x, mode=m)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            heat, paf = eng.forward(x, mode=m)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rec[f"{m}_images_per_s"] = B / (ms * 1e-3)
        rec[f"{m}_ms_per_batch"] = round(ms, 4)
        outs[m] = (heat[:check_frames].cpu().numpy().copy(), paf[:check_frames].cpu().numpy().copy())
    eng.close()
    if check_frames > 0:
        from oracle.forward_oracle import ForwardOracle
        import torch as _t
        _t.set_num_threads(os.cpu_count() or 1)
        oh, op = ForwardOracle(variables)((frames[:check_frames] / 255.).astype(np.float32))
        for m in ("bf16", "fp32"):
            rec[f"{m}_max_abs_err_vs_oracle"] = float(max(np.abs(outs[m][0] - oh).max(), np.abs(outs[m][1] - op).max()))
        rec["checked_frames"] = check_frames
        rec["tolerance"] = {"bf16": 2e-2, "fp32": 1e-4}
    return rec


def dense_scene_record(variables, source, device, mode, cpu_frames: int):
    """BASELINE.json configs[3]: 656x368 frames, 20-person scenes stressing the peak search and the grouping.
    decode-only: the scenes' heat maps / PAFs (label generator) fed to the decode chain; forward+decode:
    the forward on synthetic 656x368 frames followed by the decode of those scene maps (the synthetic weights cannot
    produce 20-person maps, so the two halves are chained on separate buffers).  The CPU port on the same scenes
    stands beside it."""
    import torch
    from lightweight_openpose_b200 import engine, synth
    HH, WW, B = 368, 656, 64
    eng = engine.Engine(HH, WW, max_batch=B, device=device)
    eng.load_variables(variables, source)
    hs, ps = zip(*[synth.synth_scene(1000 + s, 20, HH, WW)[:2] for s in range(16)])
    heat = torch.from_numpy(np.concatenate([np.stack(hs)] * 4)[:B]).cuda()
    paf = torch.from_numpy(np.concatenate([np.stack(ps)] * 4)[:B]).cuda()
    base = np.stack([synth.synth_image(s, HH, WW) for s in range(16)])
    x = torch.from_numpy(np.concatenate([base] * 4)[:B]).cuda()

    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    dec_ms = timed(lambda: eng.decode_device(heat, paf, (HH, WW), PARAM["thre1"], PARAM["thre2"]))
    res = eng.fetch(B)

    def both():
        eng.forward(x, mode=mode)
        eng.decode_device(heat, paf, (HH, WW), PARAM["thre1"], PARAM["thre2"])
    both_ms = timed(both)
    rec = {"height": HH, "width": WW, "batch": B, "persons_per_image": float(res.counts[:, 1].mean()),
           "joints_per_image": float(res.counts[:, 0].mean()),
           "decode_only_images_per_s": B / (dec_ms * 1e-3), "decode_ms_per_batch": round(dec_ms, 4),
           "forward_decode_images_per_s": B / (both_ms * 1e-3), "forward_decode_ms_per_batch": round(both_ms, 4)}
    eng.close()
    if cpu_frames > 0:
        from oracle import decode_oracle
        from oracle.forward_oracle import ForwardOracle
        import torch as _t
        _t.set_num_threads(os.cpu_count() or 1)
        orc = ForwardOracle(variables)
        img0 = np.zeros((HH, WW, 3), np.uint8)
        t0 = time.perf_counter()
        for s in range(cpu_frames):
            decode_oracle.decode(img0, PARAM, hs[s % 16], ps[s % 16], draw=False)
        t_dec = (time.perf_counter() - t0) / cpu_frames
        orc((base[0] / 255.).astype(np.float32)[None])
        t0 = time.perf_counter()
        for s in range(cpu_frames):
            orc((base[s % 16] / 255.).astype(np.float32)[None])
        t_fwd = (time.perf_counter() - t0) / cpu_frames
        rec["cpu_port"] = {"decode_only_images_per_s": 1.0 / t_dec, "forward_decode_images_per_s": 1.0 / (t_dec + t_fwd),
                           "frames": cpu_frames, "cores": os.cpu_count() or 1, "kind": "port"}
    return rec


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    from lightweight_openpose_b200 import checkpoint, engine, replicas
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    G = args.batch                                   # ONE batch of G frames per step, sharded over the ranks
    lo, hi = replicas.shard_bounds(G, world, rank)
    b = hi - lo
    # steps in flight per GPU: small slices leave more of every kernel's fill / drain and of the latency-bound decode
    # chain exposed, and more concurrent steps fill those gaps (measured: 4 lanes at 256 frames; at 32 frames 16 / 24 / 32 lanes
    # give 33.4 / 33.9 / 33.8 k images/s)
    lanes = args.lanes if args.lanes > 0 else max(3, min(24, 1024 // b))
    # the result ring lives in /dev/shm (~16 kB per frame and slot): fall back to a shallow pipeline on a box whose
    # tmpfs could not hold it
    depth = args.depth
    try:
        st = os.statvfs("/dev/shm")
        if (lanes * depth + 2) * G * 16500 * 2 > st.f_bavail * st.f_frsize:
            lanes, depth = min(lanes, 4), 1
    except OSError:
        pass
    variables, source = checkpoint.load_or_synthesize()
    pool = engine.EnginePool(H, W, max_batch=b, lanes=lanes, device=local_rank, depth=depth)
    pool.load_variables(variables, source)
    eng0 = pool.engines[0]

    frames_u8 = make_frames(G)[lo:hi]                # this rank's slice of the batch
    # device-resident inputs: enough distinct float32 copies that a buffer is re-read only after > L2 of other inputs
    # (on top of the ~19 MB of activations per frame written in between); buffer k always goes to lane k % lanes
    in_bytes = b * H * W * 3 * 4
    n_in = lanes * max(1, -(-int(160e6 // in_bytes + 1) // lanes))
    x_host = torch.from_numpy((frames_u8 / 255.).astype(np.float32)).pin_memory()      # reference feed: img / 255.
    x_devs = [torch.roll(x_host, k, 0).cuda() for k in range(n_in)]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    seq = {"dev": 0}

    def steps_device(n):
        pool.fork()
        for _ in range(n):
            pool.step_device(x_devs[seq["dev"] % n_in], args.mode, PARAM["thre1"], PARAM["thre2"])
            seq["dev"] += 1
        pool.join()

    # ---- device-resident throughput ("value"): K steps, up to `lanes` steps in flight on independent streams ----
    steps_device(n_in)                               # builds every (lane, input buffer) CUDA graph
    steps_device(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    for e in pool.engines:
        e.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    steps_device(args.steps)
    ev1.record()
    barrier()
    launches = sum(e.launch_count() for e in pool.engines)
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    res = eng0.fetch(b)                  # untimed: decode load of the workload, for the config record
    joints_per_img = float(res.counts[:, 0].mean())
    persons_per_img = float(res.counts[:, 1].mean())

    # the same with ONE step in flight (each step's kernels start after the previous step's have finished)
    def one_at_a_time(n):
        for i in range(n):
            eng0.forward_decode(x_devs[(i * lanes) % n_in], args.mode, PARAM["thre1"], PARAM["thre2"])
    cur = torch.cuda.current_stream()
    eng0.stream.wait_stream(cur)
    one_at_a_time(3)
    barrier()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record(eng0.stream)
    one_at_a_time(args.steps)
    es1.record(eng0.stream)
    barrier()
    ms_serial = es0.elapsed_time(es1)

    # ---- end to end through the host-buffer C entries (lwop_infer_host_begin / _end) + host-side gather ----
    # Frames start in pinned HOST memory as uint8 (what cv2.imread / VideoCapture hand the reference loop,
    # model_test.py:59-65); every step every rank uploads its slice, runs `/ 255.` + forward + decode on the device and
    # downloads its result tables straight into its region of a shared-memory ring; rank 0 reads the tables of ALL
    # ranks of a step in image order (the host-side gather of SURVEY 8(e)) inside the timed region.  Up to `lanes`
    # steps are in flight per rank, the way a serving loop drives the library.
    # segment name: a token drawn by rank 0 and broadcast, so that no rank can attach to a stale segment of an earlier
    # (crashed) run that used the same rendezvous port
    tok = torch.tensor([int.from_bytes(os.urandom(6), "little")], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.broadcast(tok, src=0)
    ring_name = f"lwop_bench_{int(tok.item()):x}"
    ring = replicas.SharedResultRing(ring_name, G, world, rank, slots=pool.capacity + 2)
    n_host = pool.capacity + 1
    u8_host = [torch.from_numpy(np.ascontiguousarray(np.roll(frames_u8, k, 0))).pin_memory() for k in range(n_host)]
    f32_host = [x_host] + [torch.roll(x_host, k, 0).pin_memory() for k in range(1, n_host)] if args.e2e_f32 else None
    gathered = {"persons": 0, "images": 0}

    host_s = {"submit": 0.0, "wait_result": 0.0, "gather": 0.0, "slot": 0.0}     # host time per phase (BENCH_TRACE=1)

    def finish(st):
        t0 = time.perf_counter()
        pool.result()
        ring.publish(st)
        t1 = time.perf_counter()
        if rank == 0:
            views = ring.gather(st)
            for v in views:              # consume: per-image person counts of every rank, in image order
                gathered["persons"] += int(v["counts"][:v["batch"], 1].sum())
                gathered["images"] += v["batch"]
            ring.release(st)
        host_s["wait_result"] += t1 - t0
        host_s["gather"] += time.perf_counter() - t1

    def run_e2e(bufs, n):
        pending = []
        for _ in range(n):
            st = seq.setdefault("e2e", 0)
            seq["e2e"] = st + 1
            if pool.in_flight == pool.capacity:
                finish(pending.pop(0))
            t0 = time.perf_counter()
            ring.wait_slot_free(st)
            t1 = time.perf_counter()
            pool.submit(bufs[st % len(bufs)], args.mode, PARAM["thre1"], PARAM["thre2"],
                        tables=None if os.environ.get("BENCH_NO_RING") else ring.tables(st))
            host_s["slot"] += t1 - t0
            host_s["submit"] += time.perf_counter() - t1
            pending.append(st)
        while pending:
            finish(pending.pop(0))

    def time_e2e(bufs, sampler=None):
        # warm-up: every engine must have seen each of its (upload slot, result lane) pairs once -- each new pair captures
        # a CUDA graph (eager pass + synchronisation), which must not happen inside the timed region
        run_e2e(bufs, 2 * pool.capacity + 2)
        barrier()
        for k in host_s:
            host_s[k] = 0.0
        if sampler is not None:
            sampler.start()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pool.fork()
        run_e2e(bufs, args.steps)        # returns after the last download has landed (and, on rank 0, the last gather)
        e1.record()
        wall = 1e3 * (time.perf_counter() - t0)
        barrier()
        return max(e0.elapsed_time(e1), wall), wall

    # the e2e region runs after ~0.3 s of load, usually at the power-capped clock, while the short `value` region before it
    # may still have boosted: its clocks are sampled separately (timed part only)
    sampler_e2e = ClockSampler(local_rank, period_s=0.01)
    e2e_ms, e2e_wall_ms = time_e2e(u8_host, sampler_e2e)
    clocks_e2e = sampler_e2e.stop()
    if os.environ.get("BENCH_TRACE") and rank == 0:
        print("host ms per step:", {k: round(1e3 * v / args.steps, 4) for k, v in host_s.items()},
              "wall ms per step:", round(e2e_wall_ms / args.steps, 4), file=sys.stderr)
    d2h = eng0.last_d2h_bytes
    h2d = int(u8_host[0].numel())
    e2e_f32_ms = time_e2e(f32_host)[0] if args.e2e_f32 else 0.0
    # one synchronous call per step (no overlap across steps)
    for _ in range(2):
        eng0.infer_host(u8_host[0], mode=args.mode)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng0.infer_host(u8_host[0], mode=args.mode)
    e2e_sync_ms = 1e3 * (time.perf_counter() - t0)
    barrier()

    # ---- max over ranks; totals over ranks ----
    t = torch.tensor([ms_total, e2e_ms, e2e_f32_ms, e2e_sync_ms, ms_serial], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(launches), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, e2e_f32_ms, e2e_sync_ms, ms_serial = (float(v) for v in t.cpu())
    launches, h2d, d2h = (int(v) for v in tot.cpu())
    ring_bytes = ring.nbytes
    ring.close()

    if rank == 0:
        roof, per_class, fwd_ms, peaks = roofline_record(eng0, x_devs[0], b, args.mode, args.dump_steps)
        # ---- CPU baseline on a bounded sample (rank 0, N = 1 contract; cheap enough to always run) ----
        cpu = None
        if not args.no_cpu and world == 1:          # the CPU baseline is an N = 1 line (contract); skipped under torchrun
            threads = os.cpu_count() or 1
            v, dt = cpu_path_images_per_s(args.cpu_images, threads)
            cpu = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_images} synthetic 368x368 frames, batch-1 serial forward (torch-CPU fp32, "
                             f"{threads} threads) + decode (numpy/cv2, 1 thread), {dt:.1f} s"}
        extras = {}
        if world == 1 and not args.no_extras:
            extras["latency_b1"] = latency_record(variables, source, local_rank, args.mode)
            extras["configs1_batch64_forward_only"] = batch64_forward_record(variables, source, local_rank,
                                                                               0 if args.no_cpu else 2)
            extras["configs3_656x368_20_person_scenes"] = dense_scene_record(variables, source, local_rank, args.mode,
                                                                            0 if args.no_cpu else 4)

        ms_per_step = ms_total / args.steps
        total_images = G * args.steps
        value = total_images / (ms_total * 1e-3)
        step_tflops = GFLOP_TENSOR * G / ms_per_step / world        # per GPU, whole step (decode included in the time)
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": workload_name(G),
                "global_batch": G, "frames_per_gpu_per_step": b, "height": H, "width": W, "mode": args.mode,
                "steps_in_flight_per_gpu": lanes, "e2e_steps_in_flight_per_gpu": pool.capacity,
                "weights": source, "thre1": PARAM["thre1"], "thre2": PARAM["thre2"],
                "l2": f"each GPU cycles through {n_in} distinct float32 input buffers ({n_in * in_bytes / 2**20:.0f} MiB "
                      "> 126 MB L2) and writes ~19 MB of activations per frame between two reads (no flush needed)",
                "decode_load": {"joints_per_image": joints_per_img, "persons_per_image": persons_per_img},
                "one_step_at_a_time": {"value": total_images / (ms_serial * 1e-3), "unit": "images/s",
                                       "ms_per_step": ms_serial / args.steps},
                "forward_ms_sum_of_kernels": round(fwd_ms, 3),
                "forward_tflops_tensor": round(GFLOP_TENSOR * b / fwd_ms, 1) if fwd_ms > 0 else None,
                "whole_step_tensor_frac": {"tflops_per_gpu": round(step_tflops, 1),
                                           "of_sustained": round(step_tflops / peaks["bf16_tflops_sustained"], 4),
                                           "of_burst": round(step_tflops / peaks["bf16_tflops"], 4)},
                "kernel_classes": per_class,
                "e2e_input": "uint8 frames in pinned host memory (the reference loop's frames before `/ 255.`), "
                             f"{lanes} steps in flight per GPU through lwop_infer_host_begin/_end; result tables land in a "
                             "shared-memory ring and rank 0 reads every rank's tables of each step in image order "
                             "inside the timed region",
                "e2e_gathered_on_rank0": {"images": gathered["images"], "persons": gathered["persons"],
                                          "ring_bytes": ring_bytes},
                "e2e_float32_frames": ({"value": total_images / (e2e_f32_ms * 1e-3), "unit": "images/s",
                                        "h2d_bytes_per_step": G * H * W * 3 * 4} if args.e2e_f32 else None),
                "e2e_one_synchronous_call_per_step": {"value": total_images / (e2e_sync_ms * 1e-3), "unit": "images/s"},
                "e2e_wall_ms_per_step": e2e_wall_ms / args.steps,
                "e2e_clocks": clocks_e2e,
                **extras,
            },
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": total_images / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print_json(json.dumps(line))
    pool.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm (rank 0 only) is meant to use all host cores,
    # so the variable is dropped before torch is imported.  stdout carries exactly ONE JSON line: everything else
    # that libraries print there (NCCL's version banner) is routed to stderr.
    os.environ.pop("OMP_NUM_THREADS", None)
    os.environ.pop("MKL_NUM_THREADS", None)
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str) -> None:
        os.write(real_stdout, (line + "\n").encode())

    global print_json
    print_json = emit
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per step (ONE batch, sharded over the GPUs)")
    ap.add_argument("--lanes", type=int, default=0,
                    help="steps in flight per GPU (independent engines / CUDA streams); 0 = by slice size: 1024 / frames, 3..16")
    ap.add_argument("--depth", type=int, default=2, help="host batches in flight per engine of the pool (1 or 2)")
    ap.add_argument("--e2e-f32", action="store_true", help="also time the host path with float32 frames (4x the upload)")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch-1 latency and the configs[3] sub-records")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "bf16_direct"])
    ap.add_argument("--cpu-images", type=int, default=224,
                    help="frames in the bounded CPU-baseline sample (~11 s at ~20 images/s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="separable layers as two kernels (A/B timing of the fusion)")
    ap.add_argument("--dump-steps", default=None, help="write per-kernel timings (one JSON object per line) here")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        args.warmup = min(args.warmup, 2)        # each warm-up is one full CPU frame
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it so that there is one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
