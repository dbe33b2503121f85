#!/usr/bin/env python
"""Benchmark of the lightweight_openpose hot path: 368x368 forward + decode images/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One process per GPU (under torchrun for N > 1; ranks are independent replicas,
the only cross-rank traffic is the barrier and the max-reduction of the time).
A step is one pass of the whole path -- network forward (bf16 tensor-core
kernels) + on-device pose decode -- over one batch of synthetic frames.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
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
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
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
                "samples": len(sm)}


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


def workload_name(batch: int, world: int) -> str:
    """config.workload, shared by both arms so the driver compares like with like."""
    return (f"batch {batch} per GPU, 368x368 forward+decode (BASELINE.json configs[2]), "
            f"{world} independent replica(s), no collective on the data path")


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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch, world), "global_batch": args.batch * world, "height": H,
                   "width": W, "thre1": PARAM["thre1"], "thre2": PARAM["thre2"],
                   "reference_arm": f"bounded sample of that workload: each step is the reference's batch-1 serial "
                                    f"loop (model_test.py:82-99) over {sample} of its frames; CPU port of the reference "
                                    "path (torch-CPU fp32 forward on all host threads + numpy/cv2 decode, one thread "
                                    "as in the reference loop), stand-in for TF-CPU: TensorFlow is not installable here"},
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
    return "gemm_conv3x3_im2col" if k == 3 else "gemm_conv1x1"


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


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    from lightweight_openpose_b200 import checkpoint, engine
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL logs its version there)
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    batch = args.batch
    eng = engine.Engine(H, W, max_batch=batch, device=local_rank, fuse_separable=not args.no_fuse)
    variables, source = checkpoint.load_or_synthesize()
    eng.load_variables(variables, source)

    frames_u8 = make_frames(batch)
    x_host = torch.from_numpy((frames_u8 / 255.).astype(np.float32)).pin_memory()      # reference feed: img / 255.
    x_dev = x_host.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        heat, paf = eng.forward(x_dev, mode=args.mode)
        eng.decode_device(heat, paf, (H, W), PARAM["thre1"], PARAM["thre2"])
        return heat, paf

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    launches = eng.launch_count()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    res = eng.fetch(batch)               # untimed: decode load of the workload, for the config record
    joints_per_img = float(res.counts[:, 0].mean())
    persons_per_img = float(res.counts[:, 1].mean())

    if args.chunk > 0:
        eng.set_option("chunk", args.chunk)
    sweep = {}
    if args.sweep_chunks and rank == 0:
        for c in (32, 64, 128, 256):
            if c > batch:
                continue
            eng.set_option("chunk", c)
            sw = torch.from_numpy(frames_u8).pin_memory()
            for _ in range(2):
                eng.infer_host(sw, mode=args.mode)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                eng.infer_host(sw, mode=args.mode)
            torch.cuda.synchronize()
            sweep[c] = round(1e3 * (time.perf_counter() - t0) / 5, 3)
        eng.set_option("chunk", max(args.chunk, 0))
    # ---- end to end through the host-buffer C entries (lwop_infer_host_begin / _end) ----
    # Frames start in pinned HOST memory as uint8 (what cv2.imread / VideoCapture hand the reference loop,
    # model_test.py:59-65); every step uploads its frames, runs `/ 255.` + forward + decode on the device and
    # downloads the result tables.  Two batches are kept in flight (begin k+1, then end k), the way a serving loop
    # drives the library, so the upload of one batch overlaps the kernels of the previous one; all K uploads and K
    # downloads happen inside the timed region.
    u8_a = torch.from_numpy(frames_u8).pin_memory()
    u8_b = torch.from_numpy(np.ascontiguousarray(frames_u8[::-1])).pin_memory()

    def run_e2e(bufs, steps):
        ticket = eng.infer_host_begin(bufs[0], mode=args.mode)
        out = None
        for i in range(1, steps):
            nxt = eng.infer_host_begin(bufs[i % len(bufs)], mode=args.mode)
            out = eng.infer_host_end(ticket)
            ticket = nxt
        out = eng.infer_host_end(ticket)
        return out

    def time_e2e(bufs):
        run_e2e(bufs, max(2, min(args.warmup, 3)))
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(bufs, args.steps)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)     # device clock (ends after the last download)

    e2e_ms, e2e_wall_ms = time_e2e([u8_a, u8_b])
    d2h = eng.last_d2h_bytes
    h2d = int(u8_a.numel())

    # the same with float32 frames (the reference's `img / 255.` done on the host: 4x the upload) ...
    e2e_f32_ms, _ = time_e2e([x_host])
    # ... and one synchronous call per batch (no overlap across batches)
    for _ in range(2):
        eng.infer_host(u8_a, mode=args.mode)
    barrier()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(args.steps):
        eng.infer_host(u8_a, mode=args.mode)
    es1.record()
    barrier()
    e2e_sync_ms = es0.elapsed_time(es1)
    e2e_u8_ms = e2e_f32_ms      # third slot of the cross-rank reduction below

    # ---- max over ranks ----
    t = torch.tensor([ms_total, e2e_ms, e2e_f32_ms, e2e_sync_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_f32_ms, e2e_sync_ms = (float(v) for v in t.cpu())

    if rank == 0:
        # ---- live per-kernel timing for the roofline (rank 0 only, outside the timed loops) ----
        descs = eng.step_descs()
        acc = np.zeros(len(descs))
        n_prof = 5
        eng.profile_forward(x_dev, mode=args.mode)
        for _ in range(n_prof):
            acc += np.asarray(eng.profile_forward(x_dev, mode=args.mode))
        acc /= n_prof
        classes = {}
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
        peaks = measured_peaks()
        dom = max(classes, key=lambda k: classes[k]["ms"])
        dc = classes[dom]
        # a fused separable layer is bounded by whichever of the two rooflines is slower for its shape mix:
        # report it against the tensor roofline when its tensor time exceeds its HBM time
        tensor_bound = dom.startswith("gemm") or dom.startswith("head") or (dom.startswith("sep") and
                                                  dc["flops"] / (peaks["bf16_tflops_sustained"] * 1e12) >
                                                  dc["bytes"] / (peaks["hbm_gbs"] * 1e9))
        if tensor_bound:
            achieved = dc["flops"] / (dc["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None}
        else:
            achieved = dc["bytes"] / (dc["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": None}
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture of the same workload
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                tr = json.load(fh)
            if tr.get("batch") == batch and dom in tr["traffic_bytes_per_launch"]:
                roof["traffic"] = tr["traffic_bytes_per_launch"][dom]
                roof["traffic_source"] = "profiles/ncu_traffic.json (ncu --set full, one launch)"
                roof["algorithmic_bytes_per_launch"] = dc["bytes"] / dc["launches"]
        except (OSError, ValueError, KeyError):
            pass
        roof["kernel"] = dom
        roof["launches_per_step"] = dc["launches"]
        roof["avg_launch_ms"] = dc["ms"] / dc["launches"]
        roof["peak_source"] = peaks["source"] + (" (sustained)" if tensor_bound else "")
        per_class = {}
        for k, c in classes.items():
            per_class[k] = {"ms": round(c["ms"], 4), "launches": c["launches"],
                            "tflops": round(c["flops"] / (c["ms"] * 1e-3) / 1e12, 2),
                            "gbs": round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1)}
        fwd_ms = float(acc.sum())
        if args.dump_steps:
            with open(args.dump_steps, "w") as fh:
                for i, (d, ms) in enumerate(zip(descs, acc)):
                    kind = classify_step(d, descs[i + 1] if i + 1 < len(descs) else None)
                    if kind is None:
                        continue
                    fl, by = step_work(d, batch)
                    if d[9] == 2:
                        fl += step_work(descs[i + 1], batch)[0]
                        by = 2.0 * (d[2] * d[3] * d[4] + d[2] * d[3] * descs[i + 1][5]) * batch
                    if d[9] == 4:
                        fl = sum(step_work(descs[i + k], batch)[0] for k in range(4))
                        by = (2.0 * d[4] + 4.0 * (descs[i + 1][5] + descs[i + 3][5])) * d[2] * d[3] * batch
                    fh.write(json.dumps({"layer": d[0], "kind": kind, "in_hw": [d[2], d[3]], "cin": d[4],
                                         "cout": descs[i + 1][5] if d[9] == 2 else d[5], "k": d[6], "stride": d[7], "dil": d[8], "ms": round(float(ms), 4),
                                         "tflops": round(fl / (ms * 1e-3) / 1e12, 1),
                                         "gbs": round(by / (ms * 1e-3) / 1e9, 1)}) + "\n")

        # ---- CPU baseline on a bounded sample (rank 0, N = 1 contract; cheap enough to always run) ----
        cpu = None
        if not args.no_cpu and world == 1:          # the CPU baseline is an N = 1 line (contract); skipped under torchrun
            threads = os.cpu_count() or 1
            v, dt = cpu_path_images_per_s(args.cpu_images, threads)
            cpu = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_images} synthetic 368x368 frames, batch-1 serial forward (torch-CPU fp32, "
                             f"{threads} threads) + decode (numpy/cv2, 1 thread), {dt:.1f} s"}

        ms_per_step = ms_total / args.steps
        total_images = batch * world * args.steps
        value = total_images / (ms_total * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": workload_name(batch, world),
                "global_batch": batch * world, "height": H, "width": W, "mode": args.mode,
                "weights": source, "thre1": PARAM["thre1"], "thre2": PARAM["thre2"],
                "l2": f"inputs {x_dev.numel() * 4 / 2**20:.0f} MiB fp32 per step > 126 MB L2 (no flush needed)",
                "decode_load": {"joints_per_image": joints_per_img, "persons_per_image": persons_per_img},
                "forward_ms_sum_of_kernels": round(fwd_ms, 3),
                "forward_tflops_tensor": round(GFLOP_TENSOR * batch / fwd_ms, 1) if fwd_ms > 0 else None,
                "kernel_classes": per_class,
                "e2e_input": "uint8 frames in pinned host memory (the reference loop's frames before `/ 255.`), "
                             "two batches in flight through lwop_infer_host_begin/_end",
                "e2e_float32_frames": {"value": total_images / (e2e_f32_ms * 1e-3), "unit": "images/s",
                                       "h2d_bytes_per_step": int(x_host.numel() * 4)},
                "e2e_one_synchronous_call_per_batch": {"value": total_images / (e2e_sync_ms * 1e-3), "unit": "images/s"},
                "e2e_wall_ms_per_step": e2e_wall_ms / args.steps,
                "e2e_chunk_sweep_ms": sweep or None,
            },
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": total_images / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print_json(json.dumps(line))
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
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "bf16_direct"])
    ap.add_argument("--cpu-images", type=int, default=224,
                    help="frames in the bounded CPU-baseline sample (~11 s at ~20 images/s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="separable layers as two kernels (A/B timing of the fusion)")
    ap.add_argument("--chunk", type=int, default=0, help="images per pipeline stage of the host-buffer path (0 = library default)")
    ap.add_argument("--sweep-chunks", action="store_true", help="also time the host-buffer path for several chunk sizes")
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
