"""Small-batch / multi-stream probe (one GPU).

For each batch size: ms per step of (a) forward alone, (b) forward + decode on one stream, (c) the same with
`nlanes` independent engines on their own streams, steps issued round-robin (tests whether concurrent forwards fill
the per-kernel fill/drain gaps of the persistent kernels).  Prints one JSON object per configuration.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lightweight_openpose_b200 import checkpoint, engine, synth


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    H, W = (int(os.environ.get("PROBE_H", 368)), int(os.environ.get("PROBE_W", 368)))
    batches = [int(b) for b in (sys.argv[1:] or ["1", "8", "32", "64", "256"])]
    variables, src = checkpoint.load_or_synthesize()
    base = np.stack([synth.synth_image(s, H, W) for s in range(32)])
    for B in batches:
        frames = np.concatenate([base] * (-(-B // 32)))[:B]
        iters = max(10, min(200, 4096 // B))
        lanes = []
        for k in range(3):
            eng = engine.Engine(H, W, max_batch=B)
            eng.load_variables(variables, src)
            x = torch.from_numpy(frames).cuda()
            st = torch.cuda.Stream()
            lanes.append((eng, x, st))
        eng, x, _ = lanes[0]

        def fwd():
            eng.forward(x)

        def fwd_dec():
            heat, paf = eng.forward(x)
            eng.decode_device(heat, paf, (H, W))

        def make_multi(n):
            ev = [torch.cuda.Event() for _ in range(n)]
            state = {"i": 0}

            def step():
                k = state["i"] % n
                state["i"] += 1
                e, xx, st = lanes[k]
                with torch.cuda.stream(st):
                    heat, paf = e.forward(xx)
                    e.decode_device(heat, paf, (H, W))
                    ev[k].record(st)
            def run():
                step()
            def join():
                for k in range(n):
                    torch.cuda.current_stream().wait_event(ev[k])
            return run, join

        rec = {"batch": B, "H": H, "W": W}
        rec["fwd_ms"] = timed(fwd, iters)
        rec["fwd_dec_ms"] = timed(fwd_dec, iters)
        for n in (2, 3):
            run, join = make_multi(n)
            for _ in range(6):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(n):      # lanes start after e0
                lanes[k][2].wait_event(e0)
            for _ in range(iters):
                run()
            join()
            e1.record()
            torch.cuda.synchronize()
            rec[f"lanes{n}_ms"] = e0.elapsed_time(e1) / iters
        rec["img_s_fwd_dec"] = B / rec["fwd_dec_ms"] * 1e3
        rec["img_s_lanes2"] = B / rec["lanes2_ms"] * 1e3
        rec["img_s_lanes3"] = B / rec["lanes3_ms"] * 1e3
        print(json.dumps(rec), flush=True)
        for e, _, _ in lanes:
            e.close()


if __name__ == "__main__":
    main()
