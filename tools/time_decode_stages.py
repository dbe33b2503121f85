"""Per-stage timing of the decode chain on the forward's own maps: NMS (peaks), limbs, grouping, and the whole
chain, via the stage entries of the C ABI.  usage: time_decode_stages.py [batch ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lightweight_openpose_b200 import checkpoint, engine, synth


def timed(fn, iters=20, warm=3):
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
    H = int(os.environ.get("PROBE_H", 368))
    W = int(os.environ.get("PROBE_W", 368))
    scenes = int(os.environ.get("PROBE_SCENES", 0))          # > 0: label-style scenes with this many persons
    variables, src = checkpoint.load_or_synthesize()
    for B in [int(b) for b in (sys.argv[1:] or ["32", "256"])]:
        eng = engine.Engine(H, W, max_batch=B)
        eng.load_variables(variables, src)
        if scenes:
            hs, ps = zip(*[synth.synth_scene(1000 + s, scenes, H, W)[:2] for s in range(min(B, 16))])
            reps = -(-B // len(hs))
            heat = torch.from_numpy(np.concatenate([np.stack(hs)] * reps)[:B]).cuda()
            paf = torch.from_numpy(np.concatenate([np.stack(ps)] * reps)[:B]).cuda()
        else:
            base = np.stack([synth.synth_image(s, H, W) for s in range(32)])
            x = torch.from_numpy(np.concatenate([base] * (-(-B // 32)))[:B]).cuda()
            heat, paf = (t.clone() for t in eng.forward(x))
        h, w = heat.shape[1:3]
        rec = {"batch": B, "H": H, "W": W, "scene_persons": scenes}
        import ctypes as C
        from lightweight_openpose_b200 import _lib
        lib, hd = eng.lib, eng.handle
        peaks, pc, counts = eng.nms_device(heat, H / h, 0.1)
        limbs = eng.connect_device(paf, False, (h, w), (H, W), 0.0, peaks, pc, counts)
        joints, persons, kp = eng.group_device(peaks, pc, limbs, counts)
        t = _lib.DecodeTables(counts.data_ptr(), joints.data_ptr(), limbs.data_ptr(), persons.data_ptr(), kp.data_ptr())
        st = torch.cuda.current_stream().cuda_stream
        rec["nms_ms"] = timed(lambda: lib.lwop_nms(hd, heat.data_ptr(), B, h, w, float(H / h), 0.1, 0, peaks.data_ptr(),
                                                   pc.data_ptr(), counts.data_ptr(), st))
        lib.lwop_find_connected_joints(hd, paf.data_ptr(), 0, B, h, w, H, W, 0.0, 10, 1.0, peaks.data_ptr(), pc.data_ptr(),
                                       limbs.data_ptr(), counts.data_ptr(), st)
        rec["limbs_ms"] = timed(lambda: lib.lwop_find_connected_joints(hd, paf.data_ptr(), 0, B, h, w, H, W, 0.0, 10, 1.0,
                                                                       peaks.data_ptr(), pc.data_ptr(), limbs.data_ptr(),
                                                                       counts.data_ptr(), st))
        rec["group_ms"] = timed(lambda: lib.lwop_group_limbs(hd, peaks.data_ptr(), pc.data_ptr(), B, C.byref(t), st))
        rec["chain_ms"] = timed(lambda: eng.decode_device(heat, paf, (H, W)))
        res = eng.fetch(B)
        rec["joints_per_image"] = float(res.counts[:, 0].mean())
        rec["max_joints"] = int(res.counts[:, 0].max())
        rec["persons_per_image"] = float(res.counts[:, 1].mean())
        rec["conns_per_image"] = float(res.counts[:, 3].mean())
        rec["max_peaks_channel"] = int(res.counts[:, 4:18].max())
        rec["cand_frac"] = float((heat > 0.1).float().mean())
        print(json.dumps(rec), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
