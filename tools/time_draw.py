"""Timing of the canvas rasteriser (lwop_draw_poses) on the decode tables of label-style scenes, beside the host cv2
drawing of the same tables.  usage: time_draw.py [batch] [persons per scene]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lightweight_openpose_b200 import engine, synth
from lightweight_openpose_b200.src import pose_decode


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    persons = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    H = int(os.environ.get("PROBE_H", 368))
    W = int(os.environ.get("PROBE_W", 368))
    hs, ps = zip(*[synth.synth_scene(1000 + s, persons, H, W)[:2] for s in range(min(B, 16))])
    reps = -(-B // len(hs))
    heat = torch.from_numpy(np.concatenate([np.stack(hs)] * reps)[:B]).cuda()
    paf = torch.from_numpy(np.concatenate([np.stack(ps)] * reps)[:B]).cuda()
    eng = engine.Engine(H, W, max_batch=B, use_graph=False)
    res = eng.decode(heat, paf, (H, W), 0.1, 0.0)
    frames = torch.randint(0, 255, (B, H, W, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(frames)
    for _ in range(3):
        eng.draw_poses(frames, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.draw_poses(frames, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fh = frames[:16].cpu().numpy()
    t0 = time.perf_counter()
    for b in range(16):
        want, _ = pose_decode.plot_pose(fh[b], res[b].joint_list, res[b].person_to_joint_assoc)
        assert np.array_equal(out[b].cpu().numpy(), want), b
    host_ms = (time.perf_counter() - t0) / 16 * 1e3
    bytes_moved = B * H * W * (3 + 3 + 4)
    print(json.dumps({"batch": B, "H": H, "W": W, "persons_per_scene": persons, "draw_ms_per_batch": round(ms, 4),
                      "frames_per_s": round(B / ms * 1e3), "GBps_frame_in_canvas_out_priority_read": round(bytes_moved / ms / 1e6, 1),
                      "host_cv2_ms_per_frame_incl_compare": round(host_ms, 3)}))


if __name__ == "__main__":
    main()
