"""BASELINE.json configs[4]: standalone decode on N synthetic scenes (1-20 persons, seed = frame index), CUDA chain vs
the CPU oracle: integer outputs (peak x/y, ids, limb pairs, grouping, keypoint table) must be exact, float scores
within 1e-5.  Writes a one-line JSON summary (mismatch rate, max float difference).

    python tools/decode_parity_10k.py --frames 10000 --out profiles/r1_decode_parity_10k.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARAM = {"thre1": 0.1, "thre2": 0.0}
HW = (368, 368)
USE_IPP = os.environ.get("LWOP_ORACLE_IPP", "0") == "1"      # default: OpenCV's own arithmetic (bit-exact target)


def make_scene(i):
    from lightweight_openpose_b200 import synth
    return synth.synth_scene(i, 1 + i % 20, HW[0], HW[1])[:2]


def oracle_one(i):
    from oracle import decode_oracle as orc
    import cv2
    cv2.setNumThreads(1)
    cv2.ipp.setUseIPP(USE_IPP)
    h, p = make_scene(i)
    d = orc.decode_detail(HW, PARAM, h, p)
    return i, h, p, d


def run(frames, procs=None, chunk=256, time_limit=1500.0):
    """The parity loop; returns the summary dict.  Oracle workers are spawned (not forked): the calling process may
    already hold a CUDA context (pytest)."""
    import multiprocessing as mp
    from types import SimpleNamespace
    args = SimpleNamespace(frames=frames, procs=procs or (os.cpu_count() or 1), chunk=chunk, time_limit=time_limit)
    FLOAT_TOL = 1e-5 if USE_IPP else 0.0
    from lightweight_openpose_b200.src import pose_decode
    from test_oracle_decode import compare_decode
    t0 = time.time()
    checked = bad_int = 0
    max_float = 0.0
    persons = joints = 0
    bad_frames = []
    with mp.get_context("spawn").Pool(args.procs) as pool:
        for c0 in range(0, args.frames, args.chunk):
            if time.time() - t0 > args.time_limit:
                break
            idx = list(range(c0, min(args.frames, c0 + args.chunk)))
            refs = pool.map(oracle_one, idx, chunksize=max(1, len(idx) // (4 * args.procs)))
            heat = np.stack([r[1] for r in refs])
            paf = np.stack([r[2] for r in refs])
            got = pose_decode.decode_pose_batch(HW, PARAM, heat, paf)
            for (i, _, _, ref), g in zip(refs, got):
                gd = {"joint_list": g.joint_list, "limbs": g.limbs, "persons": g.person_to_joint_assoc, "joints": g.joints}
                try:
                    compare_decode(gd, ref, float_tol=FLOAT_TOL, f64_tol=None if USE_IPP else 1e-12)
                except AssertionError:
                    bad_int += 1
                    bad_frames.append(i)
                if g.joint_list.size and np.asarray(ref["joint_list"]).size == g.joint_list.size:
                    max_float = max(max_float, float(np.abs(g.joint_list[:, 2] - np.asarray(ref["joint_list"])[:, 2]).max()))
                persons += g.person_to_joint_assoc.shape[0] if g.person_to_joint_assoc.size else 0
                joints += g.joint_list.shape[0] if g.joint_list.size else 0
                checked += 1
    line = {"config": "standalone decode, 368x368 synthetic scenes, 1-20 persons, seed = frame index, thre1 0.1 thre2 0.0",
            "frames_requested": args.frames, "frames_checked": checked, "frames_with_any_mismatch": bad_int,
            "mismatch_rate": bad_int / max(checked, 1), "first_bad_frames": bad_frames[:10],
            "max_abs_joint_score_diff": max_float, "joints": joints, "persons": persons,
            "oracle_cv2_ipp": USE_IPP,
            "tolerance": "integer fields exact; joint scores " + ("<= 1e-5 (IPP patch resize)" if USE_IPP else
                                                                  "bit-exact (OpenCV's own bicubic)") +
                         ("; limb/person scores <= 1e-5 / 3.2e-4 (they sum IPP-resized joint scores)" if USE_IPP else
                          "; float64 limb/person scores <= 1e-12"),
            "oracle_procs": args.procs,
            "seconds": round(time.time() - t0, 1)}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--time-limit", type=float, default=1500.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    line = run(args.frames, args.procs, args.chunk, args.time_limit)
    print(json.dumps(line))
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
