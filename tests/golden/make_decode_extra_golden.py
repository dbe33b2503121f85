"""Generates ``tests/golden/decode_extra_golden.npz`` by running the UNMODIFIED reference decoder
(``/root/reference/src/pose_decode.py``, OpenCV's own arithmetic: ``cv2.ipp.setUseIPP(False)``) on two cases the
first golden set does not hold:

* ``dense_46x82`` -- BASELINE configs[3] frame size (656x368, maps 46x82) with EVERY heat-map pixel above ``thre1``:
  the radius-5 search keeps close to the geometric maximum of peaks per channel (~150), i.e. more than the 128 the
  round-1 tables held; the reference returns a result, so must the CUDA path.
* ``params`` -- ``find_connected_joints`` with non-default ``num_intermed_pts`` / ``max_dist_thresh``
  (reference ``src/pose_decode.py:188``) on a 6-person scene.

Run in the build container only:  ``python tests/golden/make_decode_extra_golden.py``
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from lightweight_openpose_b200.synth import synth_scene  # noqa: E402

LIMB_PARAMS = [(5, 1), (16, 1), (10, 0.25), (7, 0.5), (1, 1), (2, 1), (64, 1), (9, 1.0), (8, 1)]


def dense_scene():
    rng = np.random.default_rng(4682)
    heat = rng.uniform(0.11, 1.0, size=(46, 82, 14)).astype(np.float32)      # every pixel is a candidate
    # smooth weak PAF field + negative offset: most candidate limbs fail the 9-of-10 criterion, so the grouping
    # stays within what a frame can plausibly hold while the peak tables are as full as they can get
    paf = (rng.normal(0.0, 0.05, size=(46, 82, 26)) - 0.03).astype(np.float32)
    return heat, paf


def main():
    warnings.simplefilter("ignore")
    import cv2
    cv2.setNumThreads(1)
    cv2.ipp.setUseIPP(False)
    from src import pose_decode as ref            # the reference module itself

    param = {"thre1": 0.1, "thre2": 0.0}
    out = {}
    # ---- dense 46x82 frame ----
    heat, paf = dense_scene()
    H, W = 368, 656
    img = np.zeros((H, W, 3), dtype=np.uint8)
    per_type = ref.NMS(param, heat, H / heat.shape[0])
    joint_list = np.array([tuple(peak) + (jt,) for jt, peaks in enumerate(per_type) for peak in peaks])
    paf_up = cv2.resize(paf, (W, H), interpolation=cv2.INTER_CUBIC)
    limbs = ref.find_connected_joints(param, paf_up, per_type)
    persons = ref.group_limbs_of_same_person(limbs, joint_list)
    _, jl2, pa2, joints = ref.decode_pose(img, param, heat, paf)
    assert np.array_equal(np.asarray(jl2), joint_list) and np.array_equal(np.asarray(pa2), persons)
    out["dense_46x82/heat"], out["dense_46x82/paf"] = heat, paf
    out["dense_46x82/hw"] = np.array([H, W], dtype=np.int64)
    out["dense_46x82/joint_list"] = np.asarray(joint_list, dtype=np.float64)
    for t in range(13):
        out[f"dense_46x82/limb{t}"] = np.asarray(limbs[t], dtype=np.float64).reshape(-1, 5)
    out["dense_46x82/persons"] = np.asarray(persons, dtype=np.float64)
    out["dense_46x82/joints"] = np.asarray(joints, dtype=np.float32)
    print("dense_46x82: peaks per channel", [len(p) for p in per_type], "persons", np.asarray(persons).shape,
          "conns", sum(len(l) for l in limbs))
    # ---- find_connected_joints keyword arguments ----
    heat, paf, _ = synth_scene(31, 6, 368, 368, noise=0.04)
    per_type = ref.NMS(param, heat, 8.0)
    paf_up = cv2.resize(paf, (368, 368), interpolation=cv2.INTER_CUBIC)
    out["params/heat"], out["params/paf"] = heat, paf
    out["params/joint_list"] = np.array([tuple(pk) + (jt,) for jt, pks in enumerate(per_type) for pk in pks],
                                        dtype=np.float64).reshape(-1, 6)
    out["params/settings"] = np.array(LIMB_PARAMS, dtype=np.float64)
    for k, (n, thr) in enumerate(LIMB_PARAMS):
        lim = ref.find_connected_joints(param, paf_up, per_type, num_intermed_pts=n, max_dist_thresh=thr)
        for t in range(13):
            out[f"params/{k}/limb{t}"] = np.asarray(lim[t], dtype=np.float64).reshape(-1, 5)
        print("params", n, thr, "conns", sum(len(l) for l in lim))
    out["thre"] = np.array([param["thre1"], param["thre2"]])
    out["versions"] = np.array([cv2.__version__, np.__version__, str(cv2.ipp.useIPP())])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decode_extra_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
