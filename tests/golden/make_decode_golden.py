"""Generates ``tests/golden/decode_golden.npz`` and ``decode_golden_noipp.npz`` by running the UNMODIFIED
reference decoder (``/root/reference/src/pose_decode.py``) on seeded scenes.

Run in the build container only (the reference tree does not exist on the GPU
box):  ``python tests/golden/make_decode_golden.py``

Two files because the reference's ``cv2.resize`` is not one function: OpenCV builds with the (binary-only)
IPP accelerator route the single-channel 5x5 patch resize to IPP, whose float32 results differ from OpenCV's
own C++ code by 1 ulp on ~25% of the pixels (the 26-channel PAF resize takes OpenCV's own code either way).
``decode_golden.npz`` is the module run with cv2 as installed (IPP on); ``decode_golden_noipp.npz`` is the same
module after ``cv2.ipp.setUseIPP(False)`` = OpenCV's own arithmetic, which the oracle restatement and the CUDA
kernels reproduce bit for bit (scores included).

For every scene the file stores the inputs (heatmaps, pafs, image size,
thresholds) and the reference outputs: ``joint_list`` (N,6) f64, the 13 limb
tables (m,5) f64, ``person_to_joint_assoc`` (P,16) f64 and ``joints`` (P,42) f32.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from lightweight_openpose_b200.synth import synth_scene  # noqa: E402

SCENES = [
    # name, seed, persons, H, W, noise
    ("three_368", 3, 3, 368, 368, 0.02),
    ("one_368", 1, 1, 368, 368, 0.02),
    ("empty_368", 7, 0, 368, 368, 0.02),
    ("eight_368", 8, 8, 368, 368, 0.02),
    ("twenty_656x368", 20, 20, 368, 656, 0.02),
    ("two_256", 2, 2, 256, 256, 0.02),
    ("five_noisy_368", 5, 5, 368, 368, 0.12),   # noise above thre1: many spurious peaks
]


def single_peak_scene():
    heat = np.random.default_rng(99).uniform(0, 0.02, size=(46, 46, 14)).astype(np.float32)
    heat[0, 45, 3] = 0.9            # isolated corner peak (clipped 3x3 patch)
    heat[20, 20, 3] = 0.8
    paf = np.random.default_rng(98).normal(0, 0.01, size=(46, 46, 26)).astype(np.float32)
    return heat, paf


def main(use_ipp=True, fname="decode_golden.npz"):
    warnings.simplefilter("ignore")
    import cv2
    cv2.setNumThreads(1)
    cv2.ipp.setUseIPP(use_ipp)
    from src import pose_decode as ref            # the reference module itself

    param = {"thre1": 0.1, "thre2": 0.0}
    out = {}
    names = []
    for name, seed, persons, H, W, noise in SCENES + [("single_peak", 0, 0, 368, 368, 0.0)]:
        if name == "single_peak":
            heat, paf = single_peak_scene()
        else:
            heat, paf, _ = synth_scene(seed, persons, H, W, noise=noise)
        img = np.zeros((H, W, 3), dtype=np.uint8)
        # stage by stage, so limb tables are captured too (decode_pose :460-517)
        scale = img.shape[0] / heat.shape[0]
        per_type = ref.NMS(param, heat, scale)
        joint_list = np.array([tuple(peak) + (jt,) for jt, peaks in enumerate(per_type) for peak in peaks])
        paf_up = cv2.resize(paf, (W, H), interpolation=cv2.INTER_CUBIC)
        limbs = ref.find_connected_joints(param, paf_up, per_type)
        persons_arr = ref.group_limbs_of_same_person(limbs, joint_list)
        _, jl2, pa2, joints = ref.decode_pose(img, param, heat, paf)
        assert np.array_equal(np.asarray(jl2), joint_list)
        assert np.array_equal(np.asarray(pa2), persons_arr)
        names.append(name)
        out[f"{name}/heat"] = heat
        out[f"{name}/paf"] = paf
        out[f"{name}/hw"] = np.array([H, W], dtype=np.int64)
        out[f"{name}/joint_list"] = np.asarray(joint_list, dtype=np.float64)
        for t in range(13):
            out[f"{name}/limb{t}"] = np.asarray(limbs[t], dtype=np.float64).reshape(-1, 5)
        out[f"{name}/persons"] = np.asarray(persons_arr, dtype=np.float64)
        out[f"{name}/joints"] = np.asarray(joints, dtype=np.float32)
        # alternate mode of SURVEY 8(f) row 4: NMS with the Gaussian-smoothed patch (:163-164), rows (x, y, score, id, 0, type)
        gauss = ref.NMS(param, heat, scale, True, True)
        out[f"{name}/nms_gauss"] = np.array([tuple(pk) + (jt,) for jt, pks in enumerate(gauss) for pk in pks],
                                            dtype=np.float64).reshape(-1, 6)
        print(name, "joints", len(joint_list), "persons", persons_arr.shape, "conns",
              sum(len(l) for l in limbs))
    out["names"] = np.array(names)
    out["thre"] = np.array([param["thre1"], param["thre2"]])
    out["versions"] = np.array([cv2.__version__, np.__version__, str(cv2.ipp.useIPP())])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), fname)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main(True, "decode_golden.npz")
    main(False, "decode_golden_noipp.npz")
