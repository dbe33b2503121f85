"""The decode oracle (oracle/decode_oracle.py) against golden vectors produced
by the unmodified reference decoder, and -- when /root/reference exists -- the
live reference module on fresh scenes.  Integer fields must be exact, float
scores within 1e-5 (cv2's own bicubic paths differ by ~4e-7)."""
import os
import sys
import warnings

import numpy as np
import pytest

from oracle import decode_oracle as orc
from lightweight_openpose_b200.synth import synth_scene

FLOAT_TOL = 1e-5


def compare_decode(got, ref, float_tol=FLOAT_TOL, f64_tol=None):
    """``float_tol``: joint scores (float32 bicubic values).  ``f64_tol``: limb / person scores, float64 dot
    products and means whose last bit depends on numpy's BLAS kernel (FMA or not); defaults to ``float_tol``."""
    limb_tol = float_tol if f64_tol is None else f64_tol
    pers_tol = float_tol * 32 if f64_tol is None else f64_tol * 32
    jl, rjl = np.asarray(got["joint_list"]), np.asarray(ref["joint_list"])
    assert jl.shape == rjl.shape, (jl.shape, rjl.shape)
    if jl.size:
        assert np.array_equal(jl[:, [0, 1, 3, 4, 5]], rjl[:, [0, 1, 3, 4, 5]])
        assert np.abs(jl[:, 2] - rjl[:, 2]).max() <= float_tol
    for t in range(13):
        a, b = np.asarray(got["limbs"][t]).reshape(-1, 5), np.asarray(ref["limbs"][t]).reshape(-1, 5)
        assert a.shape == b.shape, (t, a.shape, b.shape)
        if a.size:
            assert np.array_equal(a[:, [0, 1, 3, 4]], b[:, [0, 1, 3, 4]]), t
            assert np.abs(a[:, 2] - b[:, 2]).max() <= limb_tol
    p, rp = np.asarray(got["persons"]), np.asarray(ref["persons"])
    assert p.shape == rp.shape, (p.shape, rp.shape)
    if p.size:
        assert np.array_equal(p[:, :14], rp[:, :14])
        assert np.array_equal(p[:, 15], rp[:, 15])
        assert np.abs(p[:, 14] - rp[:, 14]).max() <= pers_tol
    j, rj = np.asarray(got["joints"]), np.asarray(ref["joints"])
    assert j.shape == rj.shape and j.dtype == rj.dtype
    assert np.array_equal(j, rj)


@pytest.mark.parametrize("use_cv2", [True, False])
def test_oracle_matches_reference_golden(decode_golden, use_cv2):
    scenes, param = decode_golden
    assert len(scenes) >= 8
    for name, sc in scenes.items():
        got = orc.decode_detail(sc["hw"], param, sc["heat"], sc["paf"], use_cv2=use_cv2)
        compare_decode(got, sc)


def test_oracle_return_types_match_reference(decode_golden):
    scenes, param = decode_golden
    sc = scenes["empty_368"]
    img = np.zeros(sc["hw"] + (3,), dtype=np.uint8)
    canvas, jl, persons, joints = orc.decode(img, param, sc["heat"], sc["paf"])
    assert canvas.shape == img.shape and canvas.dtype == np.uint8
    assert jl.shape == (0,) and persons.shape == (0,)
    assert joints.shape == (0, 42) and joints.dtype == np.float32
    sc = scenes["three_368"]
    img = np.zeros(sc["hw"] + (3,), dtype=np.uint8)
    canvas, jl, persons, joints = orc.decode(img, param, sc["heat"], sc["paf"])
    assert jl.dtype == np.float64 and jl.shape == (42, 6)
    assert persons.dtype == np.float64 and persons.shape == (3, 16)
    assert joints.shape == (3, 42)
    assert canvas.any()          # something was drawn


def test_bicubic_restatement_close_to_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    patch = rng.random((5, 5), dtype=np.float32)
    a = cv2.resize(patch, None, fx=8.0, fy=8.0, interpolation=cv2.INTER_CUBIC)
    b = orc.bicubic_resize_f32(patch, 40, 40, 0.125, 0.125)
    assert a.shape == b.shape
    assert np.abs(a - b).max() < 1e-6
    maps = rng.normal(size=(46, 82, 26)).astype(np.float32)
    a = cv2.resize(maps, (656, 368), interpolation=cv2.INTER_CUBIC)
    b = orc.bicubic_resize_f32(maps, 368, 656, 46 / 368, 82 / 656)
    assert np.abs(a - b).max() < 2e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_oracle_matches_live_reference():
    warnings.simplefilter("ignore")
    sys.path.insert(0, "/root/reference")
    try:
        from src import pose_decode as ref
    finally:
        sys.path.remove("/root/reference")
    import cv2
    param = {"thre1": 0.1, "thre2": 0.0}
    for seed, persons, H, W in [(101, 4, 368, 368), (102, 12, 368, 656), (103, 2, 256, 256),
                                (104, 6, 368, 368), (105, 0, 368, 368)]:
        heat, paf, _ = synth_scene(seed, persons, H, W)
        img = np.zeros((H, W, 3), dtype=np.uint8)
        per_type = ref.NMS(param, heat, H / heat.shape[0])
        joint_list = np.array([tuple(pk) + (jt,) for jt, pks in enumerate(per_type) for pk in pks])
        paf_up = cv2.resize(paf, (W, H), interpolation=cv2.INTER_CUBIC)
        limbs = ref.find_connected_joints(param, paf_up, per_type)
        persons_arr = ref.group_limbs_of_same_person(limbs, joint_list)
        _, _, _, joints = ref.decode_pose(img, param, heat, paf)
        refd = {"joint_list": joint_list, "limbs": [np.asarray(l).reshape(-1, 5) for l in limbs],
                "persons": persons_arr, "joints": joints}
        got = orc.decode_detail((H, W), param, heat, paf)
        compare_decode(got, refd)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_oracle_stage_variants_match_live_reference():
    """find_peaks (3x3 cross filter), find_peaks_v2 and NMS(bool_refine_center=False) of the reference module
    against the oracle's restatements -- the alternate decode modes of SURVEY.md 8(f)."""
    import warnings
    from lightweight_openpose_b200 import synth
    sys.path.insert(0, "/root/reference")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from src import pose_decode as ref
    finally:
        sys.path.remove("/root/reference")
    param = {"thre1": 0.1, "thre2": 0.0}
    for seed, persons, hw in ((700, 4, (368, 368)), (701, 9, (368, 656)), (702, 1, (256, 256))):
        heat, _, _ = synth.synth_scene(seed, persons, hw[0], hw[1], noise=0.08)
        for j in (0, 5, 13):
            assert np.array_equal(np.asarray(ref.find_peaks(param, heat[:, :, j])).reshape(-1, 2),
                                  orc.cross_peaks(heat[:, :, j], 0.1))
            assert np.array_equal(np.asarray(ref.find_peaks_v2(param, heat[:, :, j])).reshape(-1, 2),
                                  orc.greedy_radius_peaks(heat[:, :, j], 0.1))
        want = ref.NMS(param, heat, 8.0, bool_refine_center=False)
        got = orc.refine_peaks(heat, 0.1, 8.0, refine=False)
        for a, b in zip(got, want):
            assert np.array_equal(a, np.asarray(b).reshape(-1, 5))


def test_restatement_is_bit_exact_vs_reference_without_ipp(decode_golden_noipp):
    """OpenCV's own float32 bicubic (cv2.ipp.setUseIPP(False)) restated: every output of the reference module,
    float scores included, reproduced exactly by the pure-numpy path of the oracle."""
    scenes, param = decode_golden_noipp
    for name, sc in scenes.items():
        got = orc.decode_detail(sc["hw"], param, sc["heat"], sc["paf"], use_cv2=False)
        compare_decode(got, sc, float_tol=0.0, f64_tol=1e-14)
