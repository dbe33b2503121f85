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
from lightweight_openpose_b200 import synth
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


def test_gaussian_filter_restatement_equals_scipy():
    """NMS(bool_gaussian_filt=True) smooths the up-sampled patch with scipy.ndimage.gaussian_filter(sigma=3)
    (pose_decode.py:8,163-164; third-party, scipy as installed): the numpy restatement is bit-identical, also for
    lines shorter than the 12-sample kernel radius (multiple reflections)."""
    scipy_ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(11)
    for shape in [(40, 40), (24, 40), (40, 24), (24, 24), (8, 8), (5, 17), (48, 48)]:
        for _ in range(5):
            a = rng.normal(0.3, 0.2, size=shape).astype(np.float32)
            want = scipy_ndimage.gaussian_filter(a, sigma=3)
            got = orc.gaussian_filter_f32(a)
            assert want.dtype == got.dtype == np.float32 and np.array_equal(want, got)


def test_gaussian_weights_match_the_kernel_constants():
    """The CUDA kernel carries the 13 float64 weights as hex literals; they must be numpy's."""
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "lightweight_openpose_b200", "csrc", "lwop_decode.cu")).read()
    block = src[src.index("c_gauss3[kGaussRadius + 1] = {"):]
    lits = re.findall(r"0x1\.[0-9a-f]+p-\d+", block[:block.index("};")])
    w = orc.gaussian_weights()
    assert len(lits) == 13 and [float.fromhex(v) for v in lits] == [float(v) for v in w[:13]]
    assert np.array_equal(w, w[::-1])


def test_oracle_gaussian_nms_matches_reference_golden(decode_golden_noipp, decode_golden):
    """refine_peaks(gaussian=True) against the reference module's NMS(..., True, True): exact on OpenCV's own
    bicubic (IPP off); on an IPP build integer fields are exact and scores agree to 1e-6."""
    for (scenes, param), use_cv2, tol in ((decode_golden_noipp, False, 0.0), (decode_golden, True, 1e-6)):
        for name, sc in scenes.items():
            H = sc["hw"][0]
            got = orc.refine_peaks(sc["heat"], param["thre1"], H / sc["heat"].shape[0], use_cv2=use_cv2, gaussian=True)
            rows = np.array([tuple(pk) + (jt,) for jt, pks in enumerate(got) for pk in pks], dtype=np.float64).reshape(-1, 6)
            want = sc["nms_gauss"]
            assert rows.shape == want.shape, name
            if rows.size:
                if tol == 0.0:
                    assert np.array_equal(rows, want), name
                else:
                    # a last-ulp difference of IPP's patch resize can move the argmax of the smoothed (flat-topped) patch by a pixel
                    assert np.abs(rows[:, :2] - want[:, :2]).max() <= 1 and np.array_equal(rows[:, 3:], want[:, 3:]), name
                    assert np.abs(rows[:, 2] - want[:, 2]).max() <= tol, name


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_grouping_restatement_on_adversarial_tables_matches_live_reference():
    """group_limbs_of_same_person (:304-396) on random connection tables: repeated destinations, two-person merges,
    >= 3 matches (-> new person), long person lists (list.pop far down the list) -- cases natural scenes rarely hit."""
    import warnings
    sys.path.insert(0, "/root/reference")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from src import pose_decode as ref
    finally:
        sys.path.remove("/root/reference")
    for seed, n, dens in ((1, 6, 0.9), (2, 20, 0.7), (3, 40, 0.5), (4, 60, 0.35), (5, 100, 0.25), (6, 3, 1.0)):
        limbs, joint_list = synth.random_limb_tables(seed, n, dens)
        want = np.asarray(ref.group_limbs_of_same_person([l if len(l) else [] for l in limbs], joint_list))
        got = orc.group_persons(limbs, joint_list)
        assert got.shape == want.shape, (seed, got.shape, want.shape)
        if want.size:
            assert np.array_equal(got, want), seed


def _per_type_from_joint_list(jl):
    jl = np.asarray(jl, dtype=np.float64).reshape(-1, 6)
    return [jl[jl[:, 5] == j][:, :5] for j in range(14)]


def test_oracle_dense_frame_matches_reference_golden(decode_extra_golden):
    """46x82 maps with every pixel above thre1 (BASELINE configs[3] frame size): ~104 peaks per channel, 1288
    connections, 141 persons in the reference's own output -- all reproduced exactly."""
    dense, _, param = decode_extra_golden
    got = orc.decode_detail(dense["hw"], param, dense["heat"], dense["paf"], use_cv2=False)
    compare_decode(got, dense, float_tol=0.0, f64_tol=1e-13)
    assert max(np.bincount(dense["joint_list"][:, 5].astype(int))) > 100


def test_oracle_limb_keyword_arguments_match_reference_golden(decode_extra_golden):
    """find_connected_joints(num_intermed_pts, max_dist_thresh) (reference :188) for nine settings."""
    _, params, param = decode_extra_golden
    per_type = _per_type_from_joint_list(params["joint_list"])
    paf_up = orc._resize_full(params["paf"], 368, 368, False)
    differs = 0
    for (n, thr), want in zip(params["settings"], params["limbs"]):
        got = orc.score_limbs(paf_up, per_type, param["thre2"], num_intermed_pts=n, max_dist_thresh=thr)
        for t in range(13):
            a, b = got[t].reshape(-1, 5), want[t].reshape(-1, 5)
            assert a.shape == b.shape, (n, thr, t)
            if a.size:
                assert np.array_equal(a[:, [0, 1, 3, 4]], b[:, [0, 1, 3, 4]]), (n, thr, t)
                assert np.abs(a[:, 2] - b[:, 2]).max() <= 1e-13
        base = params["limbs"][params["settings"].index((9, 1.0))]
        differs += any(not np.array_equal(want[t], base[t]) for t in range(13))
    assert differs >= 5          # the settings do change the result
