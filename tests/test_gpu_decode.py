"""Decode-chain parity on the GPU: against golden vectors produced by the
unmodified reference decoder and against the CPU oracle on fresh scenes.
Integer outputs (peak x/y, ids, limb pairs, grouping, keypoint table) exact,
float scores within 1e-5."""
import numpy as np
import pytest
import torch

from lightweight_openpose_b200 import engine, synth
from lightweight_openpose_b200.src import pose_decode
from oracle import decode_oracle as orc
from test_oracle_decode import compare_decode

pytestmark = pytest.mark.gpu
PARAM = {"thre1": 0.1, "thre2": 0.0}


def _as_dict(res):
    return {"joint_list": res.joint_list, "limbs": res.limbs, "persons": res.person_to_joint_assoc,
            "joints": res.joints}


def test_decode_matches_reference_golden(decode_golden):
    scenes, param = decode_golden
    for name, sc in scenes.items():
        res = pose_decode.decode_pose_batch(sc["hw"], param, sc["heat"][None], sc["paf"][None])[0]
        compare_decode(_as_dict(res), sc)


def test_decode_bit_exact_vs_reference_without_ipp(decode_golden_noipp):
    """Against the reference module run on OpenCV's own float32 arithmetic (IPP off): everything exact,
    float scores included."""
    scenes, param = decode_golden_noipp
    for name, sc in scenes.items():
        res = pose_decode.decode_pose_batch(sc["hw"], param, sc["heat"][None], sc["paf"][None])[0]
        compare_decode(_as_dict(res), sc, float_tol=0.0, f64_tol=1e-14)


def test_decode_pose_signature_and_types(decode_golden):
    scenes, param = decode_golden
    sc = scenes["three_368"]
    img = np.zeros(sc["hw"] + (3,), np.uint8)
    canvas, jl, persons, joints = pose_decode.decode_pose(img, param, sc["heat"], sc["paf"])
    assert canvas.shape == img.shape and canvas.dtype == np.uint8 and canvas.any()
    assert jl.dtype == np.float64 and jl.shape == (42, 6)
    assert persons.dtype == np.float64 and persons.shape == (3, 16)
    assert joints.dtype == np.float32 and joints.shape == (3, 42)
    ocanvas, _, _, _ = orc.decode(img, param, sc["heat"], sc["paf"])
    assert np.array_equal(canvas, ocanvas)
    sc = scenes["empty_368"]
    canvas, jl, persons, joints = pose_decode.decode_pose(np.zeros(sc["hw"] + (3,), np.uint8), param,
                                                          sc["heat"], sc["paf"])
    assert jl.shape == (0,) and persons.shape == (0,) and joints.shape == (0, 42)


def test_decode_batched_matches_oracle_on_fresh_scenes():
    specs = [(200 + i, 1 + (i * 3) % 12) for i in range(24)]
    heats, pafs = [], []
    for seed, persons in specs:
        h, p, _ = synth.synth_scene(seed, persons, 368, 368)
        heats.append(h)
        pafs.append(p)
    res = pose_decode.decode_pose_batch((368, 368), PARAM, np.stack(heats), np.stack(pafs))
    for (seed, persons), h, p, r in zip(specs, heats, pafs, res):
        ref = orc.decode_detail((368, 368), PARAM, h, p)
        compare_decode(_as_dict(r), ref)
        assert r.person_to_joint_assoc.shape[0] == persons


def test_decode_wide_frames_dense_scene():
    """config[3]: 656x368 frames with 20-person scenes."""
    heats, pafs = zip(*[synth.synth_scene(300 + i, 20, 368, 656)[:2] for i in range(4)])
    res = pose_decode.decode_pose_batch((368, 656), PARAM, np.stack(heats), np.stack(pafs))
    for h, p, r in zip(heats, pafs, res):
        ref = orc.decode_detail((368, 656), PARAM, h, p)
        compare_decode(_as_dict(r), ref)
        assert r.person_to_joint_assoc.shape[0] == 20


def test_decode_is_deterministic_and_batch_invariant():
    heats, pafs = zip(*[synth.synth_scene(400 + i, 5, 368, 368, noise=0.12)[:2] for i in range(6)])
    a = pose_decode.decode_pose_batch((368, 368), PARAM, np.stack(heats), np.stack(pafs))
    b = pose_decode.decode_pose_batch((368, 368), PARAM, np.stack(heats[2:4]), np.stack(pafs[2:4]))
    for r, s in zip(a[2:4], b):
        assert np.array_equal(r.joint_list, s.joint_list)
        assert np.array_equal(r.person_to_joint_assoc, s.person_to_joint_assoc)
        assert np.array_equal(r.joints, s.joints)


def test_nms_entry_point(decode_golden):
    scenes, param = decode_golden
    sc = scenes["eight_368"]
    per_type = pose_decode.NMS(param, sc["heat"], 8.0)
    ref = orc.refine_peaks(sc["heat"], param["thre1"], 8.0)
    assert len(per_type) == 14
    for a, b in zip(per_type, ref):
        assert a.shape == b.shape
        assert np.array_equal(a[:, [0, 1, 3, 4]], b[:, [0, 1, 3, 4]])
        assert np.abs(a[:, 2] - b[:, 2]).max() <= 1e-5


def test_forward_then_decode_end_to_end():
    """The whole path through the host-buffer C entry (lwop_infer_host) against
    oracle forward + oracle decode fed with the GPU's own fp32-check maps."""
    from lightweight_openpose_b200 import checkpoint
    v, src = checkpoint.load_or_synthesize()
    eng = engine.Engine(368, 368, max_batch=4, device=0)
    eng.load_variables(v, src)
    x = synth.synth_batch(range(4), 368, 368)
    res, heat, paf = eng.infer_host(x, mode="fp32", want_maps=True)
    for i in range(4):
        ref = orc.decode_detail((368, 368), PARAM, heat[i], paf[i])
        compare_decode(_as_dict(res[i]), ref)
    res16 = eng.infer_host(x, mode="bf16")
    assert len(res16) == 4
    u8 = np.stack([synth.synth_image(s, 368, 368) for s in range(4)])
    res8 = eng.infer_host(u8, mode="bf16")
    for a, b in zip(res16, res8):
        assert np.array_equal(a.joint_list, b.joint_list)


def _lists_equal(got, want, tol=1e-5):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        a, b = np.asarray(a, dtype=np.float64).reshape(-1, 5), np.asarray(b, dtype=np.float64).reshape(-1, 5)
        assert a.shape == b.shape
        if a.size:
            assert np.array_equal(a[:, [0, 1, 3, 4]], b[:, [0, 1, 3, 4]])      # integer fields exact
            assert np.abs(a[:, 2] - b[:, 2]).max() <= tol


@pytest.mark.parametrize("spec", [(500, 3, 368, 368), (501, 12, 368, 656), (502, 1, 256, 256), (503, 0, 368, 368)])
def test_decode_stage_functions_match_oracle(spec):
    """The separately importable stages of the reference module (NMS, find_peaks, find_peaks_v2,
    find_connected_joints, group_limbs_of_same_person), each through its own C-ABI entry."""
    seed, persons, H, W = spec
    if persons:
        heat, paf, _ = synth.synth_scene(seed, persons, H, W, noise=0.05)
    else:
        heat = np.zeros((H // 8, W // 8, 14), np.float32)
        paf = np.zeros((H // 8, W // 8, 26), np.float32)
    factor = H / heat.shape[0]
    per_type = pose_decode.NMS(PARAM, heat, factor)
    want = orc.refine_peaks(heat, PARAM["thre1"], factor)
    _lists_equal(per_type, want)
    _lists_equal(pose_decode.NMS(PARAM, heat, factor, bool_refine_center=False),
                 orc.refine_peaks(heat, PARAM["thre1"], factor, refine=False), tol=0.0)
    for j in (0, 7):
        assert np.array_equal(pose_decode.find_peaks_v2(PARAM, heat[:, :, j]).reshape(-1, 2),
                              orc.greedy_radius_peaks(heat[:, :, j], PARAM["thre1"]))
        assert np.array_equal(pose_decode.find_peaks(PARAM, heat[:, :, j]).reshape(-1, 2),
                              orc.cross_peaks(heat[:, :, j], PARAM["thre1"]))
    paf_up = orc._resize_full(paf, H, W, True)                       # what decode_pose hands over (:487)
    limbs = pose_decode.find_connected_joints(PARAM, paf_up, want)
    ref_limbs = orc.score_limbs(paf_up, want, PARAM["thre2"])
    for t, (a, b) in enumerate(zip(limbs, ref_limbs)):
        empty = len(want[pose_decode.joint_to_limb_heatmap_relationship[t][0]]) == 0 or \
            len(want[pose_decode.joint_to_limb_heatmap_relationship[t][1]]) == 0
        if empty:
            assert isinstance(a, list) and a == []                  # reference returns [] here (:224-228)
        else:
            assert np.asarray(a).shape == b.shape
            if b.size:
                assert np.array_equal(np.asarray(a)[:, [0, 1, 3, 4]], b[:, [0, 1, 3, 4]])
                assert np.abs(np.asarray(a)[:, 2] - b[:, 2]).max() <= 1e-12     # direct samples: float64 exact to rounding
    joint_list = np.array([tuple(pk) + (jt,) for jt, pks in enumerate(want) for pk in pks])
    got_p = pose_decode.group_limbs_of_same_person(ref_limbs, joint_list)
    want_p = orc.group_persons(ref_limbs, joint_list)
    assert got_p.shape == want_p.shape
    if want_p.size:
        assert np.array_equal(got_p[:, :14], want_p[:, :14]) and np.array_equal(got_p[:, 15], want_p[:, 15])
        assert np.abs(got_p[:, 14] - want_p[:, 14]).max() <= 1e-9
        assert got_p.shape[0] == persons


def test_nms_gaussian_patch_filter(decode_golden_noipp):
    """NMS(bool_gaussian_filt=True) (pose_decode.py:163-164) on the GPU: bit-exact against the reference module's own
    output (OpenCV's own bicubic + scipy's gaussian_filter) and against the oracle on fresh scenes."""
    scenes, param = decode_golden_noipp
    for name, sc in scenes.items():
        factor = sc["hw"][0] / sc["heat"].shape[0]
        got = pose_decode.NMS(param, sc["heat"], factor, True, True)
        rows = np.array([tuple(pk) + (jt,) for jt, pks in enumerate(got) for pk in pks], dtype=np.float64).reshape(-1, 6)
        assert rows.shape == sc["nms_gauss"].shape and np.array_equal(rows, sc["nms_gauss"]), name
    for seed, persons, H, W in ((510, 5, 368, 368), (511, 14, 368, 656)):
        heat, _, _ = synth.synth_scene(seed, persons, H, W, noise=0.05)
        got = pose_decode.NMS(PARAM, heat, 8.0, bool_gaussian_filt=True)
        want = orc.refine_peaks(heat, PARAM["thre1"], 8.0, use_cv2=False, gaussian=True)
        _lists_equal(got, want, tol=0.0)
        # without centre refinement the flag has no effect, as in the reference
        _lists_equal(pose_decode.NMS(PARAM, heat, 8.0, False, True), orc.refine_peaks(heat, PARAM["thre1"], 8.0, refine=False), tol=0.0)


@pytest.mark.parametrize("spec", [(1, 6, 0.9), (2, 20, 0.7), (3, 40, 0.5), (4, 60, 0.35), (5, 100, 0.25), (6, 3, 1.0)])
def test_grouping_on_adversarial_tables(spec):
    """The grouping kernel on random connection tables (repeated destinations, merges, >= 3 matches, list.pop): up
    to ~300 working persons, i.e. beyond the 128 rows the kernel keeps in shared memory (global spill path)."""
    seed, n, dens = spec
    limbs, joint_list = synth.random_limb_tables(seed, n, dens)
    want = orc.group_persons(limbs, joint_list)
    got = pose_decode.group_limbs_of_same_person(limbs, joint_list)
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.size:
        assert np.array_equal(got[:, :14], want[:, :14]) and np.array_equal(got[:, 15], want[:, 15])
        assert np.abs(got[:, 14] - want[:, 14]).max() <= 1e-9


def test_capacity_limits_fail_loudly():
    """The device tables have fixed capacities the reference does not have (256 peaks per joint channel, 256 persons):
    exceeding one is an error, never a silently truncated result."""
    from lightweight_openpose_b200 import _lib
    heat = np.zeros((108, 108, 14), np.float32)
    heat[3::6, 3::6, 0] = 0.9                       # 18 x 18 = 324 isolated peaks, 6 px apart (> the radius-5 suppression)
    assert len(orc.greedy_radius_peaks(heat[:, :, 0], 0.1)) == 324
    with pytest.raises(_lib.LwopError):
        pose_decode.NMS(PARAM, heat, 8.0)
    with pytest.raises(_lib.LwopError):
        pose_decode.decode_pose(np.zeros((864, 864, 3), np.uint8), PARAM, heat, np.zeros((108, 108, 26), np.float32))
    heat = np.zeros((96, 96, 14), np.float32)
    heat[3::6, 3::6, 0] = 0.9                       # 16 x 16 = 256 peaks: exactly the capacity, no error
    got = pose_decode.NMS(PARAM, heat, 8.0)
    assert len(got[0]) == 256 and all(len(g) == 0 for g in got[1:])
    eng = engine.Engine(64, 64, max_batch=2, device=0)
    with pytest.raises(ValueError):
        eng.forward(torch.zeros((3, 64, 64, 3), device="cuda"))


def test_decode_dense_46x82_frame_matches_reference_golden(decode_extra_golden):
    """BASELINE configs[3] frame size with every heat-map pixel above thre1: ~104 peaks per channel (more than the 128-row
    tables of round 1 could take in a denser packing), 1288 connections, 141 persons -- the reference returns a result
    and the CUDA chain returns the same one, bit for bit (register-resident grouping hands over to the general kernel)."""
    dense, _, param = decode_extra_golden
    res = pose_decode.decode_pose_batch(dense["hw"], param, dense["heat"][None], dense["paf"][None])[0]
    compare_decode(_as_dict(res), dense, float_tol=0.0, f64_tol=1e-13)
    # batched next to light frames: the fast / general grouping split is per image
    h3, p3, _ = synth.synth_scene(3, 3, 368, 656)
    heat = np.stack([h3, dense["heat"], h3])
    paf = np.stack([p3, dense["paf"], p3])
    out = pose_decode.decode_pose_batch(dense["hw"], param, heat, paf)
    compare_decode(_as_dict(out[1]), dense, float_tol=0.0, f64_tol=1e-13)
    ref3 = orc.decode_detail((368, 656), param, h3, p3, use_cv2=False)
    for k in (0, 2):
        compare_decode(_as_dict(out[k]), ref3, float_tol=0.0, f64_tol=1e-13)


def test_find_connected_joints_keyword_arguments(decode_extra_golden):
    """num_intermed_pts / max_dist_thresh of the reference signature (:188), against the reference's own output."""
    _, params, param = decode_extra_golden
    jl = params["joint_list"]
    per_type = [jl[jl[:, 5] == j][:, :5] for j in range(14)]
    paf_up = orc._resize_full(params["paf"], 368, 368, False)
    for (n, thr), want in zip(params["settings"], params["limbs"]):
        got = pose_decode.find_connected_joints(param, paf_up, per_type, num_intermed_pts=n, max_dist_thresh=thr)
        for t in range(13):
            a, b = np.asarray(got[t], dtype=np.float64).reshape(-1, 5), want[t].reshape(-1, 5)
            assert a.shape == b.shape, (n, thr, t)
            if a.size:
                assert np.array_equal(a[:, [0, 1, 3, 4]], b[:, [0, 1, 3, 4]]), (n, thr, t)
                assert np.abs(a[:, 2] - b[:, 2]).max() <= 1e-12
    with pytest.raises(ValueError):
        pose_decode.find_connected_joints(param, paf_up, per_type, num_intermed_pts=65)


def test_peak_search_on_tied_plateaus():
    """Exact ties: the reference orders candidates with an unstable argsort, so which pixel of a flat plateau survives
    is implementation-defined there.  Asserted here: what does NOT depend on the tie order -- kept peaks
    are pairwise farther than 5 px apart, every candidate is within 5 px of a kept peak of equal or higher value,
    strict maxima are always kept -- and that the CUDA kernel and the oracle apply the same documented rule (larger
    row-major index first)."""
    rng = np.random.default_rng(77)
    heat = np.zeros((46, 46, 14), np.float32)
    heat[:, :, 0] = 0.5                                            # one plateau over the whole map
    heat[10:20, 10:30, 1] = 0.7                                    # plateau + strict maxima
    heat[5, 5, 1], heat[40, 40, 1] = 0.9, 0.95
    heat[:, :, 2] = np.round(rng.random((46, 46)) * 4) / 4         # quantised map: many ties
    for ch in range(3):
        got = pose_decode.find_peaks_v2(PARAM, heat[:, :, ch]).reshape(-1, 2)
        want = orc.greedy_radius_peaks(heat[:, :, ch], 0.1)
        assert np.array_equal(got, want)
        xy = got.astype(np.int64)
        d2 = ((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1)
        assert (d2[~np.eye(len(xy), dtype=bool)] > 25).all()
        vals = heat[xy[:, 1], xy[:, 0], ch]
        assert (np.diff(vals) <= 0).all()                          # best first
        ys, xs = np.nonzero(heat[:, :, ch] > 0.1)
        for y, x in zip(ys, xs):
            near = ((xy[:, 0] - x) ** 2 + (xy[:, 1] - y) ** 2) <= 25
            assert near.any() and vals[near].max() >= heat[y, x, ch]
    got1 = pose_decode.find_peaks_v2(PARAM, heat[:, :, 1]).reshape(-1, 2).tolist()
    assert [40, 40] in got1 and [5, 5] in got1


def test_forward_decode_single_graph_equals_separate_calls():
    """lwop_forward_decode (forward + decode chain replayed from one CUDA graph) == lwop_forward then lwop_decode."""
    from lightweight_openpose_b200 import checkpoint
    eng = engine.Engine(368, 368, max_batch=4, device=0)
    variables, src = checkpoint.load_or_synthesize()
    eng.load_variables(variables, src)
    x = torch.from_numpy(np.stack([synth.synth_image(s, 368, 368) for s in range(4)])).cuda()
    heat, paf = eng.forward(x)
    want = eng.decode(heat.clone(), paf.clone(), (368, 368))
    for _ in range(3):                                             # eager pass, capture, replay
        eng.forward_decode(x)
        got = eng.fetch(4)
        assert np.array_equal(got.counts, want.counts)
        assert np.array_equal(got.joints, want.joints) and np.array_equal(got.persons, want.persons)
        assert np.array_equal(got.limbs, want.limbs) and np.array_equal(got.keypoints, want.keypoints)
    assert int(want.counts[:, 0].sum()) > 0


def test_decode_parity_1000_frames():
    """BASELINE configs[4] in the driver-run suite: the standalone decode on 1000 generator scenes (1-20 persons, seed =
    frame index) against the oracle -- integer fields and float32 scores exact, float64 scores <= 1e-12.  The 10 000
    frame run of the same loop is tools/decode_parity_10k.py (profiles/)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import decode_parity_10k
    line = decode_parity_10k.run(1000, time_limit=600.0)
    assert line["frames_checked"] == 1000
    assert line["frames_with_any_mismatch"] == 0, line
    assert line["max_abs_joint_score_diff"] == 0.0
    assert line["persons"] > 5000
