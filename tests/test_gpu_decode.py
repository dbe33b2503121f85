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
