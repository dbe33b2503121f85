"""Property / fuzz level of SURVEY.md section 4 with hypothesis: generated inputs instead of a fixed seed list.

CPU: the oracles against what they restate -- cv2's drawing and resizing, and (when /root/reference exists, i.e. in the
build container) the unmodified reference decoder.  GPU: the CUDA decode chain, rasteriser and pre-processing kernel
against the oracles on generated scenes.  Integer and byte outputs exact, float scores <= 1e-5."""
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

from lightweight_openpose_b200.synth import synth_scene  # noqa: E402
from oracle import decode_oracle, draw_oracle, preprocess_oracle  # noqa: E402
from test_oracle_decode import compare_decode  # noqa: E402

PARAM = {"thre1": 0.1, "thre2": 0.0}
COMMON = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
HAVE_REF = os.path.isdir("/root/reference/src")


def _reference_module():
    warnings.simplefilter("ignore")
    sys.path.insert(0, "/root/reference")
    try:
        from src import pose_decode as ref
    finally:
        sys.path.remove("/root/reference")
    return ref


# ---------------------------------------------------------------- CPU: oracles against what they restate
@settings(max_examples=300, **COMMON)
@given(h=st.integers(4, 90), w=st.integers(4, 90), x0=st.integers(-15, 105), y0=st.integers(-15, 105),
       x1=st.integers(-15, 105), y1=st.integers(-15, 105), col=st.tuples(*[st.integers(0, 255)] * 3))
def test_draw_oracle_equals_cv2(h, w, x0, y0, x1, y1, col):
    cv2 = pytest.importorskip("cv2")
    img = ((np.arange(h * w * 3).reshape(h, w, 3) * 7) % 251).astype(np.uint8)
    a, b = img.copy(), img.copy()
    cv2.circle(a, (x0, y0), 3, (255, 255, 255), thickness=-1)
    cv2.circle(a, (x1, y1), 3, (255, 255, 255), thickness=-1)
    cv2.line(a, (x0, y0), (x1, y1), color=col, thickness=2)
    draw_oracle.draw_limb(b, (x0, y0), (x1, y1), col)
    assert np.array_equal(a, b)


@settings(max_examples=60, **COMMON)
@given(sh=st.integers(1, 70), sw=st.integers(1, 70), dh=st.integers(1, 70), dw=st.integers(1, 70), seed=st.integers(0, 2 ** 31))
def test_preprocess_oracle_equals_cv2(sh, sw, dh, dw, seed):
    cv2 = pytest.importorskip("cv2")
    frame = np.random.default_rng(seed).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    want = cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB), (dw, dh))          # model_test.py:63-64
    assert np.array_equal(preprocess_oracle.preprocess_bgr(frame, dh, dw), want)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
@settings(max_examples=12, **COMMON)
@given(seed=st.integers(0, 10 ** 6), persons=st.integers(0, 9), wide=st.booleans(), noise=st.sampled_from([0.02, 0.05, 0.11]))
def test_decode_oracle_equals_live_reference(seed, persons, wide, noise):
    import cv2
    ref = _reference_module()
    H, W = (368, 656) if wide else (256, 256)
    heat, paf, _ = synth_scene(seed, persons, H, W, noise=noise)
    per_type = ref.NMS(PARAM, heat, H / heat.shape[0])
    joint_list = np.array([tuple(pk) + (jt,) for jt, pks in enumerate(per_type) for pk in pks])
    paf_up = cv2.resize(paf, (W, H), interpolation=cv2.INTER_CUBIC)
    limbs = ref.find_connected_joints(PARAM, paf_up, per_type)
    persons_arr = ref.group_limbs_of_same_person(limbs, joint_list)
    _, _, _, joints = ref.decode_pose(np.zeros((H, W, 3), np.uint8), PARAM, heat, paf)
    refd = {"joint_list": joint_list, "limbs": [np.asarray(l).reshape(-1, 5) for l in limbs], "persons": persons_arr,
            "joints": joints}
    compare_decode(decode_oracle.decode_detail((H, W), PARAM, heat, paf), refd)


# ---------------------------------------------------------------- GPU: kernels against the oracles
@pytest.mark.gpu
@settings(max_examples=25, **COMMON)
@given(seed=st.integers(0, 10 ** 6), persons=st.integers(0, 12), size=st.sampled_from([(368, 368), (368, 656), (256, 256), (136, 200)]),
       noise=st.sampled_from([0.02, 0.05, 0.11]))
def test_gpu_decode_and_canvas_equal_oracle(seed, persons, size, noise):
    import torch
    from lightweight_openpose_b200.src import pose_decode
    H, W = size
    heat, paf, _ = synth_scene(seed, persons, H, W, noise=noise)
    res = pose_decode.decode_pose_batch((H, W), PARAM, heat[None], paf[None])[0]
    want = decode_oracle.decode_detail((H, W), PARAM, heat, paf, use_cv2=False)      # OpenCV's own arithmetic, restated
    got = {"joint_list": res.joint_list, "limbs": res.limbs, "persons": res.person_to_joint_assoc, "joints": res.joints}
    compare_decode(got, want, float_tol=0.0, f64_tol=1e-12)
    frame = np.random.default_rng(seed).integers(0, 256, (H, W, 3), dtype=np.uint8)
    persons_t = np.asarray(res.person_to_joint_assoc, dtype=np.float64).reshape(-1, 16)
    jl = np.asarray(res.joint_list, dtype=np.float64).reshape(-1, 6)
    canvas = pose_decode.plot_pose_device(frame, jl, persons_t)
    assert np.array_equal(canvas, draw_oracle.plot_pose_canvas(frame, jl, persons_t, pose_decode.joint_to_limb_heatmap_relationship,
                                                               pose_decode.colors))
    assert torch.cuda.is_available()


@pytest.mark.gpu
@settings(max_examples=25, **COMMON)
@given(sh=st.integers(1, 300), sw=st.integers(1, 300), dst=st.sampled_from([(256, 256), (368, 368), (64, 48), (17, 33)]),
       seed=st.integers(0, 2 ** 31))
def test_gpu_preprocess_equals_oracle(sh, sw, dst, seed):
    import torch
    from lightweight_openpose_b200 import _lib, engine
    frame = np.random.default_rng(seed).integers(0, 256, (1, sh, sw, 3), dtype=np.uint8)
    e = engine.Engine(max(dst[0], 16), max(dst[1], 16), max_batch=1, device=0, use_graph=False)
    s = torch.from_numpy(frame).cuda()
    d = torch.empty((1, dst[0], dst[1], 3), dtype=torch.uint8, device="cuda")
    _lib.check(e.lib.lwop_preprocess_bgr_u8(e.handle, s.data_ptr(), 1, sh, sw, d.data_ptr(), dst[0], dst[1],
                                            torch.cuda.current_stream().cuda_stream), e.handle)
    got = d.cpu().numpy()[0]
    e.close()
    assert np.array_equal(got, preprocess_oracle.preprocess_bgr(frame[0], dst[0], dst[1]))
