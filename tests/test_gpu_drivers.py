"""Driver-level path: device pre-processing against the cv2-restating oracle and the batched
replacement of the model_test / model_json loops against the same steps done one frame at a time."""
import json

import numpy as np
import pytest
import torch

from lightweight_openpose_b200 import _lib, drivers, engine, synth
from lightweight_openpose_b200.src import pose_decode
from oracle import preprocess_oracle as pre

pytestmark = pytest.mark.gpu


def _frames(n, hw, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        f = synth.synth_image(seed + i, hw[0], hw[1])[:, :, ::-1].copy()          # BGR like cv2.imread
        f[::7, ::5] = rng.integers(0, 256, size=f[::7, ::5].shape, dtype=np.uint8)
        out.append(f)
    return out


def test_device_preprocess_equals_cv2_golden_bytes():
    """The CUDA pre-processing against bytes the installed cv2 produced (tests/golden/preprocess_golden.npz, IPP off;
    up- and down-scaling, both axes): exact."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_golden.npz"))
    k = 0
    while f"frame{k}" in g:
        frame, want = g[f"frame{k}"], g[f"rgb{k}"]
        e = engine.Engine(want.shape[0], want.shape[1], max_batch=1, device=0, use_graph=False)
        s = torch.from_numpy(frame[None]).cuda()
        d = torch.empty((1,) + want.shape, dtype=torch.uint8, device="cuda")
        _lib.check(e.lib.lwop_preprocess_bgr_u8(e.handle, s.data_ptr(), 1, frame.shape[0], frame.shape[1], d.data_ptr(),
                                                want.shape[0], want.shape[1], torch.cuda.current_stream().cuda_stream),
                   e.handle)
        assert np.array_equal(d.cpu().numpy()[0], want), k
        e.close()
        k += 1
    assert k >= 5


@pytest.mark.parametrize("src,dst", [((480, 640), (256, 256)), ((720, 1280), (368, 368)), ((199, 301), (368, 656)),
                                     ((368, 368), (368, 368)), ((100, 100), (368, 368)), ((64, 48), (256, 256))])
def test_device_preprocess_is_bit_exact_vs_oracle(src, dst):
    frames = np.stack(_frames(3, src, 11))
    e = engine.Engine(dst[0], dst[1], max_batch=3, device=0, use_graph=False)
    s = torch.from_numpy(frames).cuda()
    d = torch.empty((3, dst[0], dst[1], 3), dtype=torch.uint8, device="cuda")
    _lib.check(e.lib.lwop_preprocess_bgr_u8(e.handle, s.data_ptr(), 3, src[0], src[1], d.data_ptr(), dst[0], dst[1],
                                            torch.cuda.current_stream().cuda_stream), e.handle)
    got = d.cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], pre.preprocess_bgr(frames[i], dst[0], dst[1]))


def test_run_frames_matches_frame_by_frame_path():
    runner = drivers.PoseRunner((256, 256), max_batch=4)
    frames = _frames(3, (480, 640), 21) + _frames(2, (360, 500), 31)              # two source sizes, 5 frames
    res = runner.run_frames(frames, keep_frames=True)
    assert len(res) == 5
    for f, r in zip(frames, res):
        img_data = pre.preprocess_bgr(f, 256, 256)                                # model_test.py:63-64
        assert np.array_equal(r.img_data, img_data)
        heat, paf = runner.engine.forward(torch.from_numpy(img_data[None]).cuda())
        one = pose_decode.decode_pose_batch((256, 256), runner.params, heat.clone(), paf.clone())[0]
        assert np.array_equal(one.joint_list, r.joint_list)
        assert np.array_equal(one.person_to_joint_assoc, r.person_to_joint_assoc)
        assert np.array_equal(one.joints, r.joints)
        assert r.canvas().shape == (256, 256, 3)
    # canvases rasterised on the device from the device-resident tables = the host cv2 drawing of the fetched tables
    drawn = runner.run_frames(frames, keep_frames=True, draw=True)
    for r in drawn:
        assert r.canvas_data is not None and r.canvas() is r.canvas_data
        host = r.img_data.copy() if not r.person_to_joint_assoc.size else \
            pose_decode.plot_pose(r.img_data, r.joint_list, r.person_to_joint_assoc)[0]
        assert np.array_equal(r.canvas_data, host)


def test_keypoint_annotations_json_shape(tmp_path):
    runner = drivers.PoseRunner((256, 256), max_batch=4)
    frames = {f"img{i}": fr for i, fr in enumerate(_frames(3, (300, 400), 41))}
    labels = [{"image_id": k, "extra": 1} for k in frames]
    preds = drivers.keypoint_annotations(runner, labels, lambda image_id: frames[image_id])
    assert [p["image_id"] for p in preds] == list(frames) and all(p["extra"] == 1 for p in preds)
    res = runner.run_frames(list(frames.values()))
    for p, r in zip(preds, res):
        kps = p["keypoint_annotations"]
        assert len(kps) == r.joints.shape[0]
        for h, row in enumerate(r.joints, start=1):
            want = np.array(row, dtype=np.float32, copy=True)
            want[0::3] *= 400 / 256
            want[1::3] *= 300 / 256
            assert kps["human" + str(h)] == want.tolist() and len(kps["human" + str(h)]) == 42
    path = tmp_path / "pred.json"
    drivers.write_predictions(preds, str(path))
    assert json.load(open(path))[0]["image_id"] == "img0"
