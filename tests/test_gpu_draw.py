"""The CUDA canvas rasteriser (``lwop_draw_poses``; reference src/pose_decode.py:431-436) through the C ABI against
canvases drawn by the unmodified reference (tests/golden/canvas_golden.npz) and against the CPU oracle on random
tables.  Byte work: exact."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def base_frame(H, W):
    y, x, c = np.meshgrid(np.arange(H), np.arange(W), np.arange(3), indexing="ij")
    return ((x * 3 + y * 5 + c * 70) % 251).astype(np.uint8)


@pytest.fixture(scope="module")
def goldens():
    return (np.load(os.path.join(GOLDEN, "canvas_golden.npz"), allow_pickle=True),
            np.load(os.path.join(GOLDEN, "decode_golden_noipp.npz"), allow_pickle=True))


def golden_tables(name, goldens):
    cg, dg = goldens
    src = cg if f"{name}/persons" in cg.files else dg
    return tuple(int(v) for v in src[f"{name}/hw"]), src[f"{name}/joint_list"], src[f"{name}/persons"]


def test_device_canvas_equals_reference_golden(goldens):
    from lightweight_openpose_b200.src import pose_decode
    for name in goldens[0]["names"]:
        name = str(name)
        (H, W), jl, persons = golden_tables(name, goldens)
        canvas = pose_decode.plot_pose_device(base_frame(H, W), jl, persons)
        want = goldens[0][f"{name}/canvas"]
        assert canvas.dtype == np.uint8 and canvas.shape == want.shape
        assert np.array_equal(canvas, want), (name, int((canvas != want).any(-1).sum()))


def test_plot_pose_env_switch_selects_the_device_rasteriser(goldens, monkeypatch):
    from lightweight_openpose_b200.src import pose_decode
    (H, W), jl, persons = golden_tables("eight_368", goldens)
    monkeypatch.setenv("LWOP_DRAW", "gpu")
    e = pose_decode._decode_engine(16, 16, 1)
    before = e.launch_count()
    canvas, joints = pose_decode.plot_pose(base_frame(H, W), jl, persons)
    assert e.launch_count() > before                      # the CUDA kernels ran, not cv2
    assert np.array_equal(canvas, goldens[0]["eight_368/canvas"])
    assert np.array_equal(joints.reshape(-1, 42), goldens[1]["eight_368/joints"])


def test_device_canvas_equals_oracle_on_random_tables_batched():
    """A batch of frames with random person tables (joints anywhere incl. borders, missing joints, shared joints,
    odd frame size so the scalar tail of the resolve pass runs), drawn in one call; in place as well."""
    import torch
    from lightweight_openpose_b200 import _lib, engine
    from lightweight_openpose_b200.src import pose_decode
    from oracle import draw_oracle
    rng = np.random.default_rng(5)
    B, H, W = 5, 61, 87                                    # 61 * 87 * 5 pixels: not a multiple of 4
    dev = torch.device("cuda", 0)
    counts = np.zeros((B, _lib.COUNTS), dtype=np.int32)
    tj = np.zeros((B, _lib.MAX_JOINTS, 6))
    tp = np.zeros((B, _lib.MAX_PERSONS, 16))
    frames = rng.integers(0, 255, (B, H, W, 3), dtype=np.uint8)
    want = []
    for b in range(B):
        nj, npers = int(rng.integers(1, 60)), int(rng.integers(0, 40)) if b else 0     # frame 0: nobody
        tj[b, :nj, 0], tj[b, :nj, 1] = rng.integers(0, W, nj), rng.integers(0, H, nj)
        tj[b, :nj, 5] = rng.integers(0, 14, nj)
        ids = rng.integers(-1, nj, (npers, 14)).astype(np.float64)
        ids[rng.random((npers, 14)) < 0.3] = -1
        tp[b, :npers, :14] = ids
        counts[b, 0], counts[b, 1] = nj, npers
        want.append(draw_oracle.plot_pose_canvas(frames[b], tj[b, :nj], tp[b, :npers],
                                                 pose_decode.joint_to_limb_heatmap_relationship, pose_decode.colors))
    e = engine.Engine(16, 16, max_batch=1, device=0, use_graph=False)
    tables = {"counts": torch.from_numpy(counts).to(dev), "joints": torch.from_numpy(tj).to(dev),
              "persons": torch.from_numpy(tp).to(dev)}
    fd = torch.from_numpy(frames).to(dev)
    for rep in range(2):                                   # the second call finds the priority map reset by the first
        got = e.draw_poses(fd, tables).cpu().numpy()
        for b in range(B):
            assert np.array_equal(got[b], want[b]), (rep, b, int((got[b] != want[b]).any(-1).sum()))
    e.draw_poses(fd, tables, out=fd)                       # in place
    assert np.array_equal(fd.cpu().numpy(), np.stack(want))
    # an unaligned frame pointer takes the scalar resolve kernel
    flat = torch.empty(B * H * W * 3 + 1, dtype=torch.uint8, device=dev)
    off = flat[1:].view(B, H, W, 3)
    off.copy_(torch.from_numpy(frames).to(dev))
    assert np.array_equal(e.draw_poses(off, tables).cpu().numpy(), np.stack(want))
    e.close()


def test_forward_decode_then_draw_on_device_matches_host_drawing():
    """Engine path: decode tables stay on the device and feed the rasteriser; the canvas equals what the host cv2 path
    of plot_pose draws from the fetched tables."""
    import torch
    from lightweight_openpose_b200 import engine, synth
    from lightweight_openpose_b200.src import pose_decode
    H = W = 368
    scenes = [synth.synth_scene(s, n, H, W) for s, n in ((3, 3), (8, 8), (7, 0))]
    heat = torch.from_numpy(np.stack([s[0] for s in scenes])).cuda()
    paf = torch.from_numpy(np.stack([s[1] for s in scenes])).cuda()
    e = engine.Engine(H, W, max_batch=3, device=0, use_graph=False)
    res = e.decode(heat, paf, (H, W), 0.1, 0.0)
    frames = np.stack([base_frame(H, W)] * 3)
    canv = e.draw_poses(torch.from_numpy(frames).cuda()).cpu().numpy()
    for b in range(3):
        want, _ = pose_decode.plot_pose(frames[b], res[b].joint_list, res[b].person_to_joint_assoc)
        assert np.array_equal(canv[b], want), b
    e.close()
