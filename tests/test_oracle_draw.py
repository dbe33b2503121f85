"""oracle/draw_oracle.py (CPU restatement of the drawing half of plot_pose, reference src/pose_decode.py:431-436)
against the installed cv2 on random limbs and against canvases drawn by the UNMODIFIED reference ``plot_pose``
(tests/golden/canvas_golden.npz, made by tests/golden/make_canvas_golden.py).  Byte work: exact."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import draw_oracle  # noqa: E402
from lightweight_openpose_b200.src import pose_decode as mirror  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def base_frame(H, W):
    y, x, c = np.meshgrid(np.arange(H), np.arange(W), np.arange(3), indexing="ij")
    return ((x * 3 + y * 5 + c * 70) % 251).astype(np.uint8)


@pytest.fixture(scope="module")
def canvas_golden():
    return np.load(os.path.join(GOLDEN, "canvas_golden.npz"), allow_pickle=True)


@pytest.fixture(scope="module")
def decode_golden():
    return np.load(os.path.join(GOLDEN, "decode_golden_noipp.npz"), allow_pickle=True)


def golden_tables(name, canvas_golden, decode_golden):
    src = canvas_golden if f"{name}/persons" in canvas_golden.files else decode_golden
    return tuple(int(v) for v in src[f"{name}/hw"]), src[f"{name}/joint_list"], src[f"{name}/persons"]


def test_palette_and_limb_table_are_the_references():
    # the tables the rasterisers use: reference pose_decode.py:17-18,26-31 (exported by the mirror module)
    assert len(mirror.colors) == 16 and mirror.colors[6] == [100, 60, 39] and mirror.colors[15] == [100, 20, 0]
    assert mirror.joint_to_limb_heatmap_relationship[9:] == [[13, 0], [13, 3], [13, 6], [13, 9]]


def test_draw_limb_matches_cv2_on_random_limbs():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for t in range(1500):
        h, w = int(rng.integers(8, 80)), int(rng.integers(8, 80))
        img = rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
        lo, hi = (-12, 12) if t % 3 == 0 else (0, 0)         # a third of the limbs start or end off the frame
        j0 = (int(rng.integers(lo, w + hi)), int(rng.integers(lo, h + hi)))
        j1 = (int(rng.integers(lo, w + hi)), int(rng.integers(lo, h + hi)))
        if t % 7 == 0:                                       # very short and zero-length limbs
            j1 = (j0[0] + int(rng.integers(-2, 3)), j0[1] + int(rng.integers(-2, 3)))
        col = tuple(int(c) for c in rng.integers(0, 255, 3))
        a, b = img.copy(), img.copy()
        cv2.circle(a, j0, 3, (255, 255, 255), thickness=-1)
        cv2.circle(a, j1, 3, (255, 255, 255), thickness=-1)
        cv2.line(a, j0, j1, color=col, thickness=2)
        draw_oracle.draw_limb(b, j0, j1, col)
        assert np.array_equal(a, b), (h, w, j0, j1)


def test_oracle_canvas_equals_reference_golden(canvas_golden, decode_golden):
    for name in canvas_golden["names"]:
        name = str(name)
        (H, W), jl, persons = golden_tables(name, canvas_golden, decode_golden)
        canvas = draw_oracle.plot_pose_canvas(base_frame(H, W), jl, persons, mirror.joint_to_limb_heatmap_relationship,
                                              mirror.colors)
        assert np.array_equal(canvas, canvas_golden[f"{name}/canvas"]), name


def test_mirror_plot_pose_host_path_equals_reference_golden(canvas_golden, decode_golden):
    pytest.importorskip("cv2")
    for name in ("three_368", "empty_368", "border_64x80"):
        (H, W), jl, persons = golden_tables(name, canvas_golden, decode_golden)
        canvas, joints = mirror.plot_pose(base_frame(H, W), jl, persons)
        assert np.array_equal(canvas, canvas_golden[f"{name}/canvas"]), name
        assert joints.shape == (persons.shape[0] if persons.size else 0, 14, 3) and joints.dtype == np.float32
