"""Generates ``tests/golden/canvas_golden.npz``: canvases drawn by the UNMODIFIED reference ``plot_pose``
(``/root/reference/src/pose_decode.py:399-457``, i.e. the installed cv2's ``circle`` / ``line``) for the joint lists and
person tables of ``decode_golden_noipp.npz`` plus hand-made tables that put limbs on the frame border, on top of each
other and at zero length.

Run in the build container only: ``python tests/golden/make_canvas_golden.py``.  The base frame is a formula of
(y, x, c) (``base_frame``), so only the canvases are stored.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")


def base_frame(H, W):
    y, x, c = np.meshgrid(np.arange(H), np.arange(W), np.arange(3), indexing="ij")
    return ((x * 3 + y * 5 + c * 70) % 251).astype(np.uint8)


def border_tables(H, W, seed):
    """One joint of every type per person; persons hug the borders / corners, overlap, or collapse to a point."""
    rng = np.random.default_rng(seed)
    persons, joints = [], []
    for p in range(18):                      # more than 16 persons: the palette wraps (colors[person % 16])
        kind = p % 6
        row = -np.ones(16)
        for t in range(14):
            if kind == 0:      # anywhere
                x, y = rng.integers(0, W), rng.integers(0, H)
            elif kind == 1:    # left / right border columns
                x, y = rng.choice([0, 1, W - 2, W - 1]), rng.integers(0, H)
            elif kind == 2:    # top / bottom border rows
                x, y = rng.integers(0, W), rng.choice([0, 1, H - 2, H - 1])
            elif kind == 3:    # corners
                x, y = rng.choice([0, W - 1]), rng.choice([0, H - 1])
            elif kind == 4:    # all joints within a few pixels: zero-length and 1-2 pixel limbs
                x, y = W // 2 + rng.integers(-2, 3), H // 2 + rng.integers(-2, 3)
            else:              # long near-horizontal / near-vertical limbs
                x, y = (rng.integers(0, W), H // 3 + rng.integers(-1, 2)) if t % 2 else (W // 3 + rng.integers(-1, 2), rng.integers(0, H))
            if kind == 0 and t in (2, 7):    # missing joints: those limbs are skipped (:413-416)
                continue
            row[t] = len(joints)
            joints.append([float(x), float(y), 0.5, float(len(joints)), 0.0, float(t)])
        row[14], row[15] = 10.0, 14.0
        persons.append(row)
    return np.array(joints, dtype=np.float64), np.array(persons, dtype=np.float64)


def main():
    warnings.simplefilter("ignore")
    import cv2
    cv2.setNumThreads(1)
    from src import pose_decode as ref            # the reference module itself
    z = np.load(os.path.join(HERE, "decode_golden_noipp.npz"), allow_pickle=True)
    out, names = {}, []
    for name in z["names"]:
        name = str(name)
        H, W = (int(v) for v in z[f"{name}/hw"])
        jl, persons = z[f"{name}/joint_list"], z[f"{name}/persons"]
        canvas, _ = ref.plot_pose(base_frame(H, W), jl, persons)
        out[f"{name}/canvas"] = canvas
        names.append(name)
        print(name, persons.shape, int((canvas != base_frame(H, W)).any(-1).sum()), "pixels drawn")
    for k, (H, W) in enumerate([(64, 80), (368, 368), (37, 23), (368, 656)]):
        name = f"border_{H}x{W}"
        jl, persons = border_tables(H, W, 100 + k)
        canvas, _ = ref.plot_pose(base_frame(H, W), jl, persons)
        out[f"{name}/canvas"], out[f"{name}/joint_list"], out[f"{name}/persons"] = canvas, jl, persons
        out[f"{name}/hw"] = np.array([H, W], dtype=np.int64)
        names.append(name)
        print(name, persons.shape, int((canvas != base_frame(H, W)).any(-1).sum()), "pixels drawn")
    out["names"] = np.array(names)
    out["versions"] = np.array([cv2.__version__, np.__version__])
    path = os.path.join(HERE, "canvas_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
