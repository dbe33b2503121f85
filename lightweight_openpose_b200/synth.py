"""Synthetic workloads: camera-like input images and stick-figure label scenes.

There is no dataset and no network access, so benchmarks, the smoke test and
the parity tests draw their inputs from here.

* :func:`synth_image` -- a seeded uint8 RGB frame with natural-image-like
  (multi-octave, ~1/f) spatial statistics plus a few hard-edged rectangles.
  The reference feeds ``uint8 / 255.`` frames (reference
  ``test&eval/model_test.py:63-65``); iid uniform noise would be averaged away
  by the stride-8 network and leave spatially constant maps, which is not what
  the decode stage sees in practice.
* :func:`synth_scene` -- keypoint heatmaps and part-affinity fields for N
  stick-figure persons, built the way the reference builds its training labels
  (reference ``src/get_heatmap.py:13-107``: per joint channel the max of
  truncated Gaussians, sigma 1 at 1/8 resolution, cut-off 4.6052;
  ``src/get_paf.py:12-100``: per limb the unit vector field within
  ``paf_width_thre`` of the segment, averaged over persons), restated here so
  the generator also runs on machines where the reference tree is absent.  A
  little noise is added so that no two above-threshold pixels tie exactly
  (SURVEY.md 8(a) parity hazard 1).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

# limb table: (src joint, dst joint) per limb type -- reference
# src/pose_decode.py:17-18 == src/get_paf.py:31-32
LIMBS = [(0, 1), (1, 2), (3, 4), (4, 5), (6, 7), (7, 8), (9, 10), (10, 11),
         (12, 13), (13, 0), (13, 3), (13, 6), (13, 9)]
NUM_JOINTS = 14


def synth_image(seed: int, height: int = 368, width: int = 368) -> np.ndarray:
    """uint8 [H, W, 3] RGB frame."""
    import cv2
    rng = np.random.default_rng([seed, height, width, 7])
    acc = np.zeros((height, width, 3), dtype=np.float32)
    for s in (4, 8, 16, 32, 64, 128):
        gh, gw = -(-height // s) + 1, -(-width // s) + 1
        grid = rng.random((gh, gw, 3), dtype=np.float32) - 0.5
        up = cv2.resize(grid, (gw * s, gh * s), interpolation=cv2.INTER_LINEAR)
        acc += math.sqrt(s) * up[:height, :width]
    acc -= acc.min()
    acc /= max(float(acc.max()), 1e-6)
    for _ in range(6):                                   # a few hard edges
        y0, x0 = int(rng.integers(0, height - 8)), int(rng.integers(0, width - 8))
        hh, ww = int(rng.integers(8, height // 3)), int(rng.integers(8, width // 3))
        col = rng.random(3, dtype=np.float32)
        a = float(rng.uniform(0.3, 0.9))
        acc[y0:y0 + hh, x0:x0 + ww] = (1 - a) * acc[y0:y0 + hh, x0:x0 + ww] + a * col
    acc += (rng.random(acc.shape, dtype=np.float32) - 0.5) * 0.06   # sensor noise
    return np.clip(np.rint(acc * 255.0), 0, 255).astype(np.uint8)


def synth_batch(seeds, height: int = 368, width: int = 368) -> np.ndarray:
    """float32 [B, H, W, 3] in [0, 1] = uint8 / 255 (reference model_test.py:65)."""
    return np.stack([synth_image(s, height, width) for s in seeds]).astype(np.float32) / np.float32(255.0)


# ----------------------------------------------------------------------------
# label-style scenes
# ----------------------------------------------------------------------------
# A crude standing figure in a unit box (x right, y down), AI-Challenger joint
# order: 0-2 right arm (shoulder, elbow, wrist), 3-5 left arm, 6-8 right leg
# (hip, knee, ankle), 9-11 left leg, 12 head top, 13 neck.
_TEMPLATE = np.array([
    [0.30, 0.25], [0.22, 0.42], [0.18, 0.58],
    [0.70, 0.25], [0.78, 0.42], [0.82, 0.58],
    [0.38, 0.58], [0.36, 0.78], [0.35, 0.97],
    [0.62, 0.58], [0.64, 0.78], [0.65, 0.97],
    [0.50, 0.03], [0.50, 0.20],
], dtype=np.float64)


def synth_keypoints(seed: int, num_persons: int, height: int, width: int) -> np.ndarray:
    """[P, 14, 3] keypoints (x, y, v=1) in image pixels; figures are jittered
    copies of a template, placed on a jittered grid so they rarely overlap."""
    rng = np.random.default_rng([seed, num_persons, height, width, 11])
    kps = np.zeros((num_persons, NUM_JOINTS, 3), dtype=np.float64)
    cols = max(1, int(math.ceil(math.sqrt(num_persons * width / height))))
    rows = max(1, -(-num_persons // cols))
    cw, ch = width / cols, height / rows
    for p in range(num_persons):
        r, c = divmod(p, cols)
        bw = cw * rng.uniform(0.70, 0.92)
        bh = ch * rng.uniform(0.75, 0.95)
        ox = c * cw + rng.uniform(0.02, 0.98) * (cw - bw)
        oy = r * ch + rng.uniform(0.02, 0.98) * (ch - bh)
        pts = _TEMPLATE + rng.normal(0, 0.025, size=_TEMPLATE.shape)
        if rng.random() < 0.5:
            pts[:, 0] = 1.0 - pts[:, 0]
        xy = np.stack([ox + pts[:, 0] * bw, oy + pts[:, 1] * bh], axis=1)
        xy[:, 0] = np.clip(xy[:, 0], 1.0, width - 2.0)
        xy[:, 1] = np.clip(xy[:, 1], 1.0, height - 2.0)
        kps[p, :, :2] = xy
        kps[p, :, 2] = 1
    return kps


def _gaussian_max(hm: np.ndarray, cx: float, cy: float, sigma: float) -> None:
    th = 4.6052
    delta = math.sqrt(th * 2)
    h, w = hm.shape
    x0 = int(max(0, cx - delta * sigma + 0.5))
    y0 = int(max(0, cy - delta * sigma + 0.5))
    x1 = int(min(w - 1, cx + delta * sigma + 0.5))
    y1 = int(min(h - 1, cy + delta * sigma + 0.5))
    if x1 < x0 or y1 < y0:
        return
    ef = 1.0 / 2.0 / sigma / sigma
    xv = (np.arange(x0, x1 + 1) - cx) ** 2
    yv = (np.arange(y0, y1 + 1) - cy) ** 2
    s = ef * (xv[None, :] + yv[:, None])
    g = np.exp(-s)
    g[s > th] = 0
    hm[y0:y1 + 1, x0:x1 + 1] = np.maximum(hm[y0:y1 + 1, x0:x1 + 1], g)


def label_heatmaps(kps: np.ndarray, height: int, width: int, hh: int, hw: int,
                   sigma: float = 1.0) -> np.ndarray:
    fx, fy = hw / width, hh / height
    out = np.zeros((hh, hw, NUM_JOINTS), dtype=np.float32)
    for j in range(NUM_JOINTS):
        ch = np.zeros((hh, hw), dtype=np.float32)
        for p in range(kps.shape[0]):
            cx, cy = kps[p, j, 0] * fx, kps[p, j, 1] * fy
            if cx >= hw or cy >= hh or cx < 0 or cy < 0 or (cx == 0 and cy == 0):
                continue
            _gaussian_max(ch, cx, cy, sigma)
        out[:, :, j] = ch
    return out


def label_pafs(kps: np.ndarray, height: int, width: int, hh: int, hw: int,
               thresh: float = 1.0) -> np.ndarray:
    fx, fy = hw / width, hh / height
    out = np.zeros((hh, hw, 2 * len(LIMBS)), dtype=np.float32)
    ys, xs = np.mgrid[0:hh, 0:hw]
    for li, (ja, jb) in enumerate(LIMBS):
        acc = np.zeros((2, hh, hw), dtype=np.float64)
        count = np.zeros((hh, hw), dtype=np.float64)
        for p in range(kps.shape[0]):
            a = np.array([kps[p, ja, 0] * fx, kps[p, ja, 1] * fy])
            b = np.array([kps[p, jb, 0] * fx, kps[p, jb, 1] * fy])
            norm = float(np.linalg.norm(b - a))
            if norm == 0.0:
                continue
            u = (b - a) / norm
            x_lo = max(int(round(min(a[0], b[0]) - thresh)), 0)
            x_hi = min(int(round(max(a[0], b[0]) + thresh)), hw)
            y_lo = max(int(round(min(a[1], b[1]) - thresh)), 0)
            y_hi = min(int(round(max(a[1], b[1]) + thresh)), hh)
            box = (xs >= x_lo) & (xs < x_hi) & (ys >= y_lo) & (ys < y_hi)
            dist = np.abs((xs - a[0]) * u[1] - (ys - a[1]) * u[0])
            mask = box & (dist < thresh)
            acc[0][mask] += u[0]
            acc[1][mask] += u[1]
            count[mask & (np.abs(u[0]) > 0)] += 1     # reference counts via the x component only
        count[count == 0] = 1
        out[:, :, 2 * li] = (acc[0] / count).astype(np.float32)
        out[:, :, 2 * li + 1] = (acc[1] / count).astype(np.float32)
    return out


def synth_scene(seed: int, num_persons: int, height: int = 368, width: int = 368,
                noise: float = 0.02) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns ``(heatmaps [h,w,14] f32, pafs [h,w,26] f32, keypoints [P,14,3])``
    at stride 8.  ``noise`` adds U(0, noise) to heatmaps and N(0, noise/2) to
    PAFs (tie breaker)."""
    hh, hw = -(-height // 8), -(-width // 8)
    kps = synth_keypoints(seed, num_persons, height, width)
    heat = label_heatmaps(kps, height, width, hh, hw)
    paf = label_pafs(kps, height, width, hh, hw)
    rng = np.random.default_rng([seed, 13])
    if noise > 0:
        heat = heat + rng.uniform(0.0, noise, size=heat.shape).astype(np.float32)
        paf = paf + rng.normal(0.0, noise / 2, size=paf.shape).astype(np.float32)
    return heat.astype(np.float32), paf.astype(np.float32), kps


# joint types joined by each limb: joint_to_limb_heatmap_relationship of the 14-joint variant (src/pose_decode.py:16-17)
_LIMB_JOINTS = [[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8], [9, 10], [10, 11], [12, 13], [13, 0], [13, 3], [13, 6],
                [13, 9]]


def random_limb_tables(seed, joints_per_type, density):
    """Adversarial input of the grouping stage: random connections between random joints, every source joint at most
    once per limb type (what find_connected_joints emits), destinations may repeat (:295 keeps them)."""
    rng = np.random.default_rng(seed)
    rows, nid = [], 0
    ids = []
    for jt in range(14):
        n = int(joints_per_type)
        ids.append(np.arange(nid, nid + n))
        for _ in range(n):
            rows.append((float(rng.integers(0, 368)), float(rng.integers(0, 368)), float(np.float32(rng.uniform(0.1, 1.0))),
                         float(nid), 0.0, float(jt)))
            nid += 1
    joint_list = np.array(rows, dtype=np.float64).reshape(-1, 6)
    limbs = []
    for t in range(13):
        ja, jb = _LIMB_JOINTS[t]
        out = []
        for i, a in enumerate(ids[ja]):
            if rng.uniform() < density and len(ids[jb]):
                j = int(rng.integers(0, len(ids[jb])))
                out.append((float(a), float(ids[jb][j]), float(rng.uniform(0.05, 1.0)), float(i), float(j)))
        limbs.append(np.array(out, dtype=np.float64).reshape(-1, 5))
    return limbs, joint_list
