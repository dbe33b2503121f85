"""``decode_pose`` and friends -- drop-in for the reference ``src/pose_decode.py``.

``decode_pose(img_orig, param, heatmaps, pafs)`` keeps the reference signature
and return value ``(canvas, joint_list, person_to_joint_assoc, joints)``
(reference ``src/pose_decode.py:460,517``): ``joint_list`` float64 ``(N, 6)``
(``(0,)`` when empty), ``person_to_joint_assoc`` float64 ``(P, 16)`` (``(0,)`` when
empty), ``joints`` float32 ``(P, 42)``.  The peak search, patch refinement, PAF
line-integral scoring and grouping run in the CUDA decode chain (``lwop_decode``);
only the canvas drawing of ``plot_pose`` stays on the host (cv2), as the survey's
scope table allows.

``decode_pose_batch`` is the batched form the GPU path is built for.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from .. import _lib, engine, graph

# tables -- reference pose_decode.py:17-34
joint_to_limb_heatmap_relationship = [[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8],
                                      [9, 10], [10, 11], [12, 13], [13, 0], [13, 3], [13, 6], [13, 9]]
paf_xy_coords_per_limb = [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9], [10, 11], [12, 13], [14, 15], [16, 17],
                          [18, 19], [20, 21], [22, 23], [24, 25]]
colors = [[255, 0, 0], [0, 255, 0], [0, 0, 255],
          [255, 0, 255], [0, 255, 255], [255, 0, 255],
          [100, 60, 39], [35, 200, 111], [200, 40, 145],
          [238, 248, 220], [255, 165, 100], [20, 205, 50],
          [0, 191, 255], [0, 0, 205], [220, 20, 60],
          [100, 20, 0]]
NUM_JOINTS = 14
NUM_LIMBS = len(joint_to_limb_heatmap_relationship)

_decode_engines: Dict[tuple, "engine.Engine"] = {}


def _torch():
    import torch
    return torch


def _decode_engine(img_h: int, img_w: int, batch: int, device: int = 0) -> "engine.Engine":
    """Decode-only engines are keyed by frame size; no weights are needed."""
    key = (device, img_h, img_w)
    e = _decode_engines.get(key)
    if e is None or e.max_batch < batch:
        if e is not None:
            e.close()
        e = engine.Engine(max(img_h, 16), max(img_w, 16), max_batch=batch, device=device, use_graph=False)
        _decode_engines[key] = e
    return e


def compute_resized_coords(coords, resizeFactor):
    """Index of a cell after resizing the array by ``resizeFactor`` (reference :80-104)."""
    return (np.array(coords, dtype=float) + 0.5) * resizeFactor - 0.5


def decode_pose_batch(img_hw, param, heatmaps, pafs, device: int = 0) -> List["engine.DecodeResult"]:
    """heatmaps [B,h,w,14], pafs [B,h,w,26] float32 (CUDA tensors or numpy);
    ``img_hw`` = (H, W) of the frames the maps belong to."""
    torch = _torch()
    engine.require_gpu(device)
    dev = torch.device("cuda", device)

    def to_dev(a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return a.to(dev, dtype=torch.float32).contiguous()

    heat, paf = to_dev(heatmaps), to_dev(pafs)
    if heat.dim() != 4 or heat.shape[-1] != NUM_JOINTS or paf.shape[-1] != 2 * NUM_LIMBS:
        raise ValueError("heatmaps must be [B,h,w,14] and pafs [B,h,w,26]")
    H, W = int(img_hw[0]), int(img_hw[1])
    e = _decode_engine(H, W, heat.shape[0], device)
    return e.decode(heat, paf, (H, W), float(param['thre1']), float(param['thre2']))


def plot_pose(img_orig, joint_list, person_to_joint_assoc, bool_fast_plot=True):
    """Keypoint table ``(P, 14, 3)`` float32 and the drawn canvas (reference :399-457)."""
    import cv2
    canvas = img_orig.copy()
    persons = np.asarray(person_to_joint_assoc)
    n = persons.shape[0]
    joints = np.zeros((n, 14, 3), dtype=np.float32)
    for k in range(n):
        row = persons[k]
        for a, b in joint_to_limb_heatmap_relationship:
            ids = row[[a, b]].astype(int)
            if -1 in ids:
                continue
            pts = joint_list[ids, 0:2]
            for jid, pt in zip(ids, pts):
                joints[k, int(joint_list[jid, -1]), :] = [pt[0], pt[1], 1]
            for pt in pts:
                cv2.circle(canvas, tuple(pt[0:2].astype(int)), 3, (255, 255, 255), thickness=-1)
            cv2.line(canvas, tuple(pts[0].astype(int)), tuple(pts[1].astype(int)),
                     color=colors[k % len(colors)], thickness=2)
    return canvas, joints


def decode_pose(img_orig, param, heatmaps, pafs):
    """Reference :460-517.  ``heatmaps [h,w,14]``, ``pafs [h,w,26]`` (numpy or CUDA tensors)."""
    H, W = img_orig.shape[0], img_orig.shape[1]
    res = decode_pose_batch((H, W), param, heatmaps[None], pafs[None])[0]
    joint_list, persons = res.joint_list, res.person_to_joint_assoc
    canvas, _ = plot_pose(img_orig, joint_list, persons) if persons.size else (img_orig.copy(), None)
    return canvas, joint_list, persons, res.joints


def NMS(param, heatmaps, upsampFactor=1., bool_refine_center=True, bool_gaussian_filt=False):
    """Per joint type an ``(n, 5)`` float64 array ``(x, y, score, id, 0)`` (reference :107-185),
    computed by the CUDA peak kernel (greedy radius-5 search of ``find_peaks_v2`` +
    bicubic patch refinement)."""
    if not bool_refine_center or bool_gaussian_filt:
        raise NotImplementedError("only the reference's default flags (refine centre, no Gaussian filter) "
                                  "are on the GPU path")
    h, w = heatmaps.shape[0], heatmaps.shape[1]
    Hf = h * float(upsampFactor)
    if abs(Hf - round(Hf)) > 1e-9:
        raise NotImplementedError("upsampFactor * map height must be an integer frame height")
    H, W = int(round(Hf)), int(round(w * float(upsampFactor)))
    torch = _torch()
    paf0 = torch.zeros((1, h, w, 2 * NUM_LIMBS), dtype=torch.float32, device="cuda")
    res = decode_pose_batch((H, max(W, w)), param, np.asarray(heatmaps)[None] if isinstance(heatmaps, np.ndarray)
                            else heatmaps[None], paf0)[0]
    out = []
    jl = res.joint_list
    for j in range(NUM_JOINTS):
        rows = jl[jl[:, 5] == j][:, :5] if jl.size else np.zeros((0, 5))
        out.append(np.ascontiguousarray(rows))
    return out


def find_peaks_v2(param, img):
    """``[[x, y], ...]`` of the greedy radius-5 peak search (reference :52-76), via the CUDA kernel."""
    img = np.asarray(img, dtype=np.float32)
    maps = np.zeros(img.shape + (NUM_JOINTS,), dtype=np.float32)
    maps[:, :, 0] = img
    rows = NMS_lowres(param, maps)[0]
    return rows[:, :2].astype(np.int64)


def NMS_lowres(param, heatmaps):
    """Peaks at map resolution (factor 1: the 'patch refinement' of a 1x patch is the identity)."""
    return NMS(param, heatmaps, 1.0)
