"""``decode_pose`` and friends -- drop-in for the reference ``src/pose_decode.py``.

``decode_pose(img_orig, param, heatmaps, pafs)`` keeps the reference signature
and return value ``(canvas, joint_list, person_to_joint_assoc, joints)``
(reference ``src/pose_decode.py:460,517``): ``joint_list`` float64 ``(N, 6)``
(``(0,)`` when empty), ``person_to_joint_assoc`` float64 ``(P, 16)`` (``(0,)`` when
empty), ``joints`` float32 ``(P, 42)``.  The peak search, patch refinement, PAF
line-integral scoring and grouping run in the CUDA decode chain (``lwop_decode``);
the canvas drawing of ``plot_pose`` runs on the host with cv2 by default, as the survey's
scope table allows, or in the CUDA rasteriser (``lwop_draw_poses``: ``LWOP_DRAW=gpu`` /
``plot_pose_device`` / ``Engine.draw_poses``), which paints the same bytes.

``decode_pose_batch`` is the batched form the GPU path is built for.
"""
from __future__ import annotations

import math
import os
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


def _keypoint_table(joint_list, persons) -> np.ndarray:
    """``joints (P, 14, 3)`` of ``plot_pose`` (reference :403-429)."""
    n = persons.shape[0]
    joints = np.zeros((n, 14, 3), dtype=np.float32)
    for k in range(n):
        row = persons[k]
        for a, b in joint_to_limb_heatmap_relationship:
            ids = row[[a, b]].astype(int)
            if -1 in ids:
                continue
            for jid, pt in zip(ids, joint_list[ids, 0:2]):
                joints[k, int(joint_list[jid, -1]), :] = [pt[0], pt[1], 1]
    return joints


def plot_pose_device(img_orig, joint_list, person_to_joint_assoc, device: int = 0) -> np.ndarray:
    """The canvas of ``plot_pose`` rasterised by the CUDA kernel (``lwop_draw_poses``): same circles and lines, byte
    for byte, as cv2 draws them (reference :431-436).  Host arrays in, host canvas out; the batched, device-resident
    form is :meth:`engine.Engine.draw_poses`."""
    torch = _torch()
    engine.require_gpu(device)
    dev = torch.device("cuda", device)
    img = np.ascontiguousarray(img_orig)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("img_orig must be uint8 [H, W, 3]")
    persons = np.asarray(person_to_joint_assoc, dtype=np.float64)
    persons = persons.reshape(-1, 16) if persons.size else np.zeros((0, 16))
    jl = np.asarray(joint_list, dtype=np.float64)
    jl = jl.reshape(-1, 6) if jl.size else np.zeros((0, 6))
    if persons.shape[0] > _lib.MAX_PERSONS or jl.shape[0] > _lib.MAX_JOINTS:
        raise _lib.LwopError(_lib.ERR_OVERFLOW, "more persons / joints than the device tables hold")
    ids = persons[:, :14]
    if ids.size and (ids.max() >= jl.shape[0] or ids.min() < -1):
        raise IndexError("person table refers to a joint id outside joint_list")      # the reference's fancy index raises too
    counts = np.zeros((1, _lib.COUNTS), dtype=np.int32)
    counts[0, 0], counts[0, 1] = jl.shape[0], persons.shape[0]
    tj = np.zeros((1, _lib.MAX_JOINTS, 6), dtype=np.float64)
    tj[0, :jl.shape[0]] = jl
    tp = np.zeros((1, _lib.MAX_PERSONS, 16), dtype=np.float64)
    tp[0, :persons.shape[0]] = persons
    tables = {"counts": torch.from_numpy(counts).to(dev), "joints": torch.from_numpy(tj).to(dev),
              "persons": torch.from_numpy(tp).to(dev)}
    e = _decode_engine(16, 16, 1, device)
    canvas = e.draw_poses(torch.from_numpy(img).to(dev)[None], tables)
    return canvas[0].cpu().numpy()


def plot_pose(img_orig, joint_list, person_to_joint_assoc, bool_fast_plot=True):
    """Keypoint table ``(P, 14, 3)`` float32 and the drawn canvas (reference :399-457).  The canvas is drawn on the
    host with cv2, as the reference does, unless ``LWOP_DRAW=gpu`` selects the CUDA rasteriser
    (:func:`plot_pose_device`; same bytes)."""
    persons = np.asarray(person_to_joint_assoc)
    if persons.size == 0:
        persons = np.zeros((0, 16))
    joints = _keypoint_table(joint_list, persons)
    if os.environ.get("LWOP_DRAW", "") == "gpu":
        return plot_pose_device(img_orig, joint_list, persons), joints
    import cv2
    canvas = img_orig.copy()
    for k in range(persons.shape[0]):
        row = persons[k]
        for a, b in joint_to_limb_heatmap_relationship:
            ids = row[[a, b]].astype(int)
            if -1 in ids:
                continue
            pts = joint_list[ids, 0:2]
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


def _heat_to_dev(heatmaps):
    torch = _torch()
    engine.require_gpu(0)
    if isinstance(heatmaps, np.ndarray):
        heatmaps = torch.from_numpy(np.ascontiguousarray(heatmaps, dtype=np.float32))
    return heatmaps.to(torch.device("cuda", 0), dtype=torch.float32).contiguous()


def _check_overflow(counts) -> None:
    if int((counts[:, 2] & 1).max()) != 0:
        raise _lib.LwopError(_lib.ERR_OVERFLOW, f"more than {_lib.MAX_PEAKS} peaks in one joint channel")


def _peak_lists(peaks, pc, n_channels=NUM_JOINTS):
    """Device peak table of ONE image -> the reference's list of ``(n, 5)`` float64 arrays
    ``(x, y, score, id, 0)`` per joint type, ids = running counter (reference :135,182-183)."""
    pk = peaks[0].cpu().numpy()
    cnt = pc[0].cpu().numpy()
    out, uid = [], 0
    for j in range(n_channels):
        n = int(cnt[j])
        rows = np.zeros((n, 5), dtype=np.float64)
        rows[:, :3] = pk[j, :n, :3]
        rows[:, 3] = np.arange(uid, uid + n)
        uid += n
        out.append(rows)
    return out


def NMS(param, heatmaps, upsampFactor=1., bool_refine_center=True, bool_gaussian_filt=False):
    """Per joint type an ``(n, 5)`` float64 array ``(x, y, score, id, 0)`` (reference :107-185), computed by
    the CUDA peak kernel: greedy radius-5 search of ``find_peaks_v2`` (:146) and, with ``bool_refine_center``,
    the bicubic refinement of the clipped 5x5 patch (:149-175); without it the low-res peak snapped to the
    up-scaled grid (:176-182).  ``bool_gaussian_filt`` smooths the up-sampled patch with scipy's
    ``gaussian_filter(sigma=3)`` arithmetic before the argmax (:163-164); like the reference it only has an effect
    together with ``bool_refine_center``."""
    heat = _heat_to_dev(heatmaps)
    if heat.dim() != 3 or heat.shape[-1] != NUM_JOINTS:
        raise ValueError("heatmaps must be [h, w, 14]")
    h, w = int(heat.shape[0]), int(heat.shape[1])
    e = _decode_engine(max(16, h), max(16, w), 1)
    mode = 0 if bool_refine_center else _lib.PEAKS_NO_REFINE
    if bool_gaussian_filt and bool_refine_center:
        mode |= _lib.PEAKS_GAUSS
    peaks, pc, counts = e.nms_device(heat[None], float(upsampFactor), float(param['thre1']), mode)
    _check_overflow(counts.cpu().numpy())
    return _peak_lists(peaks, pc)


def _single_channel_peaks(param, img, mode):
    torch = _torch()
    img = _heat_to_dev(np.asarray(img, dtype=np.float32) if not hasattr(img, "device") else img)
    if img.dim() != 2:
        raise ValueError("img must be a 2d map")
    maps = torch.full(tuple(img.shape) + (NUM_JOINTS,), -np.inf, dtype=torch.float32, device=img.device)
    maps[:, :, 0] = img
    h, w = int(img.shape[0]), int(img.shape[1])
    e = _decode_engine(max(16, h), max(16, w), 1)
    peaks, pc, counts = e.nms_device(maps[None], 1.0, float(param['thre1']), mode | _lib.PEAKS_NO_REFINE)
    _check_overflow(counts.cpu().numpy())
    n = int(pc[0, 0])
    return peaks[0, 0, :n, :2].cpu().numpy().astype(np.int64)


def find_peaks_v2(param, img):
    """``[[x, y], ...]`` of the greedy radius-5 peak search, best first (reference :52-76)."""
    return _single_channel_peaks(param, img, 0)


def find_peaks(param, img):
    """``[[x, y], ...]`` of the local maxima over the 3x3 cross footprint above ``thre1`` in row-major order
    (reference :37-49; unused by the reference's own NMS, kept for API completeness)."""
    return _single_channel_peaks(param, img, _lib.PEAKS_CROSS)


def _peak_table_from_lists(joint_list_per_joint_type, dev):
    """Reference-shaped list of 14 ``(n, >=3)`` arrays -> device peak table of one image.  Joint ids must be
    the running counter NMS assigns; they are re-derived from the per-type counts on the device."""
    torch = _torch()
    pk = np.zeros((1, NUM_JOINTS, _lib.MAX_PEAKS, 4), dtype=np.float32)
    pc = np.zeros((1, 16), dtype=np.int32)
    uid = 0
    for j, rows in enumerate(joint_list_per_joint_type):
        rows = np.asarray(rows, dtype=np.float64).reshape(-1, 5) if len(rows) else np.zeros((0, 5))
        n = rows.shape[0]
        if n > _lib.MAX_PEAKS:
            raise _lib.LwopError(_lib.ERR_OVERFLOW, f"more than {_lib.MAX_PEAKS} joints of type {j}")
        if n and not np.array_equal(rows[:, 3], np.arange(uid, uid + n)):
            raise ValueError("joint ids must be the running counter assigned by NMS (pose_decode.py:135,182-183)")
        pk[0, j, :n, :3] = rows[:, :3]
        pc[0, j] = n
        uid += n
    return torch.from_numpy(pk).to(dev), torch.from_numpy(pc).to(dev)


def find_connected_joints(param, paf_upsamp, joint_list_per_joint_type, num_intermed_pts=10, max_dist_thresh=1,
                          max_paf_score_thresh=0.7):
    """Per limb type an ``(n, 5)`` float64 array ``(src_id, dst_id, score, i, j)`` -- or ``[]`` when either joint
    list is empty -- exactly as the reference returns (:188-301).  ``paf_upsamp`` is the PAF map already
    up-sampled to the frame ``[H, W, 26]`` (what ``decode_pose`` hands over, :487,493)."""
    if not 1 <= int(num_intermed_pts) <= 64:
        raise ValueError("num_intermed_pts must be in 1..64")
    torch = _torch()
    engine.require_gpu(0)
    dev = torch.device("cuda", 0)
    paf = paf_upsamp
    if isinstance(paf, np.ndarray):
        paf = torch.from_numpy(np.ascontiguousarray(paf, dtype=np.float32))
    paf = paf.to(dev, dtype=torch.float32).contiguous()
    if paf.dim() != 3 or paf.shape[-1] != 2 * NUM_LIMBS:
        raise ValueError("paf_upsamp must be [H, W, 26]")
    H, W = int(paf.shape[0]), int(paf.shape[1])
    peaks, pc = _peak_table_from_lists(joint_list_per_joint_type, dev)
    counts = torch.zeros((1, _lib.COUNTS), dtype=torch.int32, device=dev)
    e = _decode_engine(max(16, H), max(16, W), 1)
    limbs = e.connect_device(paf[None], True, (H, W), (H, W), float(param['thre2']), peaks, pc, counts,
                             int(num_intermed_pts), float(max_dist_thresh))
    lim, cnt = limbs[0].cpu().numpy(), counts[0].cpu().numpy()
    out = []
    for t, (a, b) in enumerate(joint_to_limb_heatmap_relationship):
        if len(joint_list_per_joint_type[a]) == 0 or len(joint_list_per_joint_type[b]) == 0:
            out.append([])
        else:
            out.append(np.array(lim[t, :int(cnt[18 + t])], copy=True))
    return out


def group_limbs_of_same_person(connected_limbs, joint_list):
    """``(P, 16)`` float64 person table (``(0,)`` when nobody is found) from the limb connections and the
    flattened ``(N, 6)`` joint list (reference :304-396), by the CUDA grouping kernel."""
    torch = _torch()
    engine.require_gpu(0)
    dev = torch.device("cuda", 0)
    jl = np.asarray(joint_list, dtype=np.float64)
    per_type = [np.zeros((0, 5))] * NUM_JOINTS
    if jl.size:
        jl = jl.reshape(-1, 6)
        per_type = [jl[jl[:, 5] == j][:, :5] for j in range(NUM_JOINTS)]
    peaks, pc = _peak_table_from_lists(per_type, dev)
    lim = np.zeros((1, NUM_LIMBS, _lib.MAX_PEAKS, 5), dtype=np.float64)
    counts = np.zeros((1, _lib.COUNTS), dtype=np.int32)
    for t, rows in enumerate(connected_limbs):
        rows = np.asarray(rows, dtype=np.float64).reshape(-1, 5) if len(rows) else np.zeros((0, 5))
        if rows.shape[0] > _lib.MAX_PEAKS:
            raise _lib.LwopError(_lib.ERR_OVERFLOW, f"more than {_lib.MAX_PEAKS} connections of limb type {t}")
        lim[0, t, :rows.shape[0]] = rows
        counts[0, 18 + t] = rows.shape[0]
    limbs_d, counts_d = torch.from_numpy(lim).to(dev), torch.from_numpy(counts).to(dev)
    e = _decode_engine(16, 16, 1)
    _, persons, _ = e.group_device(peaks, pc, limbs_d, counts_d)
    c = counts_d[0].cpu().numpy()
    if c[2] & 2:
        raise _lib.LwopError(_lib.ERR_OVERFLOW, f"more than {_lib.MAX_PERSONS} persons")
    n = int(c[1])
    return np.array(persons[0, :n].cpu().numpy(), copy=True) if n else np.array([])
