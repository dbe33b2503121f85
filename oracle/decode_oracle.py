"""ORACLE (test infrastructure, not product code) -- CPU restatement of the
reference pose decoder ``/root/reference/src/pose_decode.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package never does.

PARITY PINNED: ``tests/golden/decode_*.npz`` hold inputs and outputs of the
unmodified reference module (imported from ``/root/reference`` by
``tests/golden/make_decode_golden.py`` in the build container); this file is
checked against them in ``tests/test_oracle_decode.py``, and -- when the
reference tree is present -- against the live reference module on fresh scenes.

Third-party arithmetic on this path (unpinned by the reference, versions in the
build image): OpenCV 4.13.0 ``cv2.resize(..., INTER_CUBIC)`` (bicubic, A=-0.75,
half-pixel centres, replicated border), numpy 2.3.5 (``argsort``, ``linspace``,
``round`` = half-to-even, ``dot``, pairwise ``mean``).  The oracle calls the same
libraries; :func:`bicubic_resize_f32` additionally restates cv2's float32
two-pass bicubic in numpy (it agrees with cv2 to ~1 ulp, not bit for bit --
cv2's own code paths do not agree with each other either).

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

try:  # cv2 is present in the build/GPU image; the numpy restatement is the fallback
    import cv2
    cv2.setNumThreads(1)
except Exception:  # pragma: no cover
    cv2 = None

# reference pose_decode.py:17-34
LIMB_JOINTS = [[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8], [9, 10], [10, 11],
               [12, 13], [13, 0], [13, 3], [13, 6], [13, 9]]
LIMB_PAF_CHANNELS = [[2 * i, 2 * i + 1] for i in range(13)]
NUM_JOINTS = 14
NUM_LIMBS = 13
PALETTE = [[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 0, 255], [0, 255, 255], [255, 0, 255],
           [100, 60, 39], [35, 200, 111], [200, 40, 145], [238, 248, 220], [255, 165, 100],
           [20, 205, 50], [0, 191, 255], [0, 0, 205], [220, 20, 60], [100, 20, 0]]

NMS_RADIUS = 5          # pose_decode.py:53  dis_thres
PATCH_HALF = 2          # pose_decode.py:141 win_size
NUM_SAMPLES = 10        # pose_decode.py:188 num_intermed_pts
MIN_PARTS = 3           # pose_decode.py:384
MIN_MEAN_SCORE = 0.2    # pose_decode.py:384


# ----------------------------------------------------------------------------
# bicubic (cv2 INTER_CUBIC, float32) restated
# ----------------------------------------------------------------------------
def cubic_coeffs_f32(frac: np.float32) -> np.ndarray:
    """cv2 ``interpolateCubic`` with A = -0.75, evaluated in float32."""
    A = np.float32(-0.75)
    x = np.float32(frac)
    one = np.float32(1.0)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    xm = one - x
    c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def _axis_table(dst: int, src: int, scale: float) -> Tuple[np.ndarray, np.ndarray]:
    """Per destination index: the 4 clamped source indices and float32 taps.
    ``scale`` = src/dst ratio as cv2 computes it (``1/fx`` or ``src/dst``)."""
    idx = np.zeros((dst, 4), dtype=np.int64)
    wts = np.zeros((dst, 4), dtype=np.float32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(float(f)))
        f = np.float32(f - np.float32(s))
        wts[d] = cubic_coeffs_f32(f)
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, src - 1)
    return idx, wts


def bicubic_resize_f32(src: np.ndarray, dst_h: int, dst_w: int, scale_y: float, scale_x: float) -> np.ndarray:
    """float32 two-pass bicubic (horizontal 4-tap, then vertical 4-tap) of a
    [h, w] or [h, w, c] array, in the operation order of OpenCV's own resize engine
    (modules/imgproc/src/resize.cpp): the horizontal pass sums its four products left to
    right, ((v0*c0 + v1*c1) + v2*c2) + v3*c3, the vertical pass right to left,
    ((v3*c3 + v2*c2) + v1*c1) + v0*c0, no fused multiply-adds.  With that order this
    function equals ``cv2.resize(..., INTER_CUBIC)`` BIT FOR BIT when OpenCV runs its own
    code (``cv2.ipp.setUseIPP(False)``); IPP-accelerated builds differ from it by 1 ulp on
    about a third of the pixels (tests/test_oracle_decode.py pins both facts)."""
    a = np.asarray(src, dtype=np.float32)
    squeeze = a.ndim == 2
    if squeeze:
        a = a[:, :, None]
    xi, xw = _axis_table(dst_w, a.shape[1], scale_x)
    yi, yw = _axis_table(dst_h, a.shape[0], scale_y)
    rows = np.zeros((a.shape[0], dst_w, a.shape[2]), dtype=np.float32)
    for k in range(4):
        term = a[:, xi[:, k], :] * xw[:, k][None, :, None]
        rows = term if k == 0 else (rows + term).astype(np.float32)
    out = np.zeros((dst_h, dst_w, a.shape[2]), dtype=np.float32)
    for k in (3, 2, 1, 0):
        term = rows[yi[:, k], :, :] * yw[:, k][:, None, None]
        out = term if k == 3 else (out + term).astype(np.float32)
    return out[:, :, 0] if squeeze else out


def _resize_patch(patch: np.ndarray, factor: float, use_cv2: bool) -> np.ndarray:
    """pose_decode.py:157-158 ``cv2.resize(patch, None, fx=f, fy=f, INTER_CUBIC)``;
    cv2 derives ``dsize = round(src*f)`` and the source step ``1/f``."""
    if use_cv2 and cv2 is not None:
        return cv2.resize(np.ascontiguousarray(patch), None, fx=factor, fy=factor,
                          interpolation=cv2.INTER_CUBIC)
    dh = int(round(patch.shape[0] * factor))
    dw = int(round(patch.shape[1] * factor))
    return bicubic_resize_f32(patch, dh, dw, 1.0 / factor, 1.0 / factor)


def _resize_full(maps: np.ndarray, height: int, width: int, use_cv2: bool) -> np.ndarray:
    """pose_decode.py:487 ``cv2.resize(pafs, (W, H), INTER_CUBIC)``; the source
    step is ``src/dst`` per axis."""
    if use_cv2 and cv2 is not None:
        return cv2.resize(np.ascontiguousarray(maps), (width, height), interpolation=cv2.INTER_CUBIC)
    return bicubic_resize_f32(maps, height, width, maps.shape[0] / height, maps.shape[1] / width)


# ----------------------------------------------------------------------------
# D1: greedy radius NMS -- pose_decode.py:52-76 (find_peaks_v2)
# ----------------------------------------------------------------------------
def greedy_radius_peaks(channel: np.ndarray, thre1: float) -> np.ndarray:
    """Pixels above ``thre1``, visited in descending score order; a pixel is
    kept unless an earlier kept pixel lies within Euclidean distance 5.
    Returns [[x, y], ...] in kept order.

    Exact score ties: the reference orders them by ``argsort()[::-1]`` of an
    unstable sort (implementation-defined).  Here ties go to the larger
    row-major index first (= reversed stable sort), which is also the rule the
    CUDA path implements; fixtures never contain ties."""
    ys, xs = np.nonzero(channel > thre1)             # row-major, like np.where (:54)
    if ys.size == 0:
        return np.zeros((0, 2), dtype=np.int64)
    scores = channel[ys, xs]
    order = np.argsort(scores, kind="stable")[::-1]   # (:60)
    kept: List[int] = []
    alive = np.ones(order.size, dtype=bool)
    for pos, cand in enumerate(order):                # (:62-70)
        if not alive[pos]:
            continue
        kept.append(cand)
        rest = order[pos + 1:]
        d2 = (ys[rest] - ys[cand]) ** 2 + (xs[rest] - xs[cand]) ** 2
        alive[pos + 1:] &= d2 > NMS_RADIUS * NMS_RADIUS   # keeps dis > dis_thres (:68)
    kept = np.asarray(kept, dtype=np.int64)
    return np.stack([xs[kept], ys[kept]], axis=1).astype(np.int64)   # [x, y] (:72-75)


def cross_peaks(channel: np.ndarray, thre1: float) -> np.ndarray:
    """pose_decode.py:37-49 (find_peaks, unused by the reference's NMS): pixels above ``thre1`` equal to the
    maximum over the 3x3 cross footprint (scipy ``maximum_filter`` with ``generate_binary_structure(2, 1)``,
    default 'reflect' border: the mirrored neighbour of a border pixel is the pixel itself), as
    ``[[x, y], ...]`` in ``np.nonzero`` (row-major) order."""
    ch = np.asarray(channel)
    m = ch.copy()
    m[1:, :] = np.maximum(m[1:, :], ch[:-1, :])
    m[:-1, :] = np.maximum(m[:-1, :], ch[1:, :])
    m[:, 1:] = np.maximum(m[:, 1:], ch[:, :-1])
    m[:, :-1] = np.maximum(m[:, :-1], ch[:, 1:])
    ys, xs = np.nonzero((m == ch) & (ch > thre1))
    return np.stack([xs, ys], axis=1).astype(np.int64).reshape(-1, 2)


# ----------------------------------------------------------------------------
# D2: patch refinement -- pose_decode.py:107-185 (NMS) + :80-104
# ----------------------------------------------------------------------------
GAUSS_SIGMA = 3.0                                  # NMS passes sigma=3 (:163-164)
GAUSS_RADIUS = int(4.0 * GAUSS_SIGMA + 0.5)        # scipy: truncate = 4.0 -> 12


def gaussian_weights(sigma: float = GAUSS_SIGMA, radius: int = GAUSS_RADIUS) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d (order 0): float64 weights at offsets -radius..radius."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _reflect_index(i: int, n: int) -> int:
    """scipy 'reflect' border (half-sample symmetric: d c b a | a b c d | d c b a), any distance."""
    period = 2 * n
    i %= period
    return i if i < n else period - 1 - i


def gaussian_filter_f32(img: np.ndarray, sigma: float = GAUSS_SIGMA) -> np.ndarray:
    """``scipy.ndimage.gaussian_filter(img, sigma)`` for a float32 2-D array, restated from scipy's
    NI_Correlate1D (third-party, scipy >= 1.0; the reference imports it at pose_decode.py:8): axis 0 then axis 1,
    each line extended by reflection, float64 accumulation in the symmetric-kernel order
    ``tmp = x[l] * w[0]; for j = -r..-1: tmp += (x[l + j] + x[l - j]) * w[j]``, every pass rounded to float32."""
    radius = int(4.0 * sigma + 0.5)
    w = gaussian_weights(sigma, radius)
    out = np.asarray(img, dtype=np.float32)
    for axis in (0, 1):
        src = out if axis == 1 else out.T                       # lines along `axis` as rows
        n = src.shape[1]
        idx = np.array([_reflect_index(i, n) for i in range(-radius, n + radius)])
        ext = src[:, idx].astype(np.float64)                     # [lines, n + 2r]
        centre = np.arange(n) + radius
        tmp = ext[:, centre] * w[radius]
        for j in range(-radius, 0):
            tmp = tmp + (ext[:, centre + j] + ext[:, centre - j]) * w[j + radius]
        res = tmp.astype(np.float32)
        out = res if axis == 1 else res.T
    return np.ascontiguousarray(out)


def refine_peaks(heatmaps: np.ndarray, thre1: float, factor: float, use_cv2: bool = True,
                 refine: bool = True, gaussian: bool = False) -> List[np.ndarray]:
    """Per joint type an array [n, 5] float64 of ``(x, y, score, id, 0)`` at
    input-image resolution; ``id`` counts peaks over channels then peaks.
    ``refine=False`` is ``NMS(bool_refine_center=False)`` (:176-182); ``gaussian=True`` is
    ``NMS(bool_gaussian_filt=True)`` (:163-164): the up-sampled patch is smoothed before the argmax."""
    h, w = heatmaps.shape[:2]
    per_type: List[np.ndarray] = []
    next_id = 0
    for j in range(NUM_JOINTS):
        ch = heatmaps[:, :, j]
        pk = greedy_radius_peaks(ch, thre1)
        rows = np.zeros((len(pk), 5), dtype=np.float64)
        for i, (px, py) in enumerate(pk):
            if not refine:
                rows[i] = (math.floor((px + 0.5) * factor - 0.5), math.floor((py + 0.5) * factor - 0.5),
                           ch[py, px], next_id, 0)
                next_id += 1
                continue
            x0, y0 = max(0, px - PATCH_HALF), max(0, py - PATCH_HALF)          # (:150)
            x1, y1 = min(w - 1, px + PATCH_HALF), min(h - 1, py + PATCH_HALF)  # (:151-152)
            up = _resize_patch(ch[y0:y1 + 1, x0:x1 + 1], factor, use_cv2)      # (:156-158)
            if gaussian:
                up = gaussian_filter_f32(up)                                   # (:163-164)
            flat = int(np.argmax(up))                                          # first max, row-major (:167)
            my, mx = divmod(flat, up.shape[1])
            # (:171-182) floor(resized(peak) + (argmax - resized(peak - patch_origin)))
            cy = (py - y0 + 0.5) * factor - 0.5
            cx = (px - x0 + 0.5) * factor - 0.5
            rx = math.floor((px + 0.5) * factor - 0.5 + (mx - cx))
            ry = math.floor((py + 0.5) * factor - 0.5 + (my - cy))
            rows[i] = (int(rx), int(ry), up[my, mx], next_id, 0)
            next_id += 1
        per_type.append(rows)
    return per_type


# ----------------------------------------------------------------------------
# D4: limb scoring -- pose_decode.py:188-301 (find_connected_joints)
# ----------------------------------------------------------------------------
def score_limbs(paf_up: np.ndarray, per_type: List[np.ndarray], thre2: float, num_intermed_pts: int = NUM_SAMPLES,
                max_dist_thresh: float = 1) -> List[np.ndarray]:
    """Per limb type an array [m, 5] float64 ``(src_id, dst_id, score, i, j)``
    -- for every source joint its single best destination (ties: first j).  ``num_intermed_pts`` and
    ``max_dist_thresh`` are the reference's keyword arguments (:188): samples per candidate limb and the
    multiple of the frame size beyond which a pair is skipped (:209,250)."""
    H, W = paf_up.shape[:2]
    long_dist_thresh = np.array([W, H]) * max_dist_thresh          # (:209)
    out: List[np.ndarray] = []
    for t in range(NUM_LIMBS):
        src = per_type[LIMB_JOINTS[t][0]]
        dst = per_type[LIMB_JOINTS[t][1]]
        if len(src) == 0 or len(dst) == 0:                         # (:224-228)
            out.append(np.zeros((0, 5), dtype=np.float64))
            continue
        cx, cy = LIMB_PAF_CHANNELS[t]
        rows = []
        for i, a in enumerate(src):
            best, best_row = 0.0, None                             # (:243)
            for j, b in enumerate(dst):
                d = b[:2] - a[:2]                                  # (:249)
                if (np.abs(d) > long_dist_thresh).any():           # (:250)
                    continue
                norm = np.sqrt(np.sum(d ** 2)) + 1e-8              # (:255)
                u = d / norm
                xs = np.round(np.linspace(a[0], b[0], num=num_intermed_pts)).astype(np.intp)   # (:260-264)
                ys = np.round(np.linspace(a[1], b[1], num=num_intermed_pts)).astype(np.intp)
                vx = paf_up[ys, xs, cx].astype(np.float64)         # f32 samples promoted (:267-270)
                vy = paf_up[ys, xs, cy].astype(np.float64)
                s = vx * u[0] + vy * u[1]
                mean = s.mean()                                    # (:271)
                if (np.count_nonzero(s > thre2) > 0.8 * num_intermed_pts   # (:276-277)
                        and mean > thre2 and mean > best):         # (:283,285)
                    best = mean
                    best_row = (a[3], b[3], mean, i, j)            # (:287)
            if best_row is not None:
                rows.append(best_row)
        out.append(np.asarray(rows, dtype=np.float64).reshape(-1, 5))
    return out


# ----------------------------------------------------------------------------
# D5: grouping -- pose_decode.py:304-396 (group_limbs_of_same_person)
# ----------------------------------------------------------------------------
def group_persons(limbs: List[np.ndarray], joint_list: np.ndarray) -> np.ndarray:
    """[P, 16] float64: 14 joint ids (-1 = missing), total score, part count."""
    persons: List[np.ndarray] = []
    for t in range(NUM_LIMBS):
        jt_a, jt_b = LIMB_JOINTS[t]
        for a_id, b_id, s, _, _ in limbs[t]:
            hit = [k for k, p in enumerate(persons) if p[jt_a] == a_id or p[jt_b] == b_id]   # (:329-331)
            if len(hit) == 1:                                                               # (:336-347)
                p = persons[hit[0]]
                if p[jt_b] != b_id:
                    p[jt_b] = b_id
                    p[15] += 1
                    p[14] += joint_list[int(b_id), 2] + s
            elif len(hit) == 2:                                                             # (:348-366)
                p, q = persons[hit[0]], persons[hit[1]]
                if not np.any((p[:14] >= 0) & (q[:14] >= 0)):
                    p[:14] += q[:14] + 1
                    p[14:] += q[14:]
                    p[14] += s
                    persons.pop(hit[1])
                else:
                    p[jt_b] = b_id
                    p[15] += 1
                    p[14] += joint_list[int(b_id), 2] + s
            else:                                                                           # 0 or >= 3 (:367-379)
                p = -np.ones(16, dtype=np.float64)
                p[jt_a], p[jt_b] = a_id, b_id
                p[15] = 2
                p[14] = (0 + joint_list[int(a_id), 2] + joint_list[int(b_id), 2]) + s
                persons.append(p)
    persons = [p for p in persons if not (p[15] < MIN_PARTS or p[14] / p[15] < MIN_MEAN_SCORE)]  # (:382-391)
    return np.array(persons)                                                                # shape (0,) if empty (:396)


# ----------------------------------------------------------------------------
# D6: keypoint table + canvas -- pose_decode.py:399-457 (plot_pose)
# ----------------------------------------------------------------------------
def keypoint_table(joint_list: np.ndarray, persons: np.ndarray) -> np.ndarray:
    """[P, 14, 3] float32 ``(x, y, 1)`` for joints that are an endpoint of at
    least one complete limb of the person, zeros otherwise."""
    n = persons.shape[0]
    out = np.zeros((n, NUM_JOINTS, 3), dtype=np.float32)
    for k in range(n):
        for t in range(NUM_LIMBS):
            ids = persons[k][LIMB_JOINTS[t]].astype(int)
            if -1 in ids:
                continue
            for jid in ids:
                jt = int(joint_list[jid, 5])
                out[k, jt] = (joint_list[jid, 0], joint_list[jid, 1], 1)
    return out


def draw_canvas(img: np.ndarray, joint_list: np.ndarray, persons: np.ndarray) -> np.ndarray:
    canvas = img.copy()
    if cv2 is None:
        return canvas
    for k in range(persons.shape[0]):
        for t in range(NUM_LIMBS):
            ids = persons[k][LIMB_JOINTS[t]].astype(int)
            if -1 in ids:
                continue
            pts = joint_list[ids, 0:2]
            for p in pts:
                cv2.circle(canvas, tuple(p.astype(int)), 3, (255, 255, 255), thickness=-1)
            cv2.line(canvas, tuple(pts[0].astype(int)), tuple(pts[1].astype(int)),
                     color=PALETTE[k % len(PALETTE)], thickness=2)
    return canvas


# ----------------------------------------------------------------------------
# D0: decode_pose -- pose_decode.py:460-517
# ----------------------------------------------------------------------------
def decode(img: np.ndarray, param: Dict[str, float], heatmaps: np.ndarray, pafs: np.ndarray,
           use_cv2: bool = True, draw: bool = True):
    """Returns ``(canvas, joint_list, persons, joints)`` with the reference's
    dtypes and shapes (``(0,)``-shaped empties included)."""
    factor = img.shape[0] / heatmaps.shape[0]                                      # (:462)
    per_type = refine_peaks(heatmaps, param["thre1"], factor, use_cv2)             # (:471)
    joint_list = np.array([tuple(r) + (jt,) for jt, rows in enumerate(per_type) for r in rows])   # (:481-482)
    paf_up = _resize_full(pafs, img.shape[0], img.shape[1], use_cv2)               # (:487)
    limbs = score_limbs(paf_up, per_type, param["thre2"])                          # (:493)
    persons = group_persons(limbs, joint_list)                                     # (:496)
    joints = keypoint_table(joint_list, persons)                                   # (:505)
    canvas = draw_canvas(img, joint_list, persons) if draw else img
    return canvas, joint_list, persons, joints.reshape(-1, NUM_JOINTS * 3)         # (:516-517)


def decode_detail(img_shape: Tuple[int, int], param: Dict[str, float], heatmaps: np.ndarray,
                  pafs: np.ndarray, use_cv2: bool = True):
    """Same as :func:`decode` but also returns the per-stage tables (peaks per
    type, limb tables) for stage-by-stage parity checks."""
    H, W = img_shape
    factor = H / heatmaps.shape[0]
    per_type = refine_peaks(heatmaps, param["thre1"], factor, use_cv2)
    joint_list = np.array([tuple(r) + (jt,) for jt, rows in enumerate(per_type) for r in rows])
    paf_up = _resize_full(pafs, H, W, use_cv2)
    limbs = score_limbs(paf_up, per_type, param["thre2"])
    persons = group_persons(limbs, joint_list)
    joints = keypoint_table(joint_list, persons)
    return {"per_type": per_type, "joint_list": joint_list, "limbs": limbs,
            "persons": persons, "joints": joints.reshape(-1, NUM_JOINTS * 3)}
