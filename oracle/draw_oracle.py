"""CPU restatement of the drawing half of ``plot_pose`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Reference: ``src/pose_decode.py:431-436``: per person and limb type with both joints set, in this order,

    cv2.circle(canvas, joint0, 3, (255, 255, 255), thickness=-1); cv2.circle(canvas, joint1, 3, ...)
    cv2.line(canvas, joint0, joint1, color=colors[person % 16], thickness=2)

``cv2.circle`` / ``cv2.line`` are third-party arithmetic (OpenCV, version unpinned by the reference; 4.13 here).  Restated
from OpenCV's ``modules/imgproc/src/drawing.cpp``:

* filled circle of radius r: the midpoint-circle span fill of ``Circle()``; for the two radii that occur (3 for joints,
  1 for the end caps of a 2-pixel line) it is a fixed stencil, clipped to the image;
* ``ThickLine`` for thickness 2, LINE_8: corners ``p +- (round(dy * r), round(dx * r))`` in 16.16 fixed point with
  ``r = thickness * 32768 / |p1 - p0|`` (round-half-even), ``FillConvexPoly`` of the 4 corners (first the outline with the
  fixed-point DDA ``Line2`` after ``clipLine``, then the scan-line fill with its edge stepping
  ``dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y))``), then a radius-1 filled circle at both end points.

``tests/test_oracle_draw.py`` checks this against the installed cv2 on random segments (interior and border).
"""
from __future__ import annotations

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT

# pixels (dx, dy) of cv2's filled circle, radius 3 and radius 1 (what Circle(img, c, r, color, fill=1) covers)
CIRCLE3 = [(dx, dy) for dy, half in zip(range(-3, 4), (0, 2, 2, 3, 2, 2, 0)) for dx in range(-half, half + 1)]
CIRCLE1 = [(0, -1), (-1, 0), (0, 0), (1, 0), (0, 1)]


def _idiv(a: int, b: int) -> int:
    """C++ integer division (truncation towards zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def _cv_round(v: float) -> int:
    return int(np.rint(v))


def _circle(plot, cx, cy, stencil):
    for dx, dy in stencil:
        plot(cx + dx, cy + dy)


def _clip_line(w: int, h: int, p1, p2):
    """cv::clipLine on int64 points (here in 16.16 coordinates: w, h already scaled)."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = w - 1, h - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def _line2(plot, w, h, p1, p2):
    """Line2(): fixed-point DDA between two 16.16 points."""
    ok, (x1, y1), (x2, y2) = _clip_line(w << XY_SHIFT, h << XY_SHIFT, p1, p2)
    if not ok:
        return
    dx, dy = x2 - x1, y2 - y1
    ax, ay = abs(dx), abs(dy)
    if ax > ay:
        if dx < 0:
            dy = -dy
            x1, x2 = x2, x1
            y1, y2 = y2, y1
        x_step = XY_ONE
        y_step = _idiv(dy << XY_SHIFT, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        if dy < 0:
            dx = -dx
            x1, x2 = x2, x1
            y1, y2 = y2, y1
        x_step = _idiv(dx << XY_SHIFT, ay | 1)
        y_step = XY_ONE
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += XY_ONE >> 1
    y1 += XY_ONE >> 1
    # the end point is put first, rounded on its own (ICV_PUT_POINT((pt2 + half) >> XY_SHIFT)), then the DDA runs
    xe, ye = (x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT
    if 0 <= xe < w and 0 <= ye < h:
        plot(xe, ye)
    for _ in range(ecount + 1):
        x, y = x1 >> XY_SHIFT, y1 >> XY_SHIFT
        if 0 <= x < w and 0 <= y < h:
            plot(x, y)
        x1 += x_step
        y1 += y_step


def _fill_convex_poly(plot, w, h, v):
    """FillConvexPoly(img, v, npts, color, LINE_8, shift = XY_SHIFT) on 16.16 points."""
    npts = len(v)
    delta = XY_ONE >> 1
    p0 = v[-1]
    xmin = xmax = v[0][0]
    ymin = ymax = v[0][1]
    imin = 0
    for i, p in enumerate(v):
        if p[1] < ymin:
            ymin, imin = p[1], i
        ymax = max(ymax, p[1])
        xmax = max(xmax, p[0])
        xmin = min(xmin, p[0])
        _line2(plot, w, h, p0, p)
        p0 = p
    xmin = (xmin + delta) >> XY_SHIFT
    xmax = (xmax + delta) >> XY_SHIFT
    ymin = (ymin + delta) >> XY_SHIFT
    ymax = (ymax + delta) >> XY_SHIFT
    if npts < 3 or xmax < 0 or ymax < 0 or xmin >= w or ymin >= h:
        return
    ymax = min(ymax, h - 1)
    edge = [dict(idx=imin, di=1, x=-XY_ONE, dx=0, ye=ymin), dict(idx=imin, di=npts - 1, x=-XY_ONE, dx=0, ye=ymin)]
    edges = npts
    y = ymin
    while True:
        for e in edge:
            if y >= e["ye"]:
                idx0, di = e["idx"], e["di"]
                idx = idx0 + di
                if idx >= npts:
                    idx -= npts
                while True:
                    edges -= 1
                    if edges < 0:          # `for (; edges-- > 0; )` ran out
                        break
                    ty = (v[idx][1] + delta) >> XY_SHIFT
                    if ty > y:
                        xs, xe = v[idx0][0], v[idx][0]
                        e["ye"] = ty
                        e["dx"] = _idiv((xe - xs) * 2 + (ty - y), 2 * (ty - y))
                        e["x"] = xs
                        e["idx"] = idx
                        break
                    idx0 = idx
                    idx += di
                    if idx >= npts:
                        idx -= npts
        if edges < 0:
            break
        if y >= 0:
            left, right = (1, 0) if edge[0]["x"] > edge[1]["x"] else (0, 1)
            xx1 = (edge[left]["x"] + (XY_ONE >> 1)) >> XY_SHIFT
            xx2 = (edge[right]["x"] + (XY_ONE >> 1)) >> XY_SHIFT
            if xx2 >= 0 and xx1 < w:
                xx1 = max(xx1, 0)
                xx2 = min(xx2, w - 1)
                for x in range(xx1, xx2 + 1):
                    plot(x, y)
        edge[0]["x"] += edge[0]["dx"]
        edge[1]["x"] += edge[1]["dx"]
        y += 1
        if y > ymax:
            break


def thick_line2(plot, w, h, p0, p1):
    """cv2.line(img, p0, p1, color, thickness=2) (LINE_8, shift 0) for integer end points.

    ``cv::line`` first clips the segment to the image rectangle grown by ``thickness`` on every side (probed against
    cv2 4.13: margins 0, 1, 3, 4, 8 all disagree, 2 agrees on every case) and draws nothing if it misses it; the
    thick line and its end caps are then built from the CLIPPED end points."""
    ok, q0, q1 = _clip_line(w + 4, h + 4, (p0[0] + 2, p0[1] + 2), (p1[0] + 2, p1[1] + 2))
    if not ok:
        return
    p0, p1 = (q0[0] - 2, q0[1] - 2), (q1[0] - 2, q1[1] - 2)
    inv = 1.0 / XY_ONE
    x0, y0 = p0[0] << XY_SHIFT, p0[1] << XY_SHIFT
    x1, y1 = p1[0] << XY_SHIFT, p1[1] << XY_SHIFT
    dx, dy = (x0 - x1) * inv, (y1 - y0) * inv
    r = dx * dx + dy * dy
    thickness = 2 << (XY_SHIFT - 1)
    if abs(r) > np.finfo(np.float64).eps:
        r = thickness / np.sqrt(r)
        dpx, dpy = _cv_round(dy * r), _cv_round(dx * r)
        pts = [(x0 + dpx, y0 + dpy), (x0 - dpx, y0 - dpy), (x1 - dpx, y1 - dpy), (x1 + dpx, y1 + dpy)]
        _fill_convex_poly(plot, w, h, pts)
    for cx, cy in ((x0, y0), (x1, y1)):
        _circle(plot, (cx + (XY_ONE >> 1)) >> XY_SHIFT, (cy + (XY_ONE >> 1)) >> XY_SHIFT, CIRCLE1)


def draw_limb(canvas: np.ndarray, j0, j1, color) -> None:
    """The three drawing calls of one limb (pose_decode.py:431-436) on ``canvas`` (uint8 [H, W, 3]), in place."""
    h, w = canvas.shape[:2]

    def plotter(col):
        def plot(x, y):
            if 0 <= x < w and 0 <= y < h:
                canvas[y, x] = col
        return plot
    _circle(plotter((255, 255, 255)), int(j0[0]), int(j0[1]), CIRCLE3)
    _circle(plotter((255, 255, 255)), int(j1[0]), int(j1[1]), CIRCLE3)
    thick_line2(plotter(color), w, h, (int(j0[0]), int(j0[1])), (int(j1[0]), int(j1[1])))


def plot_pose_canvas(img_orig: np.ndarray, joint_list: np.ndarray, persons: np.ndarray, limb_table, colors) -> np.ndarray:
    """Canvas of ``plot_pose`` (pose_decode.py:399-436): persons in order, limb types in order."""
    canvas = img_orig.copy()
    persons = np.asarray(persons, dtype=np.float64).reshape(-1, 16) if np.size(persons) else np.zeros((0, 16))
    for person, info in enumerate(persons):
        for a, b in limb_table:
            ia, ib = int(info[a]), int(info[b])
            if ia == -1 or ib == -1:
                continue
            ja, jb = joint_list[ia, 0:2].astype(int), joint_list[ib, 0:2].astype(int)
            draw_limb(canvas, ja, jb, tuple(int(c) for c in colors[person % len(colors)]))
    return canvas
