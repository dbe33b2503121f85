"""CPU restatement of the reference drivers' frame pre-processing -- TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Reference: ``test&eval/model_test.py:63-65`` / ``model_json.py:68-70``::

    img_data = cv2.cvtColor(img_ori, code=cv2.COLOR_BGR2RGB)
    img_data = cv2.resize(img_data, (256, 256))          # uint8, INTER_LINEAR
    img = img_data / 255.

``cv2.resize`` is third-party arithmetic (OpenCV, version unpinned by the reference; 4.13 here).  Its 8-bit
INTER_LINEAR path is fixed point: source position ``(d + 0.5) * scale - 0.5`` in float32, 11-bit coefficients
``saturate_cast<short>(f * 2048)``, horizontal pass in int32, vertical pass
``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`` (modules/imgproc/src/resize.cpp,
HResizeLinear / VResizeLinear<uchar, int, short, FixedPtCast<int, uchar, 22>>).  OpenCV builds with the IPP
HAL may route this call to IPP, whose rounding can differ by one grey level; tests/test_oracle_preprocess.py
records the agreement of this restatement with the installed cv2.
"""
from __future__ import annotations

import numpy as np


def _taps(dst: int, src: int, clamp_weights: bool):
    """Source indices and 11-bit weights of one axis.  OpenCV treats the two axes differently at the border
    (resize.cpp, `resize` coefficient loops): along x an out-of-range source column is replaced by the border column
    WITH the weight pair reset to (2048, 0) (`if (sx < 0) fx = 0, sx = 0;`); along y only the row INDICES are
    clipped (`clip(sy + k, 0, ssize.height)` in resizeGeneric_Invoker) and the fractional weight pair is kept, so a
    border row is blended with itself through two truncating products -- which can be one grey level below the
    single-product value."""
    scale = src / dst
    f = ((np.arange(dst) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    if clamp_weights:
        lo = s < 0
        s[lo], f[lo] = 0, 0.0
        hi = s >= src - 1
        s[hi], f[hi] = src - 1, 0.0
    s0 = np.clip(s, 0, src - 1)
    s1 = np.clip(s + 1, 0, src - 1)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int16).astype(np.int64)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int16).astype(np.int64)
    return s0, s1, a0, a1


def resize_linear_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [dst_h, dst_w, C], OpenCV's fixed-point INTER_LINEAR."""
    src = img.astype(np.int64)
    x0, x1, ax0, ax1 = _taps(dst_w, img.shape[1], True)
    y0, y1, ay0, ay1 = _taps(dst_h, img.shape[0], False)
    hor = src[:, x0, :] * ax0[None, :, None] + src[:, x1, :] * ax1[None, :, None]        # [H, dst_w, C] * 2^11
    top, bot = hor[y0], hor[y1]
    v = (((ay0[:, None, None] * (top >> 4)) >> 16) + ((ay1[:, None, None] * (bot >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def preprocess_bgr(frame_bgr: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """BGR uint8 frame -> RGB uint8 [dst_h, dst_w, 3] (the network then takes ``/ 255.``)."""
    return resize_linear_u8(frame_bgr[:, :, ::-1], dst_h, dst_w)
