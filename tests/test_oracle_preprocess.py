"""The pre-processing oracle (OpenCV's fixed-point 8-bit bilinear resize, restated) against the installed cv2."""
import cv2
import numpy as np
import pytest

from oracle import preprocess_oracle as pre


@pytest.mark.parametrize("shape,dst", [((480, 640), (256, 256)), ((720, 1280), (368, 368)), ((368, 368), (368, 368)),
                                        ((199, 301), (368, 656)), ((1080, 1920), (368, 656))])
def test_resize_restatement_vs_cv2(shape, dst):
    rng = np.random.default_rng(shape[0] * 7 + dst[1])
    img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    # smooth half of the frame so that both noise and natural-image gradients are covered
    img[: shape[0] // 2] = cv2.GaussianBlur(img[: shape[0] // 2], (0, 0), 3)
    want = cv2.resize(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), (dst[1], dst[0]))
    got = pre.preprocess_bgr(img, dst[0], dst[1])
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1, diff.max()                 # IPP-backed builds may round one grey level differently
    assert (diff == 0).mean() >= 0.97
