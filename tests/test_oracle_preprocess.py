"""The pre-processing oracle (OpenCV's fixed-point 8-bit bilinear resize, restated) against the installed cv2."""
import cv2
import numpy as np
import pytest

from oracle import preprocess_oracle as pre


@pytest.mark.parametrize("ipp", [True, False])
@pytest.mark.parametrize("shape,dst", [((480, 640), (256, 256)), ((720, 1280), (368, 368)), ((368, 368), (368, 368)),
                                        ((199, 301), (368, 656)), ((1080, 1920), (368, 656)), ((100, 100), (368, 368)),
                                        ((333, 517), (368, 368)), ((64, 48), (368, 368)), ((367, 369), (368, 368)),
                                        ((2, 3), (256, 256)), ((1, 1), (16, 16)), ((2000, 37), (368, 656))])
def test_resize_restatement_vs_cv2(shape, dst, ipp):
    """Byte work: EXACT against the installed cv2, with and without IPP (up- and down-scaling, degenerate sizes)."""
    cv2.ipp.setUseIPP(ipp)
    rng = np.random.default_rng(shape[0] * 7 + dst[1])
    img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    # smooth half of the frame so that both noise and natural-image gradients are covered
    if shape[0] >= 2:
        img[: shape[0] // 2] = cv2.GaussianBlur(img[: shape[0] // 2], (0, 0), 3)
    want = cv2.resize(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), (dst[1], dst[0]))
    got = pre.preprocess_bgr(img, dst[0], dst[1])
    cv2.ipp.setUseIPP(True)
    assert np.array_equal(got, want)


def test_oracle_vs_committed_cv2_golden():
    """tests/golden/preprocess_golden.npz (made by make_preprocess_golden.py from the installed cv2, IPP off)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_golden.npz"))
    k = 0
    while f"frame{k}" in g:
        want = g[f"rgb{k}"]
        assert np.array_equal(pre.preprocess_bgr(g[f"frame{k}"], want.shape[0], want.shape[1]), want)
        k += 1
    assert k >= 5
