"""Generates tests/golden/preprocess_golden.npz: frames and what the reference drivers' pre-processing makes of them,

    img = cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB), (W, H))        # test&eval/model_test.py:63-64

computed by the installed OpenCV with IPP switched off (cv2.ipp.setUseIPP(False): OpenCV's own resize engine).  On
this build the IPP-on result is byte-identical (asserted below).  Run from the repo root:
    python tests/golden/make_preprocess_golden.py
"""
import os

import cv2
import numpy as np

CASES = [((37, 53), (64, 80)), ((120, 90), (48, 64)), ((96, 128), (96, 128)), ((23, 200), (40, 56)), ((150, 11), (64, 64))]


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    for k, (src, dst) in enumerate(CASES):
        rng = np.random.default_rng(100 + k)
        frame = rng.integers(0, 256, size=src + (3,), dtype=np.uint8)
        frame[: src[0] // 2] = cv2.GaussianBlur(frame[: src[0] // 2], (0, 0), 2)
        cv2.ipp.setUseIPP(False)
        want = cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB), (dst[1], dst[0]))
        cv2.ipp.setUseIPP(True)
        assert np.array_equal(want, cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB), (dst[1], dst[0])))
        out[f"frame{k}"] = frame
        out[f"rgb{k}"] = want
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "preprocess_golden.npz"), **out)


if __name__ == "__main__":
    main()
