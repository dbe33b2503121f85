import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: longer-running CPU test")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _load_golden(fname):
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", fname)
    z = np.load(path)
    scenes = {}
    for name in z["names"]:
        name = str(name)
        scenes[name] = {
            "heat": z[f"{name}/heat"], "paf": z[f"{name}/paf"],
            "hw": tuple(int(v) for v in z[f"{name}/hw"]),
            "joint_list": z[f"{name}/joint_list"],
            "limbs": [z[f"{name}/limb{t}"] for t in range(13)],
            "persons": z[f"{name}/persons"], "joints": z[f"{name}/joints"],
            "nms_gauss": z[f"{name}/nms_gauss"],
        }
    return scenes, {"thre1": float(z["thre"][0]), "thre2": float(z["thre"][1])}


@pytest.fixture(scope="session")
def decode_golden():
    """Reference module with cv2 as installed (IPP-accelerated patch resize)."""
    return _load_golden("decode_golden.npz")


@pytest.fixture(scope="session")
def decode_golden_noipp():
    """Reference module with cv2.ipp.setUseIPP(False): OpenCV's own float32 arithmetic (bit-exact target)."""
    return _load_golden("decode_golden_noipp.npz")


@pytest.fixture(scope="session")
def decode_extra_golden():
    """Reference module (IPP off) on the dense 46x82 frame and with non-default find_connected_joints keyword
    arguments (tests/golden/make_decode_extra_golden.py)."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "decode_extra_golden.npz"))
    dense = {"heat": z["dense_46x82/heat"], "paf": z["dense_46x82/paf"], "hw": tuple(int(v) for v in z["dense_46x82/hw"]),
             "joint_list": z["dense_46x82/joint_list"], "limbs": [z[f"dense_46x82/limb{t}"] for t in range(13)],
             "persons": z["dense_46x82/persons"], "joints": z["dense_46x82/joints"]}
    settings = [(int(n), float(thr)) for n, thr in z["params/settings"]]
    params = {"heat": z["params/heat"], "paf": z["params/paf"], "joint_list": z["params/joint_list"],
              "settings": settings,
              "limbs": [[z[f"params/{k}/limb{t}"] for t in range(13)] for k in range(len(settings))]}
    return dense, params, {"thre1": float(z["thre"][0]), "thre2": float(z["thre"][1])}
