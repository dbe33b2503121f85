"""ctypes binding of ``_lwop.so`` (the C ABI declared in ``include/lwop.h``).

There is no CPU fallback: if the shared library is missing or a symbol the
header declares is not exported, loading raises, and every product entry point
that needs the device raises :class:`LwopError` when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
# LWOP_LIB_PATH: development A/B of differently compiled builds of the same library (tools/ab_variant.sh); same loud
# failure if the file is missing or lacks a symbol
LIB_PATH = os.environ.get("LWOP_LIB_PATH") or os.path.join(_HERE, "_lwop.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "lwop.h")

# constants mirrored from include/lwop.h
ABI_VERSION = 2
OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_WEIGHTS, ERR_OVERFLOW, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4, -5, -6
MODE_BF16, MODE_FP32_CHECK, MODE_BF16_DIRECT = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
DT_F32, DT_BF16 = 0, 1
FLAG_NO_GRAPH = 1
FLAG_NO_FUSE_SEP = 2
FLAG_NO_FUSE_HEADS = 4
PEAKS_NO_REFINE, PEAKS_CROSS, PEAKS_GAUSS = 1, 2, 4
NUM_JOINTS, NUM_PAFS, NUM_LIMBS = 14, 26, 13
MAX_PEAKS, MAX_JOINTS, MAX_PERSONS, COUNTS = 256, 3584, 256, 32


class LwopError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lwop error {code}: {message}")
        self.code = code


class DecodeTables(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("joints", C.c_void_p), ("limbs", C.c_void_p),
                ("persons", C.c_void_p), ("keypoints", C.c_void_p)]


class PackedTables(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("joints", C.c_void_p), ("limbs", C.c_void_p),
                ("persons", C.c_void_p), ("keypoints", C.c_void_p),
                ("joints_cap", C.c_size_t), ("limbs_cap", C.c_size_t), ("persons_cap", C.c_size_t)]


_vp, _i, _u, _f, _d, _sz = C.c_void_p, C.c_int, C.c_uint, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); must cover every function include/lwop.h declares
SIGNATURES: Dict[str, tuple] = {
    "lwop_abi_version": (_i, []),
    "lwop_error_string": (C.c_char_p, [_i]),
    "lwop_last_error": (C.c_char_p, [_vp]),
    "lwop_create": (_i, [_i, _i, _i, _i, _u, C.POINTER(_vp)]),
    "lwop_destroy": (_i, [_vp]),
    "lwop_weights_num_floats": (_sz, []),
    "lwop_load_weights": (_i, [_vp, _vp, _sz]),
    "lwop_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "lwop_forward_u8": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "lwop_preprocess_bgr_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "lwop_decode": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _d, C.POINTER(DecodeTables), _vp]),
    "lwop_forward_decode": (_i, [_vp, _vp, _i, _i, _i, _f, _d, _vp, _vp, C.POINTER(DecodeTables), _vp]),
    "lwop_nms": (_i, [_vp, _vp, _i, _i, _i, _d, _f, _i, _vp, _vp, _vp, _vp]),
    "lwop_find_connected_joints": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _d, _i, _d, _vp, _vp, _vp, _vp, _vp]),
    "lwop_group_limbs": (_i, [_vp, _vp, _vp, _i, C.POINTER(DecodeTables), _vp]),
    "lwop_draw_poses": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(DecodeTables), _vp, _vp]),
    "lwop_decode_fetch": (_i, [_vp, C.POINTER(DecodeTables), _i, C.POINTER(PackedTables), _vp]),
    "lwop_infer_host": (_i, [_vp, _vp, _i, _i, _i, _f, _d, C.POINTER(PackedTables), _vp, _vp, _vp]),
    "lwop_infer_host_begin": (_i, [_vp, _vp, _i, _i, _i, _f, _d, C.POINTER(PackedTables), _vp, _vp, _vp, C.POINTER(_i)]),
    "lwop_infer_host_end": (_i, [_vp, _i]),
    "lwop_conv2d": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "lwop_first_conv": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lwop_depthwise3x3": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "lwop_depthwise2d": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "lwop_separable_conv2d": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "lwop_head_pair": (_i, [_vp, _vp, C.c_longlong, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lwop_same_out": (_i, [_i, _i]),
    "lwop_crc32c": (C.c_uint32, [_vp, _sz]),
    "lwop_launch_count": (C.c_longlong, [_vp, _i]),
    "lwop_set_option": (_i, [_vp, C.c_char_p, C.c_longlong]),
    "lwop_num_steps": (_i, [_vp]),
    "lwop_step_desc": (_i, [_vp, _i, C.POINTER(_i)]),
    "lwop_profile_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, C.POINTER(_f), _vp]),
}

_lib = None


def header_functions() -> List[str]:
    """Function names declared in include/lwop.h (used by the symbol-export test)."""
    with open(HEADER_PATH, "r", encoding="utf-8") as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lwop_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load ``_lwop.so`` and bind every declared entry point.  Raises if the
    library has not been built (``python __graft_entry__.py``) or lacks a symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                          "(there is no CPU fallback for the CUDA path)")
    lib = C.CDLL(LIB_PATH)
    declared = header_functions()
    missing_sig = [n for n in declared if n not in SIGNATURES]
    if missing_sig:
        raise ImportError(f"_lib.py has no signature for {missing_sig}")
    for name in declared:
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise ImportError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype, fn.argtypes = SIGNATURES[name]
    if lib.lwop_abi_version() != ABI_VERSION:
        raise ImportError("lwop ABI version mismatch")
    _lib = lib
    return lib


def check(code: int, handle=None) -> None:
    if code == OK:
        return
    lib = load()
    msg = lib.lwop_last_error(handle) if handle else lib.lwop_last_error(None)
    msg = msg.decode("utf-8", "replace") if msg else ""
    if not msg:
        msg = lib.lwop_error_string(code).decode()
    raise LwopError(code, msg)
