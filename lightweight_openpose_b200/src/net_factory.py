"""Op factory mirroring the reference ``src/net_factory.py:4-37`` in eager form.

Each function takes an NHWC CUDA tensor and returns one; variables are looked up
in the bound checkpoint by the TF auto-names of the current variable scope
(``_scope``).  Kernels: ``lwop_conv2d`` / ``lwop_depthwise3x3`` (tcgen05 implicit
GEMM for stride-1 1x1/3x3 convs with Cin % 32 == 0 in bf16 mode, CUDA-core
kernels otherwise).  Only inference is implemented: batch norm uses the moving
statistics, so ``training=True`` on a BN layer raises.

Activation storage follows the engine mode (``set_mode``): bf16 tensors between
layers in 'bf16' mode, float32 in 'fp32' (check) mode; linear (``relu=False``)
convs always return float32, as the whole-net executor does for the final heads.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from .. import _lib, engine, graph
from . import _scope

_mode = "bf16"
_dev_cache: Dict[Tuple[str, int], tuple] = {}
_op_engine: Dict[int, "engine.Engine"] = {}


def set_mode(mode: str) -> None:
    """'bf16' (tensor cores), 'bf16_direct' (bf16 storage, CUDA cores) or 'fp32' (check path)."""
    global _mode
    if mode not in ("bf16", "fp32", "bf16_direct"):
        raise ValueError(mode)
    _mode = mode


def _torch():
    import torch
    return torch


def _handle(device_index: int):
    e = _op_engine.get(device_index)
    if e is None:
        e = engine.Engine(16, 16, max_batch=1, device=device_index, use_graph=False)   # op-level calls only
        _op_engine[device_index] = e
    return e


def _training_error():
    raise NotImplementedError(
        "training=True: this package implements the inference hot path only (batch norm uses the "
        "moving statistics; reference callers pass is_training=False, test&eval/model_test.py:42)")


def _device_weights(name: str, layer: "graph.Layer", dev):
    torch = _torch()
    key = (name, dev.index or 0)
    hit = _dev_cache.get(key)
    if hit is not None and hit[0] is engine.get_variables():
        return hit[1]
    f = graph.fold_layer(layer, engine.get_variables())
    t = {"w": torch.from_numpy(f.w).to(dev), "b": torch.from_numpy(f.b).to(dev),
         "dw": torch.from_numpy(f.dw).to(dev) if f.dw is not None else None}
    _dev_cache[key] = (engine.get_variables(), t)
    return t


def _act_dtype():
    torch = _torch()
    return torch.float32 if _mode == "fp32" else torch.bfloat16


def _dt(t) -> int:
    torch = _torch()
    if t.dtype == torch.float32:
        return _lib.DT_F32
    if t.dtype == torch.bfloat16:
        return _lib.DT_BF16
    raise ValueError(f"unsupported activation dtype {t.dtype}")


def _conv_call(x, wts, cout, ksize, stride, dil, act, out_dtype, residual=None):
    torch = _torch()
    x = x.contiguous()
    B, H, W, cin = x.shape
    Ho, Wo = -(-H // stride), -(-W // stride)
    out = torch.empty((B, Ho, Wo, cout), dtype=out_dtype, device=x.device)
    e = _handle(x.device.index or 0)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    direct = _lib.MODE_FP32_CHECK if _mode == "fp32" else _lib.MODE_BF16_DIRECT
    # tensor-core path where the implicit-GEMM kernel supports the shape, CUDA-core kernel otherwise
    try_tc = (_mode == "bf16" and x.dtype == torch.bfloat16 and stride == 1 and ksize in (1, 3) and cin % 32 == 0)
    for mode in ([_lib.MODE_BF16] if try_tc else []) + [direct]:
        rc = e.lib.lwop_conv2d(e.handle, mode, x.data_ptr(), _dt(x), B, H, W, cin, wts["w"].data_ptr(),
                               wts["b"].data_ptr(), cout, ksize, stride, dil, act,
                               residual.data_ptr() if residual is not None else None,
                               out.data_ptr(), _dt(out), stream)
        if rc == _lib.ERR_UNSUPPORTED and mode == _lib.MODE_BF16:
            continue
        _lib.check(rc, e.handle)
        break
    return out


def conv(inputs, out_channels, kernel_size=3, bn=True, dilation=1, stride=1, relu=True, bias=True,
         training=True, name=''):
    """conv2d 'same' (+bias) (+batch norm) (+ReLU) -- reference net_factory.py:4-12."""
    torch = _torch()
    if bn and training:
        _training_error()
    vname = _scope.next_name("conv2d")
    bn_name = _scope.next_name("batch_normalization") if bn else None
    cin = inputs.shape[-1]
    layer = graph.Layer(-1, graph.KIND_CONV, vname, bn_name, cin, out_channels, kernel_size, stride, dilation,
                        graph.ACT_RELU if relu else graph.ACT_NONE, bias)
    wts = _device_weights(vname, layer, inputs.device)
    x = inputs
    if x.dtype == torch.float32 and _mode != "fp32" and cin != 3:
        x = x.to(torch.bfloat16)
    out_dtype = torch.float32 if (not relu or _mode == "fp32") else torch.bfloat16
    return _conv_call(x, wts, out_channels, kernel_size, stride, dilation,
                      _lib.ACT_RELU if relu else _lib.ACT_NONE, out_dtype)


def _separable(inputs, out_channels, stride, dilation, bn, act, kernel_size=3):
    torch = _torch()
    vname = _scope.next_name("separable_conv2d")
    bn_name = _scope.next_name("batch_normalization") if bn else None
    cin = inputs.shape[-1]
    layer = graph.Layer(-1, graph.KIND_SEPARABLE, vname, bn_name, cin, out_channels, 1, stride, dilation, act, False,
                        dw_k=int(kernel_size))
    wts = _device_weights(vname, layer, inputs.device)
    x = inputs.contiguous()
    if x.dtype == torch.float32 and _mode != "fp32":
        x = x.to(torch.bfloat16)
    B, H, W, _ = x.shape
    Ho, Wo = -(-H // stride), -(-W // stride)
    mid = torch.empty((B, Ho, Wo, cin), dtype=x.dtype, device=x.device)
    e = _handle(x.device.index or 0)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    vec_ok = x.dtype == torch.bfloat16 and cin % 8 == 0 and _mode == "bf16"
    _lib.check(e.lib.lwop_depthwise2d(e.handle, _lib.MODE_BF16 if vec_ok else _lib.MODE_FP32_CHECK, x.data_ptr(),
                                      _dt(x), B, H, W, cin, wts["dw"].data_ptr(), int(kernel_size), stride, dilation,
                                      mid.data_ptr(), stream), e.handle)
    return _conv_call(mid, wts, out_channels, 1, 1, 1, act, _act_dtype())


def conv_dw(inputs, out_channels, kernel_size=3, stride=1, dilation=1, training=True, name=''):
    """separable conv (depthwise ``kernel_size`` -> pointwise 1x1, no bias) + batch norm + ReLU -- net_factory.py:14-21.
    3x3 (what the network uses) runs the TMA / vector depthwise kernels, other sizes a generic CUDA-core kernel."""
    if training:
        _training_error()
    return _separable(inputs, out_channels, stride, dilation, True, _lib.ACT_RELU, kernel_size)


def conv_dw_no_bn(inputs, out_channels, kernel_size=3, stride=1, dilation=1, training=True, name=''):
    """separable conv + ELU (no batch norm, so ``training`` is irrelevant) -- net_factory.py:23-31."""
    return _separable(inputs, out_channels, stride, dilation, False, _lib.ACT_ELU, kernel_size)


def add(a, b):
    """tf.add of two activations (lightweight_openpose.py:27, net_factory.py:37); float32 math."""
    torch = _torch()
    return (a.to(torch.float32) + b.to(torch.float32)).to(a.dtype)


def RefinementStageBlock(inputs, out_channels, training=True):
    """1x1 conv+ReLU, 3x3 conv+BN+ReLU, 3x3 dilation-2 conv+BN+ReLU, residual add -- net_factory.py:33-37."""
    net_initial = conv(inputs, out_channels, kernel_size=1, bn=False)
    net_truck = conv(net_initial, out_channels, training=training)
    net_truck = conv(net_truck, out_channels, dilation=2, training=training)
    return add(net_initial, net_truck)
