"""A minimal stand-in for the TensorFlow 1.x calls the reference model code makes (test infrastructure only).

TensorFlow is not installable here, so the reference forward cannot be run as shipped.  This shim implements exactly
the API surface ``/root/reference/src/lightweight_openpose.py`` and ``src/net_factory.py`` touch --
``tf.variable_scope(name, reuse=tf.AUTO_REUSE)``, ``tf.layers.conv2d / separable_conv2d / batch_normalization``,
``tf.nn.relu / elu``, ``tf.add``, ``tf.concat`` -- on torch-CPU tensors, so that the UNMODIFIED reference model function
can be executed: the layer order, channel counts, strides / dilations, where BN / ReLU / ELU / bias sit and how the
residual and concat edges are wired then come from the reference source itself, not from a reading of it.  The
per-op arithmetic (TF 'same' padding, separable = depthwise then pointwise, BN eps 1e-3, tf.layers auto-naming of
variables per scope) is restated here, as it is in oracle/forward_oracle.py.

Use:  ``with installed(variables): from src.lightweight_openpose import lightweight_openpose`` (tests/test_oracle_forward.py).
"""
from __future__ import annotations

import contextlib
import sys
import types
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

AUTO_REUSE = object()
_BN_EPS = 1e-3          # tf.layers.batch_normalization default epsilon


class _State:
    def __init__(self, variables: Dict[str, np.ndarray]):
        self.variables = variables
        self.scopes: List[str] = []
        self.counters: Dict[str, Dict[str, int]] = {}
        self.used: List[str] = []

    def unique(self, kind: str) -> str:
        """tf.layers auto-naming: `kind`, `kind_1`, ... per variable scope, in creation order (layers get name='')."""
        scope = "/".join(self.scopes)
        c = self.counters.setdefault(scope, {})
        n = c.get(kind, 0)
        c[kind] = n + 1
        return (scope + "/" if scope else "") + kind + (f"_{n}" if n else "")

    def var(self, name: str) -> torch.Tensor:
        self.used.append(name)
        return torch.from_numpy(np.ascontiguousarray(self.variables[name], dtype=np.float32))


_state: _State = None


def _same_pad(size, k, stride, dil):
    out = -(-size // stride)
    total = max((out - 1) * stride + (k - 1) * dil + 1 - size, 0)
    return total // 2, total - total // 2


def _conv_nhwc(x, w_hwio, stride, dil, groups=1):
    kh, kw = int(w_hwio.shape[0]), int(w_hwio.shape[1])
    pt, pb = _same_pad(x.shape[1], kh, stride, dil)
    pl, pr = _same_pad(x.shape[2], kw, stride, dil)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    return F.conv2d(xn, w_hwio.permute(3, 2, 0, 1).contiguous(), stride=stride, dilation=dil, groups=groups).permute(0, 2, 3, 1)


def _conv2d(inputs, filters, kernel_size, strides=1, dilation_rate=1, padding="valid", use_bias=True, name=None):
    assert padding == "same" and not name
    base = _state.unique("conv2d")
    w = _state.var(base + "/kernel")
    assert tuple(w.shape) == (kernel_size, kernel_size, inputs.shape[-1], filters), (base, tuple(w.shape))
    y = _conv_nhwc(inputs, w, strides, dilation_rate)
    if use_bias:
        y = y + _state.var(base + "/bias")
    return y


def _separable_conv2d(inputs, filters, kernel_size, strides=1, dilation_rate=1, padding="valid", use_bias=True, name=None):
    assert padding == "same" and not name and not use_bias
    base = _state.unique("separable_conv2d")
    dw = _state.var(base + "/depthwise_kernel")           # [k, k, Cin, 1]
    pw = _state.var(base + "/pointwise_kernel")           # [1, 1, Cin, filters]
    cin = inputs.shape[-1]
    assert tuple(dw.shape) == (kernel_size, kernel_size, cin, 1) and tuple(pw.shape) == (1, 1, cin, filters), base
    y = _conv_nhwc(inputs, dw.permute(0, 1, 3, 2), strides, dilation_rate, groups=cin)     # HWIO with I = 1 per group
    return _conv_nhwc(y, pw, 1, 1)


def _batch_normalization(inputs, training=False):
    assert training is False
    base = _state.unique("batch_normalization")
    g, b = _state.var(base + "/gamma"), _state.var(base + "/beta")
    m, v = _state.var(base + "/moving_mean"), _state.var(base + "/moving_variance")
    return g * (inputs - m) / torch.sqrt(v + _BN_EPS) + b


@contextlib.contextmanager
def _variable_scope(name, reuse=None):
    _state.scopes.append(name)
    try:
        yield
    finally:
        _state.scopes.pop()


def _module() -> types.ModuleType:
    tf = types.ModuleType("tensorflow")
    tf.AUTO_REUSE = AUTO_REUSE
    tf.variable_scope = _variable_scope
    tf.add = lambda a, b: a + b
    tf.concat = lambda values, axis: torch.cat(list(values), dim=axis)
    tf.layers = types.SimpleNamespace(conv2d=_conv2d, separable_conv2d=_separable_conv2d,
                                      batch_normalization=_batch_normalization)
    tf.nn = types.SimpleNamespace(relu=torch.relu, elu=lambda x: F.elu(x, alpha=1.0))
    return tf


@contextlib.contextmanager
def installed(variables: Dict[str, np.ndarray], reference_root: str = "/root/reference"):
    """Installs the shim as ``tensorflow`` and puts the reference tree on sys.path for the duration of the block."""
    global _state
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "src", "src.net_factory", "src.lightweight_openpose")}
    _state = _State(variables)
    sys.modules["tensorflow"] = _module()
    for k in ("src", "src.net_factory", "src.lightweight_openpose"):
        sys.modules.pop(k, None)
    sys.path.insert(0, reference_root)
    try:
        yield _state
    finally:
        sys.path.remove(reference_root)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        _state = None
