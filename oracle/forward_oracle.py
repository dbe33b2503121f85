"""ORACLE (test infrastructure, not product code) -- CPU restatement of the
reference network forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package
``lightweight_openpose_b200`` never does.

PARITY of the forward: the reference graph
(``/root/reference/src/lightweight_openpose.py:8-48`` built from
``/root/reference/src/net_factory.py:4-37``) needs TensorFlow 1.x, which is not
installed and cannot be installed (no network), and the reference ships no
golden activations.  This file restates that graph with torch-CPU ``F.conv2d``
under the TF semantics below.
* WIRING PINNED: ``tests/test_oracle_forward.py`` executes the UNMODIFIED reference model
  function on a stand-in for the handful of TensorFlow calls it makes (``tests/tf_shim.py``) and
  this oracle agrees with it to 1e-5 -- layer order, widths, strides, dilations, where BN /
  ReLU / ELU / bias sit, the ``tf.add`` / ``tf.concat`` edges and the creation order (= names) of
  all 173 inference variables come from the reference source itself.
* OP ARITHMETIC UNPINNED: what each TensorFlow op computes ('same' padding, separable =
  depthwise then pointwise, BN epsilon) is restated from TensorFlow's documentation in both
  the shim and this file; no TensorFlow binary was available to confirm it.
The decode half of the path is pinned bit for bit (see ``oracle/decode_oracle.py``).

TF semantics restated here
* tensors NHWC, kernels HWIO, cross-correlation;
* ``padding='same'`` (net_factory.py:6,16,26): ``out = ceil(in/stride)``,
  ``pad_total = max((out-1)*stride + (k-1)*dil + 1 - in, 0)``, the smaller half
  goes before (top/left), the remainder after;
* ``tf.layers.separable_conv2d`` = depthwise (channel multiplier 1, carries the
  stride and dilation) followed by a 1x1 pointwise conv, nothing in between
  (net_factory.py:15,25);
* ``tf.layers.batch_normalization(training=False)`` =
  ``gamma*(x-moving_mean)/sqrt(moving_variance+1e-3)+beta`` (net_factory.py:9,18);
* ``tf.nn.elu`` with alpha 1 (net_factory.py:29);
* variables are auto-named per variable scope in creation order because every
  layer is created with ``name=''`` (net_factory.py:4,14,23): ``conv2d``,
  ``conv2d_1``, ... / ``separable_conv2d`` ... / ``batch_normalization`` ...
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


def same_pad(size: int, k: int, stride: int, dil: int) -> Tuple[int, int, int]:
    """Returns (out_size, pad_before, pad_after) of TF 'SAME' padding."""
    out = -(-size // stride)
    total = max((out - 1) * stride + (k - 1) * dil + 1 - size, 0)
    before = total // 2
    return out, before, total - before


class _Scope:
    """Emulates tf.variable_scope + tf.layers auto-naming counters."""

    def __init__(self, name: str):
        self.name = name
        self.counts: Dict[str, int] = {}

    def next(self, kind: str) -> str:
        n = self.counts.get(kind, 0)
        self.counts[kind] = n + 1
        return f"{self.name}/{kind}" + (f"_{n}" if n else "")


class ForwardOracle:
    """``ForwardOracle(variables)(x_nhwc)`` -> ``(heatmaps, pafs)`` numpy NHWC.

    ``dtype`` float32 (the baseline) or float64 (bounds the oracle's own
    rounding).  ``quantize`` optionally emulates the product's bf16 storage of
    weights and inter-layer activations to predict the bf16 error budget on CPU.
    """

    def __init__(self, variables: Dict[str, np.ndarray], dtype=torch.float32,
                 quantize: bool = False, num_threads: Optional[int] = None):
        self.v = {k: torch.from_numpy(np.asarray(a)).to(dtype) for k, a in variables.items()}
        self.dtype = dtype
        self.quantize = quantize
        self.trace: Optional[Dict[str, Tuple[float, float, float]]] = None  # name -> (rms, mean, std)
        if num_threads is not None:
            torch.set_num_threads(num_threads)

    # -- helpers -------------------------------------------------------------
    def _tap(self, name: str, t: torch.Tensor) -> None:
        if self.trace is not None:
            self.trace[name] = (float(t.pow(2).mean().sqrt()), float(t.mean()), float(t.std()))

    def _tap_pre(self, name: str, t: torch.Tensor) -> None:
        """Per-channel mean/std of a layer's pre-activation (calibration tool)."""
        if self.trace is not None:
            self.trace[name + ":pre"] = (t.mean(dim=(0, 2, 3)).double().numpy(),
                                         t.std(dim=(0, 2, 3)).double().numpy())

    def _q(self, t: torch.Tensor) -> torch.Tensor:
        if not self.quantize:
            return t
        return t.to(torch.bfloat16).to(self.dtype)

    def _conv_raw(self, x, w_hwio, stride, dil, groups=1):
        kh, kw = w_hwio.shape[0], w_hwio.shape[1]
        _, pt, pb = same_pad(x.shape[2], kh, stride, dil)
        _, pl, pr = same_pad(x.shape[3], kw, stride, dil)
        x = F.pad(x, (pl, pr, pt, pb))
        if groups == 1:
            w = w_hwio.permute(3, 2, 0, 1).contiguous()          # OIHW
        else:
            w = w_hwio.permute(2, 3, 0, 1).contiguous()          # (C,1,kh,kw)
        return F.conv2d(x, w, None, stride=stride, dilation=dil, groups=groups)

    def _bn(self, x, name):
        g = self.v[name + "/gamma"]
        b = self.v[name + "/beta"]
        m = self.v[name + "/moving_mean"]
        var = self.v[name + "/moving_variance"]
        inv = g / torch.sqrt(var + BN_EPS)
        return x * inv.view(1, -1, 1, 1) + (b - m * inv).view(1, -1, 1, 1)

    # -- net_factory.py:4-12 -------------------------------------------------
    def conv(self, sc: _Scope, x, out_channels, kernel_size=3, bn=True, dilation=1,
             stride=1, relu=True, bias=True):
        name = sc.next("conv2d")
        w = self.v[name + "/kernel"]
        assert w.shape[0] == kernel_size and w.shape[3] == out_channels, name
        if self.quantize and not bn:
            w = self._q(w)
        if self.quantize and bn:
            # product folds the BN scale into the weights before rounding to bf16
            bn_name = sc.next("batch_normalization")
            g = self.v[bn_name + "/gamma"]
            inv = g / torch.sqrt(self.v[bn_name + "/moving_variance"] + BN_EPS)
            y = self._conv_raw(x, self._q(w * inv.view(1, 1, 1, -1)), stride, dilation)
            b0 = self.v[name + "/bias"] if bias else 0.0
            y = y + ((b0 - self.v[bn_name + "/moving_mean"]) * inv
                     + self.v[bn_name + "/beta"]).view(1, -1, 1, 1)
        else:
            y = self._conv_raw(x, w, stride, dilation)
            if bias:
                y = y + self.v[name + "/bias"].view(1, -1, 1, 1)
            if bn:
                y = self._bn(y, sc.next("batch_normalization"))
        self._tap_pre(name, y)
        if relu:
            y = torch.relu(y)
        self._tap(name, y)
        return y

    # -- net_factory.py:14-21 ------------------------------------------------
    def conv_dw(self, sc: _Scope, x, out_channels, kernel_size=3, stride=1, dilation=1):
        name = sc.next("separable_conv2d")
        bn_name = sc.next("batch_normalization")
        y = self._conv_raw(x, self.v[name + "/depthwise_kernel"], stride, dilation,
                           groups=x.shape[1])
        pw = self.v[name + "/pointwise_kernel"]
        if self.quantize:
            y = self._q(y)
            g = self.v[bn_name + "/gamma"]
            inv = g / torch.sqrt(self.v[bn_name + "/moving_variance"] + BN_EPS)
            y = self._conv_raw(y, self._q(pw * inv.view(1, 1, 1, -1)), 1, 1)
            y = y + (self.v[bn_name + "/beta"] - self.v[bn_name + "/moving_mean"] * inv).view(1, -1, 1, 1)
        else:
            y = self._conv_raw(y, pw, 1, 1)
            y = self._bn(y, bn_name)
        self._tap_pre(name, y)
        y = torch.relu(y)
        self._tap(name, y)
        return y

    # -- net_factory.py:23-31 ------------------------------------------------
    def conv_dw_no_bn(self, sc: _Scope, x, out_channels, kernel_size=3, stride=1, dilation=1):
        name = sc.next("separable_conv2d")
        y = self._conv_raw(x, self.v[name + "/depthwise_kernel"], stride, dilation,
                           groups=x.shape[1])
        y = self._q(y)
        y = self._conv_raw(y, self._q(self.v[name + "/pointwise_kernel"]), 1, 1)
        self._tap_pre(name, y)
        y = F.elu(y)
        self._tap(name, y)
        return y

    # -- net_factory.py:33-37 ------------------------------------------------
    def refinement_block(self, sc: _Scope, x, out_channels):
        initial = self._q(self.conv(sc, x, out_channels, kernel_size=1, bn=False))
        trunk = self._q(self.conv(sc, initial, out_channels))
        trunk = self.conv(sc, trunk, out_channels, dilation=2)
        return initial + trunk

    # -- lightweight_openpose.py:8-48 ----------------------------------------
    def __call__(self, x_nhwc: np.ndarray, num_joints: int = 14, num_pafs: int = 26,
                 return_stage1: bool = False):
        q = self._q
        x = torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(self.dtype).permute(0, 3, 1, 2)
        with torch.no_grad():
            sc = _Scope("BackBone")
            net = q(self.conv(sc, x, 32, stride=2, bias=False))
            net = q(self.conv_dw(sc, net, 64))
            net = q(self.conv_dw(sc, net, 128, stride=2))
            net = q(self.conv_dw(sc, net, 128))
            net = q(self.conv_dw(sc, net, 256, stride=2))
            net = q(self.conv_dw(sc, net, 256))
            net = q(self.conv_dw(sc, net, 512))
            net = q(self.conv_dw(sc, net, 512, dilation=2))
            for _ in range(4):
                net = q(self.conv_dw(sc, net, 512))
            sc = _Scope("Cpm")
            align = q(self.conv(sc, net, 128, kernel_size=1, bn=False))
            trunk = q(self.conv_dw_no_bn(sc, align, 128))
            trunk = q(self.conv_dw_no_bn(sc, trunk, 128))
            trunk = self.conv_dw_no_bn(sc, trunk, 128)
            cpm = q(self.conv(sc, q(align + trunk), 128, bn=False))
            sc = _Scope("InitialStage")
            t = q(self.conv(sc, cpm, 128, bn=False))
            t = q(self.conv(sc, t, 128, bn=False))
            t = q(self.conv(sc, t, 128, bn=False))
            h1 = q(self.conv(sc, t, 512, kernel_size=1, bn=False))
            h1 = self.conv(sc, h1, num_joints, kernel_size=1, bn=False, relu=False)
            p1 = q(self.conv(sc, t, 512, kernel_size=1, bn=False))
            p1 = self.conv(sc, p1, num_pafs, kernel_size=1, bn=False, relu=False)
            outs1 = q(torch.cat([h1, p1], dim=1))
            sc = _Scope("RefinementStage")
            net = outs1
            for _ in range(5):
                net = q(self.refinement_block(sc, net, 128))
            h2 = q(self.conv(sc, net, 128, kernel_size=1, bn=False))
            h2 = self.conv(sc, h2, num_joints, kernel_size=1, bn=False, relu=False)
            p2 = q(self.conv(sc, net, 128, kernel_size=1, bn=False))
            p2 = self.conv(sc, p2, num_pafs, kernel_size=1, bn=False, relu=False)

        def out(t):
            return t.permute(0, 2, 3, 1).contiguous().to(torch.float32 if self.dtype == torch.float32
                                                         else self.dtype).numpy()

        if return_stage1:
            return out(h2), out(p2), out(h1), out(p1)
        return out(h2), out(p2)
