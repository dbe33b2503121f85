"""Canonical layer table of the lightweight_openpose inference graph and the
host-side weight folding / packing that feeds the C-ABI (``lwop_load_weights``).

The table restates, in creation order, the ops built by the reference graph
(reference ``src/lightweight_openpose.py:8-48`` through
``src/net_factory.py:4-37``).  Because every ``tf.layers`` call in the reference
passes ``name=''``, TensorFlow auto-names variables per variable scope in
creation order (``conv2d``, ``conv2d_1``, ... / ``separable_conv2d`` ... /
``batch_normalization`` ...); :func:`build_layers` reproduces those names so the
checkpoint tensors can be bound without TensorFlow.

The C++ executor (``csrc/lwop_net.cu``) hard-codes the same 43-entry table; the
packed blob produced here must match its expected size byte for byte, which is
how the two tables are kept honest.

Folding (inference only, BN epsilon 1e-3 = tf.layers default):
``w' = w * gamma/sqrt(var+eps)``, ``b' = (b - mean) * gamma/sqrt(var+eps) + beta``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

BN_EPS = 1e-3

ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
KIND_CONV, KIND_SEPARABLE = 0, 1


@dataclass(frozen=True)
class Layer:
    index: int
    kind: int              # KIND_CONV | KIND_SEPARABLE (depthwise 3x3 -> pointwise 1x1)
    name: str              # TF variable scope of the conv / separable conv
    bn: Optional[str]      # TF name of the batch_normalization that follows, or None
    cin: int
    cout: int
    k: int                 # kernel size of the dense conv (1 for the pointwise half)
    stride: int
    dil: int
    act: int
    bias: bool
    note: str = ""
    dw_k: int = 3          # kernel size of the depthwise half of a separable layer (the network only uses 3)


class _Scope:
    def __init__(self, name: str):
        self.name = name
        self.counts: Dict[str, int] = {}

    def next(self, kind: str) -> str:
        n = self.counts.get(kind, 0)
        self.counts[kind] = n + 1
        return f"{self.name}/{kind}" + (f"_{n}" if n else "")


def build_layers(num_joints: int = 14, num_pafs: int = 26) -> List[Layer]:
    layers: List[Layer] = []

    def conv(sc, cin, cout, k=3, bn=True, dil=1, stride=1, relu=True, bias=True, note=""):
        name = sc.next("conv2d")
        bn_name = sc.next("batch_normalization") if bn else None
        layers.append(Layer(len(layers), KIND_CONV, name, bn_name, cin, cout, k, stride, dil,
                            ACT_RELU if relu else ACT_NONE, bias, note))

    def conv_dw(sc, cin, cout, stride=1, dil=1):
        name = sc.next("separable_conv2d")
        bn_name = sc.next("batch_normalization")
        layers.append(Layer(len(layers), KIND_SEPARABLE, name, bn_name, cin, cout, 1, stride, dil,
                            ACT_RELU, False))

    def conv_dw_no_bn(sc, cin, cout, note=""):
        name = sc.next("separable_conv2d")
        layers.append(Layer(len(layers), KIND_SEPARABLE, name, None, cin, cout, 1, 1, 1,
                            ACT_ELU, False, note))

    sc = _Scope("BackBone")                                   # lightweight_openpose.py:9-21
    conv(sc, 3, 32, stride=2, bias=False)
    conv_dw(sc, 32, 64)
    conv_dw(sc, 64, 128, stride=2)
    conv_dw(sc, 128, 128)
    conv_dw(sc, 128, 256, stride=2)
    conv_dw(sc, 256, 256)
    conv_dw(sc, 256, 512)
    conv_dw(sc, 512, 512, dil=2)
    for _ in range(4):
        conv_dw(sc, 512, 512)
    sc = _Scope("Cpm")                                        # :22-27
    conv(sc, 512, 128, k=1, bn=False, note="align")
    conv_dw_no_bn(sc, 128, 128)
    conv_dw_no_bn(sc, 128, 128)
    conv_dw_no_bn(sc, 128, 128, note="+align")
    conv(sc, 128, 128, bn=False)
    sc = _Scope("InitialStage")                               # :28-36
    for _ in range(3):
        conv(sc, 128, 128, bn=False)
    conv(sc, 128, 512, k=1, bn=False)
    conv(sc, 512, num_joints, k=1, bn=False, relu=False, note="heatmaps1")
    conv(sc, 128, 512, k=1, bn=False)
    conv(sc, 512, num_pafs, k=1, bn=False, relu=False, note="pafs1")
    sc = _Scope("RefinementStage")                            # :37-46, net_factory.py:33-37
    cin = num_joints + num_pafs
    for _ in range(5):
        conv(sc, cin, 128, k=1, bn=False, note="initial")
        conv(sc, 128, 128)
        conv(sc, 128, 128, dil=2, note="+initial")
        cin = 128
    conv(sc, 128, 128, k=1, bn=False)
    conv(sc, 128, num_joints, k=1, bn=False, relu=False, note="heatmaps2")
    conv(sc, 128, 128, k=1, bn=False)
    conv(sc, 128, num_pafs, k=1, bn=False, relu=False, note="pafs2")
    return layers


LAYERS: List[Layer] = build_layers()


def variable_names(layer: Layer) -> List[str]:
    names = []
    if layer.kind == KIND_SEPARABLE:
        names += [layer.name + "/depthwise_kernel", layer.name + "/pointwise_kernel"]
    else:
        names.append(layer.name + "/kernel")
        if layer.bias:
            names.append(layer.name + "/bias")
    if layer.bn:
        names += [layer.bn + "/" + s for s in ("gamma", "beta", "moving_mean", "moving_variance")]
    return names


@dataclass
class FoldedLayer:
    dw: Optional[np.ndarray]   # [9, cin] fp32 (tap-major) for separable layers
    w: np.ndarray              # [cout, k*k, cin] fp32, BN scale folded in
    b: np.ndarray              # [cout] fp32, BN shift folded in


def fold_layer(layer: Layer, v: Dict[str, np.ndarray]) -> FoldedLayer:
    if layer.kind == KIND_SEPARABLE:
        dwk = np.asarray(v[layer.name + "/depthwise_kernel"], dtype=np.float64)   # [3,3,cin,1]
        assert dwk.shape == (layer.dw_k, layer.dw_k, layer.cin, 1), (layer.name, dwk.shape)
        dw = dwk.reshape(layer.dw_k * layer.dw_k, layer.cin)
        w = np.asarray(v[layer.name + "/pointwise_kernel"], dtype=np.float64)     # [1,1,cin,cout]
    else:
        dw = None
        w = np.asarray(v[layer.name + "/kernel"], dtype=np.float64)               # [k,k,cin,cout]
    assert w.shape == (layer.k, layer.k, layer.cin, layer.cout), (layer.name, w.shape)
    b = (np.asarray(v[layer.name + "/bias"], dtype=np.float64) if layer.bias
         else np.zeros(layer.cout, dtype=np.float64))
    if layer.bn:
        g = np.asarray(v[layer.bn + "/gamma"], dtype=np.float64)
        beta = np.asarray(v[layer.bn + "/beta"], dtype=np.float64)
        mean = np.asarray(v[layer.bn + "/moving_mean"], dtype=np.float64)
        var = np.asarray(v[layer.bn + "/moving_variance"], dtype=np.float64)
        inv = g / np.sqrt(var + BN_EPS)
        w = w * inv.reshape(1, 1, 1, -1)
        b = (b - mean) * inv + beta
    # HWIO -> [O][tap][I]
    w = np.ascontiguousarray(w.reshape(layer.k * layer.k, layer.cin, layer.cout).transpose(2, 0, 1))
    return FoldedLayer(None if dw is None else dw.astype(np.float32),
                       w.astype(np.float32), b.astype(np.float32))


def fits_executor(v: Dict[str, np.ndarray]) -> bool:
    """Whether a variable set has the shapes of the layer table the whole-network executor is built for (the shipped
    checkpoint: 14 keypoint and 26 PAF channels).  Other head widths run through the op-level graph only."""
    for layer in LAYERS:
        key = layer.name + ("/pointwise_kernel" if layer.kind == KIND_SEPARABLE else "/kernel")
        w = v.get(key)
        if w is None or tuple(np.shape(w))[2:] != (layer.cin, layer.cout):
            return False
    return True


def fold_all(v: Dict[str, np.ndarray], layers: Optional[List[Layer]] = None) -> List[FoldedLayer]:
    return [fold_layer(l, v) for l in (layers or LAYERS)]


def pack_blob(folded: List[FoldedLayer]) -> np.ndarray:
    """Concatenate, per layer in table order, ``dw`` (if any), ``w`` and ``b``
    as one float32 vector -- the layout ``lwop_load_weights`` expects."""
    parts = []
    for f in folded:
        if f.dw is not None:
            parts.append(f.dw.ravel())
        parts.append(f.w.ravel())
        parts.append(f.b.ravel())
    return np.ascontiguousarray(np.concatenate(parts).astype(np.float32))


def blob_floats(layers: Optional[List[Layer]] = None) -> int:
    n = 0
    for l in (layers or LAYERS):
        if l.kind == KIND_SEPARABLE:
            n += 9 * l.cin
        n += l.cout * l.k * l.k * l.cin + l.cout
    return n


def output_hw(h: int, w: int) -> Tuple[int, int]:
    """Stride-8 output size under three SAME stride-2 layers."""
    for _ in range(3):
        h, w = -(-h // 2), -(-w // 2)
    return h, w
