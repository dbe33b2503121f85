"""``lightweight_openpose(inputs, num_joints, num_pafs, is_training)`` -- drop-in for
the reference graph builder ``src/lightweight_openpose.py:8-48``.

The reference builds a TF-1 graph and the caller fetches values with
``sess.run`` (reference ``test&eval/model_test.py:39-42,68``).  Here the call is
eager: ``inputs`` is a ``[B, H, W, 3]`` float32 RGB batch in [0, 1] (CUDA tensor,
or a numpy array which is uploaded), and the two returned maps are float32
``[B, H/8, W/8, num_joints]`` and ``[B, H/8, W/8, num_pafs]`` -- CUDA tensors for
tensor input, numpy arrays for numpy input (what ``sess.run`` hands back).

Default path: the whole-network executor behind ``lwop_forward`` (one CUDA graph
of hand-written sm_100a kernels).  ``eager=True`` instead composes the network
from the ``net_factory`` ops exactly as the reference source does, which is the
layer-by-layer restatement the parity tests compare against the executor -- and
what runs when ``num_joints`` / ``num_pafs`` differ from the checkpoint's 14 / 26.

Weights are bound once with :func:`load_weights` (replaces
``tf.train.Saver().restore``); without a call the default checkpoint
``model/model.ckpt-61236`` is used if its data shard exists, else the seeded
synthetic stand-in.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import engine
from . import _scope, net_factory
from ._scope import variable_scope
from .net_factory import conv, conv_dw, conv_dw_no_bn, RefinementStageBlock

load_weights = engine.load_weights
set_variables = engine.set_variables


def _torch():
    import torch
    return torch


def _eager_graph(inputs, num_joints, num_pafs, is_training):
    """The reference body, line for line in structure (lightweight_openpose.py:9-46)."""
    torch = _torch()
    _scope.reset()
    with variable_scope('BackBone'):
        net = conv(inputs, 32, stride=2, bias=False, training=is_training)
        net = conv_dw(net, 64, training=is_training)
        net = conv_dw(net, 128, stride=2, training=is_training)
        net = conv_dw(net, 128, training=is_training)
        net = conv_dw(net, 256, stride=2, training=is_training)
        net = conv_dw(net, 256, training=is_training)
        net = conv_dw(net, 512, training=is_training)
        net = conv_dw(net, 512, dilation=2, training=is_training)
        net = conv_dw(net, 512, training=is_training)
        net = conv_dw(net, 512, training=is_training)
        net = conv_dw(net, 512, training=is_training)
        net = conv_dw(net, 512, training=is_training)
    with variable_scope('Cpm'):
        net_align = conv(net, 128, kernel_size=1, bn=False)
        net_truck = conv_dw_no_bn(net_align, 128)
        net_truck = conv_dw_no_bn(net_truck, 128)
        net_truck = conv_dw_no_bn(net_truck, 128)
        net_cpm = conv(net_factory.add(net_align, net_truck), 128, bn=False)
    with variable_scope('InitialStage'):
        trunk_features = conv(net_cpm, 128, bn=False)
        trunk_features = conv(trunk_features, 128, bn=False)
        trunk_features = conv(trunk_features, 128, bn=False)
        heatmaps1 = conv(trunk_features, 512, kernel_size=1, bn=False)
        heatmaps1 = conv(heatmaps1, num_joints, kernel_size=1, bn=False, relu=False)
        pafs1 = conv(trunk_features, 512, kernel_size=1, bn=False)
        pafs1 = conv(pafs1, num_pafs, kernel_size=1, bn=False, relu=False)
        outs1 = torch.cat([heatmaps1, pafs1], dim=-1)
    with variable_scope('RefinementStage'):
        net = RefinementStageBlock(outs1, 128, training=is_training)
        net = RefinementStageBlock(net, 128, training=is_training)
        net = RefinementStageBlock(net, 128, training=is_training)
        net = RefinementStageBlock(net, 128, training=is_training)
        net = RefinementStageBlock(net, 128, training=is_training)
        heatmaps2 = conv(net, 128, kernel_size=1, bn=False)
        heatmaps2 = conv(heatmaps2, num_joints, kernel_size=1, bn=False, relu=False)
        pafs2 = conv(net, 128, kernel_size=1, bn=False)
        pafs2 = conv(pafs2, num_pafs, kernel_size=1, bn=False, relu=False)
    return heatmaps2, pafs2, heatmaps1, pafs1


def lightweight_openpose(inputs, num_joints, num_pafs, is_training=True, *, return_stage1=False,
                         mode: str = "bf16", eager: bool = False, device: Optional[int] = None):
    """Returns ``(heatmaps2, pafs2)`` like the reference (``:48``); with
    ``return_stage1`` also the initial-stage ``(heatmaps1, pafs1)`` that the
    reference only keeps as the internal ``outs1`` concat (``:32-36``)."""
    if is_training:
        raise NotImplementedError(
            "is_training=True: only the inference hot path is implemented (the reference's callers on this "
            "path pass is_training=False, test&eval/model_test.py:42)")
    if num_joints != 14 or num_pafs != 26:
        # the whole-network executor (and the decode chain) are built for the shipped checkpoint's 14 keypoint / 26 PAF
        # channels (reference src/train_config.py:25-26); any other head width runs the same graph composed from the
        # net_factory ops, against whatever variables of matching shapes the caller has bound
        eager = True
    torch = _torch()
    was_numpy = isinstance(inputs, np.ndarray)
    dev_index = device if device is not None else (0 if was_numpy or inputs.device.type != "cuda"
                                                  else (inputs.device.index or 0))
    engine.require_gpu(dev_index)
    x = torch.from_numpy(np.ascontiguousarray(inputs)) if was_numpy else inputs
    if x.dim() != 4 or x.shape[-1] != 3:
        raise ValueError(f"inputs must be [B, H, W, 3], got {tuple(x.shape)}")
    if x.dtype in (torch.float64, torch.float16, torch.bfloat16):
        x = x.to(torch.float32)          # the TF placeholder is float32 (model_test.py:39)
    x = x.to(torch.device("cuda", dev_index)).contiguous()
    B, H, W, _ = x.shape
    if eager:
        net_factory.set_mode(mode)
        outs = _eager_graph(x, num_joints, num_pafs, is_training)
        outs = tuple(o.to(torch.float32) for o in outs)
    else:
        eng = engine.get_engine(H, W, B, dev_index)
        outs = tuple(o.clone() for o in eng.forward(x, mode=mode, return_stage1=return_stage1))
    outs = outs if return_stage1 else outs[:2]
    if was_numpy:
        outs = tuple(o.cpu().numpy() for o in outs)
    return outs
