"""Shared helpers for the tests: a plain torch-CPU fp32 conv with TF 'SAME'
padding (the per-op numerics reference) and comparison utilities."""
import numpy as np
import torch
import torch.nn.functional as F


def same_pad(size, k, stride, dil):
    out = -(-size // stride)
    total = max((out - 1) * stride + (k - 1) * dil + 1 - size, 0)
    return out, total // 2, total - total // 2


def ref_conv_nhwc(x, w_ock, bias, k, stride, dil, act, residual=None, depthwise=False, dtype=torch.float64):
    """x [B,H,W,C] numpy; w_ock [Cout][k*k][Cin] (or [9][C] for depthwise); returns numpy NHWC."""
    xt = torch.from_numpy(np.asarray(x, dtype=np.float64)).to(dtype).permute(0, 3, 1, 2)
    _, pt, pb = same_pad(xt.shape[2], k, stride, dil)
    _, pl, pr = same_pad(xt.shape[3], k, stride, dil)
    xt = F.pad(xt, (pl, pr, pt, pb))
    w = torch.from_numpy(np.asarray(w_ock, dtype=np.float64)).to(dtype)
    if depthwise:
        C = x.shape[-1]
        wt = w.reshape(3, 3, C).permute(2, 0, 1).reshape(C, 1, 3, 3)
        y = F.conv2d(xt, wt, None, stride=stride, dilation=dil, groups=C)
    else:
        cout, taps, cin = w.shape
        wt = w.reshape(cout, k, k, cin).permute(0, 3, 1, 2)
        y = F.conv2d(xt, wt, None, stride=stride, dilation=dil)
    if bias is not None:
        y = y + torch.from_numpy(np.asarray(bias, dtype=np.float64)).to(dtype).view(1, -1, 1, 1)
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = F.elu(y)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + torch.from_numpy(np.asarray(residual, dtype=np.float64)).to(dtype)
    return y.contiguous().numpy()


def bf16_round(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()
