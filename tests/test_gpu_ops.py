"""Per-kernel parity on the GPU through the C ABI: every CUDA conv kernel against
a plain torch-CPU reference of the same op (float64 math on the same bf16-rounded
inputs), over the distinct (Cin, Cout, k, stride, dil, H, W) of the network
 including odd sizes."""
import ctypes as C

import numpy as np
import pytest
import torch

from lightweight_openpose_b200 import _lib, engine
from util import bf16_round, ref_conv_nhwc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return engine.Engine(16, 16, max_batch=1, device=0, use_graph=False)


def _conv(eng, mode, x, w, b, k, stride, dil, act, res, out_dtype):
    B, H, W, cin = x.shape
    cout = w.shape[0]
    Ho, Wo = -(-H // stride), -(-W // stride)
    out = torch.empty((B, Ho, Wo, cout), dtype=out_dtype, device="cuda")
    dt = lambda t: _lib.DT_F32 if t.dtype == torch.float32 else _lib.DT_BF16
    rc = eng.lib.lwop_conv2d(eng.handle, mode, x.data_ptr(), dt(x), B, H, W, cin, w.data_ptr(),
                             b.data_ptr() if b is not None else None, cout, k, stride, dil, act,
                             res.data_ptr() if res is not None else None, out.data_ptr(), dt(out),
                             torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
    torch.cuda.synchronize()
    return out


FP32_CASES = [
    # B, H, W, cin, cout, k, stride, dil, act, residual
    (2, 37, 29, 3, 32, 3, 2, 1, 1, False),      # first conv, odd size
    (1, 23, 23, 128, 128, 3, 1, 2, 1, True),    # refinement dil-2 conv + residual
    (1, 16, 19, 40, 128, 1, 1, 1, 1, False),
    (2, 9, 9, 512, 14, 1, 1, 1, 0, False),
    (1, 12, 10, 128, 128, 1, 1, 1, 2, True),    # ELU + add
]


@pytest.mark.parametrize("case", FP32_CASES)
def test_conv_fp32_check_kernel(eng, case):
    B, H, W, cin, cout, k, stride, dil, act, use_res = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    x = rng.normal(size=(B, H, W, cin)).astype(np.float32)
    w = (rng.normal(size=(cout, k * k, cin)) / np.sqrt(k * k * cin)).astype(np.float32)
    b = rng.normal(size=(cout,)).astype(np.float32)
    Ho, Wo = -(-H // stride), -(-W // stride)
    res = rng.normal(size=(B, Ho, Wo, cout)).astype(np.float32) if use_res else None
    got = _conv(eng, _lib.MODE_FP32_CHECK, torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(),
                torch.from_numpy(b).cuda(), k, stride, dil, act,
                torch.from_numpy(res).cuda() if use_res else None, torch.float32).cpu().numpy()
    want = ref_conv_nhwc(x, w, b, k, stride, dil, act, res)
    assert np.abs(got - want).max() <= 2e-5


@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("shape", [(2, 368, 368), (1, 37, 29), (3, 64, 50), (1, 368, 656), (2, 9, 7)])
def test_first_conv_production_kernel(eng, shape, u8):
    """conv1_bf16_kernel<float | uint8> (3 -> 32, stride 2, 'same' = pad 0 before / 1 after for even sizes) alone,
    against float64 torch-CPU on the same inputs; output is bf16, so the bound is its rounding."""
    B, H, W = shape
    rng = np.random.default_rng(H * 31 + W + int(u8))
    w = (rng.normal(size=(32, 27)) / np.sqrt(27.0)).astype(np.float32)
    b = rng.normal(size=(32,)).astype(np.float32) * 0.1
    if u8:
        frames = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
        x = (frames / 255.).astype(np.float32)            # the kernel's LUT is float32(i / 255.)
        xin = torch.from_numpy(frames).cuda()
    else:
        x = rng.random((B, H, W, 3), dtype=np.float32)
        xin = torch.from_numpy(x).cuda()
    Ho, Wo = -(-H // 2), -(-W // 2)
    out = torch.empty((B, Ho, Wo, 32), dtype=torch.bfloat16, device="cuda")
    rc = eng.lib.lwop_first_conv(eng.handle, xin.data_ptr(), int(u8), B, H, W, w.ctypes.data_as(C.c_void_p),
                                 b.ctypes.data_as(C.c_void_p), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
    torch.cuda.synchronize()
    want = ref_conv_nhwc(x, w.reshape(32, 9, 3), b, 3, 2, 1, 1, None)
    got = out.float().cpu().numpy()
    err = np.abs(got - want)
    tol = 2.0 ** -8 * max(1.0, float(np.abs(want).max())) + 1e-5
    assert got.shape == want.shape and err.max() <= tol, (err.max(), tol)


TC_CASES = [
    # B, H, W, cin, cout, k, dil, act, residual, out_f32
    (1, 16, 16, 64, 128, 1, 1, 1, False, False),     # one full M tile pair, K = 64
    (2, 23, 23, 128, 128, 1, 1, 1, False, False),    # M tail
    (3, 46, 46, 512, 512, 1, 1, 1, False, False),    # largest GEMM, two N tiles of 256
    (2, 46, 46, 256, 256, 1, 1, 1, False, False),
    (1, 92, 92, 32, 64, 1, 1, 1, False, False),      # K = 32 (64B swizzle)
    (2, 23, 31, 128, 256, 1, 1, 1, False, False),
    (2, 46, 46, 512, 128, 1, 1, 1, False, False),
    (2, 23, 23, 128, 128, 1, 1, 2, True, False),     # ELU + residual
    (2, 23, 23, 512, 14, 1, 1, 0, False, True),      # narrow linear head, fp32 out
    (2, 23, 23, 128, 26, 1, 1, 0, False, True),
    (2, 23, 23, 64, 128, 1, 1, 1, False, False),
    (1, 16, 16, 128, 128, 3, 1, 1, False, False),    # 3x3 im2col
    (2, 46, 46, 128, 128, 3, 1, 1, False, False),
    (2, 46, 46, 128, 128, 3, 2, 1, True, False),     # dilation 2 + residual
    (3, 23, 37, 128, 128, 3, 2, 1, False, False),    # non-square, image-boundary-crossing tiles
    (1, 7, 5, 128, 128, 3, 1, 1, False, False),      # tiny tensor (< 128 KiB im2col map workaround)
    (40, 46, 46, 128, 128, 3, 1, 1, True, False),    # CTA-pair GEMM: several 512-row pair tiles per cluster, both TMEM stages
    (40, 46, 46, 512, 128, 1, 1, 1, False, False),   # CTA-pair GEMM, 1x1, deep K
    (5, 46, 82, 128, 128, 3, 2, 1, True, False),     # halo-patch 3x3: wide map (656x368 frames), dilation 2, residual
    (4, 17, 9, 128, 128, 3, 1, 2, False, False),     # halo-patch 3x3: 2 x 2 tiles per image, both clipped, ELU
    (7, 46, 46, 128, 128, 3, 1, 0, False, False),    # halo-patch 3x3: tile count not a multiple of 4 (126), linear
]


@pytest.mark.parametrize("case", [c for c in TC_CASES if c[5] == 3 and c[0] * c[1] * c[2] >= 512])
def test_conv3x3_im2col_pair_kernel_still_works(eng, case, monkeypatch):
    """LWOP_HALO=0 selects the im2col-mode TMA form of the dense 3x3 (the A/B switch of the halo-patch kernel)."""
    monkeypatch.setenv("LWOP_HALO", "0")
    test_conv_tcgen05_kernel(eng, case)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_kernel(eng, case):
    B, H, W, cin, cout, k, dil, act, use_res, out_f32 = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    x = bf16_round(rng.normal(size=(B, H, W, cin)))
    w = bf16_round(rng.normal(size=(cout, k * k, cin)) / np.sqrt(k * k * cin))
    b = rng.normal(size=(cout,)).astype(np.float32)
    res = bf16_round(rng.normal(size=(B, H, W, cout))) if use_res else None
    got = _conv(eng, _lib.MODE_BF16, torch.from_numpy(x).cuda().to(torch.bfloat16), torch.from_numpy(w).cuda(),
                torch.from_numpy(b).cuda(), k, 1, dil, act,
                torch.from_numpy(res).cuda().to(torch.bfloat16) if use_res else None,
                torch.float32 if out_f32 else torch.bfloat16).float().cpu().numpy()
    want = ref_conv_nhwc(x, w, b, k, 1, dil, act, res)
    tol = 2e-4 if out_f32 else 2.0 ** -8 * max(1.0, float(np.abs(want).max()))   # bf16 output rounding
    err = np.abs(got - want)
    assert err.max() <= tol, (err.max(), tol, np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("case", [(2, 23, 23, 128, 128, 3, 2, 1, True), (1, 20, 17, 64, 128, 1, 1, 2, False)])
def test_conv_bf16_direct_kernel(eng, case):
    B, H, W, cin, cout, k, dil, act, use_res = case
    rng = np.random.default_rng(7)
    x = bf16_round(rng.normal(size=(B, H, W, cin)))
    w = (rng.normal(size=(cout, k * k, cin)) / np.sqrt(k * k * cin)).astype(np.float32)
    b = rng.normal(size=(cout,)).astype(np.float32)
    res = bf16_round(rng.normal(size=(B, H, W, cout))) if use_res else None
    got = _conv(eng, _lib.MODE_BF16_DIRECT, torch.from_numpy(x).cuda().to(torch.bfloat16), torch.from_numpy(w).cuda(),
                torch.from_numpy(b).cuda(), k, 1, dil, act,
                torch.from_numpy(res).cuda().to(torch.bfloat16) if use_res else None,
                torch.bfloat16).float().cpu().numpy()
    want = ref_conv_nhwc(x, w, b, k, 1, dil, act, res)
    assert np.abs(got - want).max() <= 2.0 ** -8 * max(1.0, float(np.abs(want).max()))


DW_CASES = [
    # B, H, W, C, stride, dil
    (2, 184, 184, 32, 1, 1), (1, 184, 184, 64, 2, 1), (1, 92, 92, 128, 1, 1), (2, 92, 92, 128, 2, 1),
    (1, 46, 46, 256, 1, 1), (2, 46, 46, 512, 1, 2), (1, 45, 37, 128, 2, 1), (1, 23, 41, 512, 1, 2),
    (1, 33, 33, 64, 1, 1),
]


@pytest.mark.parametrize("case", DW_CASES)
@pytest.mark.parametrize("kind", ["tma", "vector", "fp32"])
def test_depthwise_kernels(eng, case, kind):
    B, H, W, Cc, stride, dil = case
    rng = np.random.default_rng(11)
    x = bf16_round(rng.normal(size=(B, H, W, Cc)))
    w = (rng.normal(size=(9, Cc)) / 3.0).astype(np.float32)
    Ho, Wo = -(-H // stride), -(-W // stride)
    xd = torch.from_numpy(x).cuda()
    vec = kind != "fp32"
    if vec:
        xd = xd.to(torch.bfloat16)
    out = torch.empty((B, Ho, Wo, Cc), dtype=xd.dtype, device="cuda")
    mode = {"tma": _lib.MODE_BF16, "vector": _lib.MODE_BF16_DIRECT, "fp32": _lib.MODE_FP32_CHECK}[kind]
    wd = torch.from_numpy(w).cuda()
    rc = eng.lib.lwop_depthwise3x3(eng.handle, mode, xd.data_ptr(),
                                   _lib.DT_BF16 if vec else _lib.DT_F32, B, H, W, Cc, wd.data_ptr(),
                                   stride, dil, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
    torch.cuda.synchronize()
    want = ref_conv_nhwc(x, w, None, 3, stride, dil, 0, depthwise=True)
    tol = 2.0 ** -8 * max(1.0, float(np.abs(want).max())) if vec else 1e-5
    assert np.abs(out.float().cpu().numpy() - want).max() <= tol


SEP_CASES = [
    # B, H, W, cin, cout, dil, act, residual
    (2, 184, 184, 32, 64, 1, 1, False),      # backbone layer 1: K = 32 (64B swizzle), one K chunk per tile
    (2, 92, 92, 128, 128, 1, 1, False),      # layer 3
    (2, 46, 46, 256, 256, 1, 1, False),      # layer 5
    (2, 46, 46, 256, 512, 1, 1, False),      # layer 6: two N tiles
    (3, 46, 46, 512, 512, 2, 1, False),      # layer 7: dilation 2
    (2, 46, 46, 512, 512, 1, 1, False),      # layers 8-11
    (2, 46, 46, 128, 128, 1, 2, True),       # Cpm: ELU + tf.add
    (1, 23, 37, 128, 128, 1, 2, False),      # odd, non-square, patches clipped at the border
    (1, 7, 5, 64, 256, 2, 1, False),         # map smaller than one patch
    (5, 16, 32, 128, 256, 1, 0, False),      # exact patches, linear
    (1, 23, 37, 512, 512, 1, 1, False),      # CTA-pair kernel: 9 tiles (odd: the last pair holds a dummy patch), clipped patches
    (3, 8, 16, 256, 512, 1, 1, False),       # CTA-pair kernel: 3 exact tiles, one K group
    (1, 7, 5, 512, 512, 2, 1, False),        # a single tile: falls back to the single-CTA Cout = 512 form
    (2, 46, 46, 128, 128, 1, 2, False),      # ELU without tf.add: 64-column epilogue
]


@pytest.mark.parametrize("case", SEP_CASES)
def test_fused_separable_kernel(eng, case):
    """conv_dw / conv_dw_no_bn as ONE kernel (depthwise on CUDA cores feeding the tcgen05 GEMM):
    reference = float64 depthwise, rounded to bf16 (the operand the tensor core sees), float64 pointwise."""
    B, H, W, cin, cout, dil, act, use_res = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    x = bf16_round(rng.normal(size=(B, H, W, cin)))
    dw = (rng.normal(size=(9, cin)) / 3.0).astype(np.float32)
    pw = bf16_round(rng.normal(size=(cout, 1, cin)) / np.sqrt(cin))
    b = rng.normal(size=(cout,)).astype(np.float32)
    res = bf16_round(rng.normal(size=(B, H, W, cout))) if use_res else None
    xd = torch.from_numpy(x).cuda().to(torch.bfloat16)
    out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device="cuda")
    dwd, pwd, bd = torch.from_numpy(dw).cuda(), torch.from_numpy(pw).cuda(), torch.from_numpy(b).cuda()
    resd = torch.from_numpy(res).cuda().to(torch.bfloat16) if use_res else None
    rc = eng.lib.lwop_separable_conv2d(eng.handle, _lib.MODE_BF16, xd.data_ptr(), B, H, W, cin, dwd.data_ptr(),
                                       pwd.data_ptr(), bd.data_ptr(), cout, dil, act,
                                       resd.data_ptr() if use_res else None, out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
    torch.cuda.synchronize()
    mid = ref_conv_nhwc(x, dw, None, 3, 1, dil, 0, depthwise=True)
    mid_lo = bf16_round(mid)
    want = ref_conv_nhwc(mid_lo, pw, b, 1, 1, 1, act, res)
    # the kernel's fp32 depthwise may round a mid value to the neighbouring bf16: allow a few such flips per row
    slack = ref_conv_nhwc(np.abs(mid) * 2.0 ** -8, np.abs(pw), None, 1, 1, 1, 0).max() * 0.25
    tol = 2.0 ** -8 * max(1.0, float(np.abs(want).max())) + slack
    err = np.abs(out.float().cpu().numpy() - want)
    assert err.max() <= tol, (err.max(), tol, np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("case", [(2 * 46 * 46, 512), (3 * 23 * 37, 128), (128, 512), (5, 128), (2 * 46 * 46 + 77, 128),
                                  # several tiles per CTA: the cross-tile MMA schedule, both acc2 stages, ragged last tile
                                  (148 * 3 * 128 + 50, 512), (148 * 5 * 128, 128)])
def test_fused_head_pair_kernel(eng, case):
    """Two 2-layer 1x1 heads on one input as ONE back-to-back GEMM kernel (initial-stage heads 128->512->14|26,
    refinement heads 128->128->14|26): float64 reference with the hidden activations rounded to bf16, which is
    what the second tensor-core GEMM consumes."""
    M, hidden = case
    rng = np.random.default_rng(M * 7 + hidden)
    x = bf16_round(rng.normal(size=(M, 128)))
    w1 = [bf16_round(rng.normal(size=(hidden, 128)) / np.sqrt(128)) for _ in range(2)]
    b1 = [rng.normal(size=(hidden,)).astype(np.float32) * 0.5 for _ in range(2)]
    w2 = [bf16_round(rng.normal(size=(n, hidden)) / np.sqrt(hidden)) for n in (14, 26)]
    b2 = [rng.normal(size=(n,)).astype(np.float32) for n in (14, 26)]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    xd = dev(x).to(torch.bfloat16)
    d = [dev(a) for a in (w1[0], b1[0], w2[0], b2[0], w1[1], b1[1], w2[1], b2[1])]
    oa = torch.full((M, 14), float("nan"), dtype=torch.float32, device="cuda")
    ob = torch.full((M, 26), float("nan"), dtype=torch.float32, device="cuda")
    rc = eng.lib.lwop_head_pair(eng.handle, xd.data_ptr(), M, hidden, *[t.data_ptr() for t in d], oa.data_ptr(),
                                ob.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, eng.handle)
    torch.cuda.synchronize()
    for got, wa, ba, wb, bb in ((oa, w1[0], b1[0], w2[0], b2[0]), (ob, w1[1], b1[1], w2[1], b2[1])):
        hid = np.maximum(x.astype(np.float64) @ wa.astype(np.float64).T + ba, 0.0)
        want = bf16_round(hid).astype(np.float64) @ wb.astype(np.float64).T + bb
        slack = (np.abs(hid) * 2.0 ** -8) @ np.abs(wb.astype(np.float64)).T * 0.25    # a few bf16 rounding flips per row
        err = np.abs(got.cpu().numpy() - want)
        assert (err <= 2e-4 + slack).all(), (err.max(), np.unravel_index(err.argmax(), err.shape))


def test_fused_separable_with_residual_is_bitwise_repeatable(eng):
    """Regression: the depthwise producers used to hand their input slab back to the TMA producer right after the FMA
    loop; nothing ordered that barrier arrive behind shared-memory loads still in flight, and in the layer with the
    slowest epilogue (ELU + tf.add, lightweight_openpose.py:26-27) a refill could land first: single pixels changed from
    run to run.  60 launches on the same tensors must agree bit for bit."""
    B, H, W, cin, cout = 64, 46, 46, 128, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((B, H, W, cin), device="cuda", generator=g).to(torch.bfloat16)
    res = torch.randn((B, H, W, cout), device="cuda", generator=g).to(torch.bfloat16)
    dw = torch.randn((9, cin), device="cuda", generator=g) / 3
    pw = torch.randn((cout, cin), device="cuda", generator=g) / cin ** 0.5
    b = torch.randn((cout,), device="cuda", generator=g)
    ref = None
    for run in range(60):
        out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device="cuda")
        rc = eng.lib.lwop_separable_conv2d(eng.handle, _lib.MODE_BF16, x.data_ptr(), B, H, W, cin, dw.data_ptr(),
                                           pw.data_ptr(), b.data_ptr(), cout, 1, 2, res.data_ptr(), out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, eng.handle)
        torch.cuda.synchronize()
        if ref is None:
            ref = out
        else:
            assert torch.equal(out, ref), f"launch {run} differs from launch 0 in {int((out != ref).sum())} elements"


@pytest.mark.parametrize("ksize,stride,dil,cin,cout", [(5, 1, 1, 32, 48), (5, 2, 1, 24, 64), (1, 1, 1, 64, 32), (7, 1, 2, 16, 16),
                                                        (2, 2, 1, 32, 32)])
def test_conv_dw_with_other_kernel_sizes(ksize, stride, dil, cin, cout):
    """conv_dw / conv_dw_no_bn take kernel_size (reference net_factory.py:14,23); the network only uses 3, other sizes run the
    generic depthwise kernel + the 1x1 GEMM.  Against an fp64 torch-CPU separable conv with TensorFlow's 'same' padding."""
    import torch.nn.functional as F
    from lightweight_openpose_b200 import engine as eng_mod
    from lightweight_openpose_b200.src import _scope, net_factory
    from util import same_pad
    rng = np.random.default_rng(ksize * 100 + stride * 10 + dil)
    B, H, W = 2, 23, 30
    x = rng.standard_normal((B, H, W, cin)).astype(np.float32)
    v = {"T/separable_conv2d/depthwise_kernel": rng.standard_normal((ksize, ksize, cin, 1)).astype(np.float32) / ksize,
         "T/separable_conv2d/pointwise_kernel": (rng.standard_normal((1, 1, cin, cout)) / np.sqrt(cin)).astype(np.float32),
         "T/batch_normalization/gamma": rng.uniform(0.5, 1.5, cout).astype(np.float32),
         "T/batch_normalization/beta": rng.normal(0, 0.1, cout).astype(np.float32),
         "T/batch_normalization/moving_mean": rng.normal(0, 0.1, cout).astype(np.float32),
         "T/batch_normalization/moving_variance": rng.uniform(0.5, 1.5, cout).astype(np.float32)}
    # conv_dw_no_bn gets a scope of its own: a variable name belongs to ONE layer (the folded weights are cached by name)
    v["U/separable_conv2d/depthwise_kernel"] = v["T/separable_conv2d/depthwise_kernel"]
    v["U/separable_conv2d/pointwise_kernel"] = v["T/separable_conv2d/pointwise_kernel"]

    def reference(bn):
        xt = torch.from_numpy(x).double().permute(0, 3, 1, 2)
        _, pt, pb = same_pad(H, ksize, stride, dil)
        _, pl, pr = same_pad(W, ksize, stride, dil)
        xt = F.pad(xt, (pl, pr, pt, pb))
        dw = torch.from_numpy(v["T/separable_conv2d/depthwise_kernel"]).double().permute(2, 3, 0, 1)        # [C,1,k,k]
        y = F.conv2d(xt, dw, None, stride=stride, dilation=dil, groups=cin)
        pw = torch.from_numpy(v["T/separable_conv2d/pointwise_kernel"]).double().permute(3, 2, 0, 1)
        y = F.conv2d(y, pw)
        if bn:
            g, b_, m, var = (torch.from_numpy(v["T/batch_normalization/" + k]).double().view(1, -1, 1, 1)
                             for k in ("gamma", "beta", "moving_mean", "moving_variance"))
            y = torch.relu(g * (y - m) / torch.sqrt(var + 1e-3) + b_)
        else:
            y = F.elu(y)
        return y.permute(0, 2, 3, 1).numpy()

    saved = eng_mod.get_variables() if getattr(eng_mod, "_variables", None) is not None else None
    eng_mod.set_variables(v, "test")
    try:
        xd = torch.from_numpy(x).cuda()
        for mode, tol in (("fp32", 2e-5), ("bf16", 6e-2)):
            net_factory.set_mode(mode)
            for bn in (True, False):
                _scope.reset()
                with _scope.variable_scope("T" if bn else "U"):
                    y = (net_factory.conv_dw(xd, cout, kernel_size=ksize, stride=stride, dilation=dil, training=False) if bn
                         else net_factory.conv_dw_no_bn(xd, cout, kernel_size=ksize, stride=stride, dilation=dil))
                want = reference(bn)
                got = y.float().cpu().numpy()
                assert got.shape == want.shape, (got.shape, want.shape)
                scale = max(1.0, float(np.abs(want).max()))
                assert float(np.abs(got - want).max()) <= tol * scale, (mode, bn, float(np.abs(got - want).max()), scale)
    finally:
        net_factory.set_mode("bf16")
        if saved is not None:
            eng_mod.set_variables(saved, "restored")
