// CUDA-core kernels: the fp32 check path (LWOP_MODE_FP32_CHECK), the first
// 3x3 stride-2 conv (3 -> 32), the depthwise 3x3 kernels and the uint8 -> float
// frame normalisation.  Reference ops: src/net_factory.py:4-31.
#include "lwop_kernels.cuh"

namespace lwop {

// ---------------------------------------------------------------------------
// Generic dense conv as a tiled implicit GEMM on CUDA cores (any k/stride/dil):
// M = B*Ho*Wo output pixels, N = Cout, K = taps*Cin.  64x64 tile, 16-deep
// slices, 4x4 outputs per thread, fp32 accumulation in a fixed k order.
// ---------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const TIn *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
                   const TOut *__restrict__ residual, TOut *__restrict__ out, int B, int H, int W, int Cin,
                   int Ho, int Wo, int Cout, int ksize, int stride, int dil, int pad_t, int pad_l, int act,
                   int ldi, int ldo, int out_col0, int ldr) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long M = (long long)B * Ho * Wo;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int taps = ksize * ksize;

    // each thread loads 4 A elements (row lm, 4 consecutive k) and 4 W elements
    const int lm = tid >> 2;          // 0..63
    const int lk = (tid & 3) * 4;     // 0,4,8,12
    long long gm = m0 + lm;
    int ob = 0, oy = 0, ox = 0;
    bool row_ok = gm < M;
    if (row_ok) {
        ox = (int)(gm % Wo);
        long long t = gm / Wo;
        oy = (int)(t % Ho);
        ob = (int)(t / Ho);
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int tap = 0; tap < taps; ++tap) {
        const int dy = tap / ksize, dx = tap % ksize;
        const int iy = oy * stride - pad_t + dy * dil;
        const int ix = ox * stride - pad_l + dx * dil;
        const bool pix_ok = row_ok && iy >= 0 && iy < H && ix >= 0 && ix < W;
        const TIn *ip = in + (((long long)ob * H + iy) * W + ix) * ldi;
        for (int c0 = 0; c0 < Cin; c0 += BK) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int c = c0 + lk + q;
                As[lk + q][lm] = (pix_ok && c < Cin) ? ld_as_float(ip + c) : 0.0f;
                int n = n0 + lm;
                Ws[lk + q][lm] = (n < Cout && c < Cin) ? w[((long long)n * taps + tap) * Cin + c] : 0.0f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= Cout) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.0f);
            v = apply_act(v, act);
            if (residual) v += ld_as_float(residual + m * ldr + n);
            st_from_float(out + m * ldo + out_col0 + n, v);
        }
    }
}

template <typename TIn, typename TOut>
static cudaError_t launch_conv_direct_t(const ConvArgs &a, cudaStream_t s) {
    SamePad ph = same_pad(a.H, a.ksize, a.stride, a.dil), pw = same_pad(a.W, a.ksize, a.stride, a.dil);
    long long M = (long long)a.B * ph.out * pw.out;
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((a.Cout + 63) / 64));
    conv_direct_kernel<TIn, TOut><<<grid, 256, 0, s>>>(
        (const TIn *)a.in, a.w, a.bias, (const TOut *)a.residual, (TOut *)a.out, a.B, a.H, a.W, a.Cin, ph.out,
        pw.out, a.Cout, a.ksize, a.stride, a.dil, ph.before, pw.before, a.act, a.ldi ? a.ldi : a.Cin,
        a.ldo ? a.ldo : a.Cout, a.out_col0, a.ldr ? a.ldr : a.Cout);
    return cudaGetLastError();
}

cudaError_t launch_conv_direct(const ConvArgs &a, cudaStream_t s) {
    if (a.in_dtype == LWOP_DT_F32 && a.out_dtype == LWOP_DT_F32) return launch_conv_direct_t<float, float>(a, s);
    if (a.in_dtype == LWOP_DT_F32 && a.out_dtype == LWOP_DT_BF16)
        return launch_conv_direct_t<float, __nv_bfloat16>(a, s);
    if (a.in_dtype == LWOP_DT_BF16 && a.out_dtype == LWOP_DT_BF16)
        return launch_conv_direct_t<__nv_bfloat16, __nv_bfloat16>(a, s);
    if (a.in_dtype == LWOP_DT_BF16 && a.out_dtype == LWOP_DT_F32)
        return launch_conv_direct_t<__nv_bfloat16, float>(a, s);
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------
// Depthwise 3x3, generic element type, one thread per (pixel, channel): the
// fp32 check path and the simple reference for the vectorised bf16 kernel.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
depthwise_simple_kernel(const T *__restrict__ in, const float *__restrict__ w, T *__restrict__ out, int B, int H,
                        int W, int C, int Ho, int Wo, int ksize, int stride, int dil, int pad_t, int pad_l) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C;
    if (idx >= total) return;
    int c = (int)(idx % C);
    long long p = idx / C;
    int ox = (int)(p % Wo);
    p /= Wo;
    int oy = (int)(p % Ho);
    int b = (int)(p / Ho);
    float acc = 0.0f;
    for (int dy = 0; dy < ksize; ++dy) {
        int iy = oy * stride - pad_t + dy * dil;
        if (iy < 0 || iy >= H) continue;
        for (int dx = 0; dx < ksize; ++dx) {
            int ix = ox * stride - pad_l + dx * dil;
            if (ix < 0 || ix >= W) continue;
            acc = fmaf(ld_as_float(in + (((long long)b * H + iy) * W + ix) * C + c), w[(dy * ksize + dx) * C + c], acc);
        }
    }
    st_from_float(out + idx, acc);
}

cudaError_t launch_depthwise_simple(const DwArgs &a, cudaStream_t s) {
    SamePad ph = same_pad(a.H, a.ksize, a.stride, a.dil), pw = same_pad(a.W, a.ksize, a.stride, a.dil);
    long long total = (long long)a.B * ph.out * pw.out * a.C;
    unsigned grid = (unsigned)((total + 255) / 256);
    if (a.dtype == LWOP_DT_F32)
        depthwise_simple_kernel<float><<<grid, 256, 0, s>>>((const float *)a.in, a.w, (float *)a.out, a.B, a.H, a.W,
                                                            a.C, ph.out, pw.out, a.ksize, a.stride, a.dil,
                                                            ph.before, pw.before);
    else
        depthwise_simple_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(
            (const __nv_bfloat16 *)a.in, a.w, (__nv_bfloat16 *)a.out, a.B, a.H, a.W, a.C, ph.out, pw.out, a.ksize,
            a.stride, a.dil, ph.before, pw.before);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Depthwise 3x3 bf16 NHWC, bandwidth-bound: each thread owns 8 channels (one
// 128-bit vector) of a strip of output pixels along x; the 9 taps x 8 channels
// of weights sit in registers; fp32 accumulation; 128-bit loads and stores.
// Rows of the input strip are re-used from registers across the x sweep
// (3-column sliding window), so every input vector is loaded from L1/L2 at most
// 3 times (once per output row that needs it) and from DRAM once.
// ---------------------------------------------------------------------------
template <int STRIDE, int DIL>
__global__ void __launch_bounds__(256)
depthwise_bf16_kernel(const uint4 *__restrict__ in, const float *__restrict__ w, uint4 *__restrict__ out, int B,
                      int H, int W, int C8, int Ho, int Wo, int pad_t, int pad_l, int strip) {
    // thread -> (channel group cg, strip sx, row oy, image b); blockIdx.x = work item * cblocks + channel block
    const int cblocks = (C8 + blockDim.x - 1) / blockDim.x;
    const int cg = (blockIdx.x % cblocks) * blockDim.x + threadIdx.x;      // fastest: channel groups (coalesced)
    if (cg >= C8) return;
    const int nstrips = (Wo + strip - 1) / strip;
    int t = blockIdx.x / cblocks;
    const int sx = t % nstrips;
    t /= nstrips;
    const int oy = t % Ho;
    const int b = t / Ho;
    const int C = C8 * 8;

    float wt[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 a = *reinterpret_cast<const float4 *>(w + k * C + cg * 8);
        const float4 c = *reinterpret_cast<const float4 *>(w + k * C + cg * 8 + 4);
        wt[k][0] = a.x; wt[k][1] = a.y; wt[k][2] = a.z; wt[k][3] = a.w;
        wt[k][4] = c.x; wt[k][5] = c.y; wt[k][6] = c.z; wt[k][7] = c.w;
    }
    const int x_begin = sx * strip;
    const int x_end = min(Wo, x_begin + strip);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int ox = x_begin; ox < x_end; ++ox) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = oy * STRIDE - pad_t + dy * DIL;
            const bool yok = iy >= 0 && iy < H;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * STRIDE - pad_l + dx * DIL;
                uint4 v = zero;
                if (yok && ix >= 0 && ix < W) v = __ldg(in + (((long long)b * H + iy) * W + ix) * C8 + cg);
                const int k = dy * 3 + dx;
                acc[0] = fmaf(bf16_lo(v.x), wt[k][0], acc[0]);
                acc[1] = fmaf(bf16_hi(v.x), wt[k][1], acc[1]);
                acc[2] = fmaf(bf16_lo(v.y), wt[k][2], acc[2]);
                acc[3] = fmaf(bf16_hi(v.y), wt[k][3], acc[3]);
                acc[4] = fmaf(bf16_lo(v.z), wt[k][4], acc[4]);
                acc[5] = fmaf(bf16_hi(v.z), wt[k][5], acc[5]);
                acc[6] = fmaf(bf16_lo(v.w), wt[k][6], acc[6]);
                acc[7] = fmaf(bf16_hi(v.w), wt[k][7], acc[7]);
            }
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]);
        o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]);
        o.w = pack_bf16x2(acc[6], acc[7]);
        out[(((long long)b * Ho + oy) * Wo + ox) * C8 + cg] = o;
    }
}

cudaError_t launch_depthwise_bf16(const DwArgs &a, cudaStream_t s) {
    if (a.C % 8 != 0) return cudaErrorInvalidValue;
    SamePad ph = same_pad(a.H, 3, a.stride, a.dil), pw = same_pad(a.W, 3, a.stride, a.dil);
    const int C8 = a.C / 8;
    const int threads = C8 >= 64 ? 64 : 32;
    const int strip = 8;
    const int nstrips = (pw.out + strip - 1) / strip;
    const long long nblocks = (long long)((C8 + threads - 1) / threads) * a.B * ph.out * nstrips;
    if (nblocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    dim3 grid((unsigned)nblocks);
#define LWOP_DW(S, D)                                                                                          \
    depthwise_bf16_kernel<S, D><<<grid, threads, 0, s>>>((const uint4 *)a.in, a.w, (uint4 *)a.out, a.B, a.H,    \
                                                         a.W, C8, ph.out, pw.out, ph.before, pw.before, strip)
    if (a.stride == 1 && a.dil == 1) LWOP_DW(1, 1);
    else if (a.stride == 2 && a.dil == 1) LWOP_DW(2, 1);
    else if (a.stride == 1 && a.dil == 2) LWOP_DW(1, 2);
    else return cudaErrorInvalidValue;
#undef LWOP_DW
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// First layer: 3x3 stride-2 conv 3 -> 32 + folded BN + ReLU (net_factory.py:4-12
// via lightweight_openpose.py:10).  fp32 NHWC frame in, bf16 NHWC out.  One
// thread per output pixel, all 32 output channels in registers, the 27x32
// weights broadcast from shared memory; the output row (64 B) is written with
// four 128-bit stores.
// ---------------------------------------------------------------------------
// Folded weights [27][32] (k = tap*3 + c major) and bias of the first conv travel as a by-value kernel parameter
// (Conv1Weights, 3.5 KB): kernel parameters live in the constant bank, so with the 27 x 32 loop fully unrolled
// every FFMA takes its weight straight from there (no shared-memory or register traffic for the 864 taps) -- and,
// unlike a __constant__ symbol, they belong to the handle that launches: two handles with different checkpoints on
// one device do not share layer 0.
// uint8 -> float32(u8 / 255.0) table (the reference's `img / 255.` in float64, fed as float32), filled by the host;
// the same 256 values for every handle
__device__ float g_u8_lut[256];

cudaError_t conv1_pack_weights(const float *w_host_n27 /*[32][27]*/, const float *b_host /*[32]*/, Conv1Weights *out) {
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 27; ++k) out->w[k * 32 + n] = w_host_n27[n * 27 + k];
    for (int n = 0; n < 32; ++n) out->b[n] = b_host[n];
    float lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = (float)((double)i / 255.0);
    return cudaMemcpyToSymbol(g_u8_lut, lut, sizeof(lut));
}

// Two horizontally adjacent output pixels per thread: every 128-bit weight load from the constant bank feeds both,
// and the 864 multiply-adds of a pixel issue as 432 packed FFMA2 (scalar input broadcast x uniform-register weight
// pair, accumulator pair = two adjacent output channels).  Each output keeps its 27-step accumulation order.
template <typename TIn>
__global__ void __launch_bounds__(128)
conv1_bf16_kernel(const __grid_constant__ Conv1Weights cw, const TIn *__restrict__ in, uint4 *__restrict__ out, int B, int H,
                  int W, int Ho, int Wo, int pad_t, int pad_l) {
    __shared__ float lut[256];
    if (sizeof(TIn) == 1) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = g_u8_lut[i];
        __syncthreads();
    }
    pdl_wait();                  // first kernel of the forward chain: the previous step's readers of `out` are done
    pdl_launch_dependents();
    const int wp = (Wo + 1) >> 1;                   // pixel pairs per output row
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * wp;
    if (p >= total) return;
    const int ox = (int)(p % wp) * 2;
    long long t = p / wp;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const bool two = ox + 1 < Wo;
    float2 acc[2][16];
#pragma unroll
    for (int n = 0; n < 16; ++n) acc[0][n] = acc[1][n] = make_float2(cw.b[2 * n], cw.b[2 * n + 1]);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int iy = oy * 2 - pad_t + dy;
        const bool yok = iy >= 0 && iy < H;
        const TIn *rowp = in + ((long long)b * H + (yok ? iy : 0)) * W * 3;
        // the two windows share one input column: 5 columns x 3 channels per row
        float v[5][3];
#pragma unroll
        for (int cx = 0; cx < 5; ++cx) {
            const int ix = ox * 2 - pad_l + cx;
            const bool ok = yok && ix >= 0 && ix < W;
            const TIn *ip = rowp + (ok ? ix : 0) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float x = 0.0f;
                if (ok) x = sizeof(TIn) == 1 ? lut[(int)ip[c]] : (float)ip[c];
                v[cx][c] = x;
            }
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 v0 = make_float2(v[dx][c], v[dx][c]), v1 = make_float2(v[dx + 2][c], v[dx + 2][c]);
                const float *wr = cw.w + ((dy * 3 + dx) * 3 + c) * 32;
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                    const float2 w2 = make_float2(wr[2 * n], wr[2 * n + 1]);
                    acc[0][n] = ffma2(v0, w2, acc[0][n]);
                    acc[1][n] = ffma2(v1, w2, acc[1][n]);
                }
            }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
        if (px == 1 && !two) break;
        uint4 *op = out + (((long long)b * Ho + oy) * Wo + ox + px) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(fmaxf(acc[px][q * 4 + 0].x, 0.f), fmaxf(acc[px][q * 4 + 0].y, 0.f));
            o.y = pack_bf16x2(fmaxf(acc[px][q * 4 + 1].x, 0.f), fmaxf(acc[px][q * 4 + 1].y, 0.f));
            o.z = pack_bf16x2(fmaxf(acc[px][q * 4 + 2].x, 0.f), fmaxf(acc[px][q * 4 + 2].y, 0.f));
            o.w = pack_bf16x2(fmaxf(acc[px][q * 4 + 3].x, 0.f), fmaxf(acc[px][q * 4 + 3].y, 0.f));
            op[q] = o;
        }
    }
}

cudaError_t launch_conv1_bf16(const Conv1Weights &cw, const void *in, int in_is_u8, void *out, int B, int H, int W,
                              cudaStream_t s) {
    SamePad ph = same_pad(H, 3, 2, 1), pw = same_pad(W, 3, 2, 1);
    long long total = (long long)B * ph.out * ((pw.out + 1) / 2);      // one thread per pair of output pixels
    unsigned grid = (unsigned)((total + 127) / 128);
    if (in_is_u8)
        return launch_chain(conv1_bf16_kernel<uint8_t>, grid, 128, 0, s, 1, cw, (const uint8_t *)in, (uint4 *)out, B, H, W,
                            ph.out, pw.out, ph.before, pw.before);
    return launch_chain(conv1_bf16_kernel<float>, grid, 128, 0, s, 1, cw, (const float *)in, (uint4 *)out, B, H, W, ph.out,
                        pw.out, ph.before, pw.before);
}

// uint8 frame -> float32 / 255 (reference test&eval/model_test.py:65 `img_data / 255.`:
// float64 division, then the feed casts to float32; u8/255 rounded once to
// float32 equals that double-rounded value for all 256 inputs, see tests).
__global__ void u8_to_f32_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)((double)in[i] / 255.0);
}

cudaError_t launch_u8_to_f32(const uint8_t *in, float *out, long long n, cudaStream_t s) {
    u8_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Frame pre-processing of the reference drivers (test&eval/model_test.py:63-64, model_json.py:68-69):
// cv2.cvtColor(BGR2RGB) + cv2.resize(img, (W, H)) on uint8 frames.  The resize restates OpenCV's 8-bit
// INTER_LINEAR fixed-point scheme: source position (d + 0.5) * scale - 0.5, 11-bit coefficients
// saturate_cast<short>(f * 2048), horizontal pass in int32, vertical pass
// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.   One thread per output pixel.
// The two axes differ at the border, as they do in OpenCV: along x an out-of-range source column becomes the border
// column with the weight pair reset to (2048, 0); along y only the row indices are clipped and the fractional weight
// pair is kept (the border row is blended with itself through two truncating products).
// ---------------------------------------------------------------------------
template <bool CLAMP_WEIGHTS>
__device__ __forceinline__ void lin_taps(int d, double scale, int n, int &s0, int &s1, int &a0, int &a1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (CLAMP_WEIGHTS) {
        if (s < 0) { s = 0; f = 0.0f; }
        if (s >= n - 1) { s = n - 1; f = 0.0f; }
    }
    s0 = min(max(s, 0), n - 1);
    s1 = min(max(s + 1, 0), n - 1);
    a0 = (int)(short)__float2int_rn((1.0f - f) * 2048.0f);
    a1 = (int)(short)__float2int_rn(f * 2048.0f);
}

__global__ void preprocess_bgr_kernel(const uint8_t *__restrict__ in, int B, int Hs, int Ws, uint8_t *__restrict__ out,
                                      int H, int W, double scale_y, double scale_x) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * H * W) return;
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int b = (int)(i / ((long long)W * H));
    int x0, x1, ax0, ax1, y0, y1, ay0, ay1;
    lin_taps<true>(x, scale_x, Ws, x0, x1, ax0, ax1);
    lin_taps<false>(y, scale_y, Hs, y0, y1, ay0, ay1);
    const uint8_t *r0 = in + ((long long)b * Hs + y0) * Ws * 3, *r1 = in + ((long long)b * Hs + y1) * Ws * 3;
    uint8_t *o = out + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = (int)r0[x0 * 3 + c] * ax0 + (int)r0[x1 * 3 + c] * ax1;     // horizontal pass, scaled by 2^11
        const int h1 = (int)r1[x0 * 3 + c] * ax0 + (int)r1[x1 * 3 + c] * ax1;
        const int v = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[2 - c] = (uint8_t)min(max(v, 0), 255);                                  // BGR -> RGB
    }
}

cudaError_t launch_preprocess_bgr(const uint8_t *in, int B, int Hs, int Ws, uint8_t *out, int H, int W, cudaStream_t s) {
    const long long n = (long long)B * H * W;
    preprocess_bgr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, B, Hs, Ws, out, H, W, (double)Hs / H,
                                                                      (double)Ws / W);
    return cudaGetLastError();
}

// fp32 [rows][cols] -> bf16 [rows][cols_padded] (zero padding), used to pack weights
__global__ void pack_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long rows,
                                 int cols, int cols_padded) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols_padded) return;
    long long r = i / cols_padded;
    int c = (int)(i % cols_padded);
    dst[i] = __float2bfloat16_rn(c < cols ? src[r * cols + c] : 0.0f);
}

cudaError_t launch_pack_bf16(const float *src, void *dst, long long rows, int cols, int cols_padded,
                             cudaStream_t s) {
    long long n = rows * cols_padded;
    pack_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, (__nv_bfloat16 *)dst, rows, cols, cols_padded);
    return cudaGetLastError();
}

}  // namespace lwop
