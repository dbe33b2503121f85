// Depthwise 3x3 (stride 1/2, dilation 1/2), bf16 NHWC -> bf16 NHWC, HBM-bound.
// Depthwise half of tf.layers.separable_conv2d (reference src/net_factory.py:15,25):
// channel multiplier 1, 'same' padding, no bias, no activation.
//
// Persistent CTAs; each tile = TH x TW output pixels x CB channels.  The input
// tile with its halo is fetched by ONE 4D TMA box load (C, W, H, N) into shared
// memory, double-buffered so the next tile's load overlaps this tile's math;
// 'same' zero padding is the tensor map's out-of-bounds fill (box coordinates may
// be negative).  Each thread owns 4 channels of one output column of the tile
// and slides down R output rows, so every shared-memory vector is read once per
// thread; the 9x4 filter taps live in registers; fp32 accumulation; 64-bit
// shared loads, 64-bit global stores (a warp writes two full 128-byte pixels).
#include <stdlib.h>

#include "lwop_kernels.cuh"

namespace lwop {

namespace {

constexpr int kDwThreads = 256;

template <int STRIDE, int DIL, int CB, int TW, int TH>
struct DwGeom {
    static constexpr int TPP = CB / 4;                 // threads per pixel
    static constexpr int SLOTS = kDwThreads / TPP;     // pixel columns processed concurrently
    static constexpr int ROWG = SLOTS / TW;            // row groups
    static constexpr int R = TH / ROWG;                // output rows per thread
    static constexpr int IW = (TW - 1) * STRIDE + 2 * DIL + 1;
    static constexpr int IH = (TH - 1) * STRIDE + 2 * DIL + 1;
    static constexpr int IRN = (R - 1) * STRIDE + 2 * DIL + 1;   // input rows one thread walks
    static constexpr int kBufBytes = ((IW * IH * CB * 2 + 127) / 128) * 128;
    static constexpr int kSmem = 2 * kBufBytes + 128 + 128;
    static_assert(SLOTS % TW == 0 && TH % ROWG == 0 && ROWG >= 1, "bad depthwise tile");
};

struct DwParams {
    const float *w;            // [9][C]
    __nv_bfloat16 *out;
    int B, Ho, Wo, C;
    int pad_t, pad_l;
    int tiles_x, tiles_y, cblocks;
    int spatial_tiles;         // B * tiles_y * tiles_x
};

template <int STRIDE, int DIL, int CB, int TW, int TH>
__global__ void __launch_bounds__(kDwThreads)
dw_tma_kernel(const __grid_constant__ CUtensorMap map_in, const DwParams p) {
    using G = DwGeom<STRIDE, DIL, CB, TW, TH>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) &
                                                            ~static_cast<uintptr_t>(127));
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * G::kBufBytes);

    const int tid = threadIdx.x;
    const int c4 = tid % G::TPP;                   // channel quad inside the channel block
    const int slot = tid / G::TPP;
    const int col = slot % TW, rg = slot / TW;

    // CTA i works on channel block i % cblocks for its whole life: taps loaded once
    const int cb = blockIdx.x % p.cblocks;
    const int sp0 = blockIdx.x / p.cblocks;
    const int sp_step = gridDim.x / p.cblocks;
    const int ch0 = cb * CB + c4 * 4;

    float2 wt[9][2];           // taps as fp32 pairs for the packed FFMA2
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 v = *reinterpret_cast<const float4 *>(p.w + (size_t)k * p.C + ch0);
        wt[k][0] = make_float2(v.x, v.y);
        wt[k][1] = make_float2(v.z, v.w);
    }
    if (tid == 0) {
        tma_prefetch_desc(&map_in);
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();                  // the previous kernel of the chain has completed
    pdl_launch_dependents();

    auto issue = [&](int sp, int buf) {
        const int tx = sp % p.tiles_x;
        const int t2 = sp / p.tiles_x;
        const int ty = t2 % p.tiles_y;
        const int b = t2 / p.tiles_y;
        mbar_arrive_expect_tx(&bar[buf], (uint32_t)(G::IW * G::IH * CB * 2));
        tma_load_4d(smem + buf * G::kBufBytes, &map_in, &bar[buf], cb * CB, tx * TW * STRIDE - p.pad_l,
                    ty * TH * STRIDE - p.pad_t, b);
    };

    if (tid == 0 && sp0 < p.spatial_tiles) issue(sp0, 0);
    int it = 0;
    for (int sp = sp0; sp < p.spatial_tiles; sp += sp_step, ++it) {
        const int buf = it & 1;
        const int nxt = sp + sp_step;
        if (tid == 0 && nxt < p.spatial_tiles) issue(nxt, buf ^ 1);     // buffer buf^1 was drained last iteration
        mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));

        const int tx = sp % p.tiles_x;
        const int t2 = sp / p.tiles_x;
        const int ty = t2 % p.tiles_y;
        const int b = t2 / p.tiles_y;
        // this thread's first input vector of the tile (shared-space address; everything else is an immediate offset)
        const uint32_t tbase = smem_u32(smem + buf * G::kBufBytes) +
                               (uint32_t)(((rg * G::R * STRIDE) * G::IW + col * STRIDE) * CB + c4 * 4) * 2u;

        float2 acc[G::R][2];
#pragma unroll
        for (int r = 0; r < G::R; ++r) {
            acc[r][0] = make_float2(0.0f, 0.0f);
            acc[r][1] = make_float2(0.0f, 0.0f);
        }

#pragma unroll
        for (int ir = 0; ir < G::IRN; ++ir) {
            float2 v[3][2];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const uint2 raw = lds_u64(tbase + (uint32_t)((ir * G::IW + dx * DIL) * CB * 2));
                v[dx][0] = bf16x2_to_float2(raw.x);
                v[dx][1] = bf16x2_to_float2(raw.y);
            }
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int num = ir - dy * DIL;
                if (num >= 0 && num % STRIDE == 0 && num / STRIDE < G::R) {
                    const int r = num / STRIDE;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        acc[r][0] = ffma2(v[dx][0], wt[dy * 3 + dx][0], acc[r][0]);
                        acc[r][1] = ffma2(v[dx][1], wt[dy * 3 + dx][1], acc[r][1]);
                    }
                }
            }
        }
        const int ox = tx * TW + col;
        const int oy0 = ty * TH + rg * G::R;
        if (ox < p.Wo) {
            __nv_bfloat16 *op = p.out + (((size_t)b * p.Ho + oy0) * p.Wo + ox) * p.C + ch0;
            const size_t row_pitch = (size_t)p.Wo * p.C;
#pragma unroll
            for (int r = 0; r < G::R; ++r) {
                if (oy0 + r < p.Ho) {
                    uint2 o;
                    o.x = pack_bf16x2(acc[r][0].x, acc[r][0].y);
                    o.y = pack_bf16x2(acc[r][1].x, acc[r][1].y);
                    *reinterpret_cast<uint2 *>(op) = o;
                }
                op += row_pitch;
            }
        }
        __syncthreads();      // every thread is done with `buf` before it is refilled next iteration
    }
}

template <int STRIDE, int DIL, int CB, int TW, int TH>
cudaError_t launch_dw_t(const DwPlan *pl, cudaStream_t s);

}  // namespace

struct DwPlan {
    CUtensorMap map_in;
    DwParams p;
    int stride, dil, cb;
    int grid;
    int max_ctas;      // resident CTAs on the device (occupancy x SMs); the grid never exceeds it
};

namespace {

template <int STRIDE, int DIL, int CB, int TW, int TH>
cudaError_t launch_dw_t(const DwPlan *pl, cudaStream_t s) {
    using G = DwGeom<STRIDE, DIL, CB, TW, TH>;
    auto kfn = dw_tma_kernel<STRIDE, DIL, CB, TW, TH>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmem);
    if (e != cudaSuccess) return e;
    return launch_chain(kfn, (unsigned)pl->grid, kDwThreads, G::kSmem, s, 1, pl->map_in, pl->p);
}

// resident CTAs per SM of the instantiation a plan will launch
template <int STRIDE, int DIL, int CB, int TW, int TH>
int occupancy_t() {
    using G = DwGeom<STRIDE, DIL, CB, TW, TH>;
    auto kfn = dw_tma_kernel<STRIDE, DIL, CB, TW, TH>;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmem) != cudaSuccess) return 1;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kfn, kDwThreads, G::kSmem) != cudaSuccess || n < 1) n = 1;
    return n;
}

// tile shapes per (stride, dilation)
constexpr int kTW1 = 16, kTH1 = 8;     // stride 1 (dilation 1 or 2)
constexpr int kTW2 = 8, kTH2 = 8;      // stride 2

}  // namespace

bool dw_plan_supported(int C, int stride, int dil) {
    return (C % 32 == 0) && ((stride == 1 && (dil == 1 || dil == 2)) || (stride == 2 && dil == 1));
}

DwPlan *dw_plan_create(const DwArgs &a, char *err, size_t errlen) {
    if (!dw_plan_supported(a.C, a.stride, a.dil) || a.dtype != LWOP_DT_BF16) {
        snprintf(err, errlen, "depthwise TMA kernel: unsupported C=%d stride=%d dil=%d", a.C, a.stride, a.dil);
        return nullptr;
    }
    if (!gemm_driver_ready(err, errlen)) return nullptr;
    DwPlan *pl = (DwPlan *)calloc(1, sizeof(DwPlan));
    SamePad ph = same_pad(a.H, 3, a.stride, a.dil), pw = same_pad(a.W, 3, a.stride, a.dil);
    const int cb = a.C % 64 == 0 ? 64 : 32;
    const int tw = a.stride == 1 ? kTW1 : kTW2, th = a.stride == 1 ? kTH1 : kTH2;
    pl->stride = a.stride;
    pl->dil = a.dil;
    pl->cb = cb;
    DwParams &p = pl->p;
    p.w = a.w;
    p.out = (__nv_bfloat16 *)a.out;
    p.B = a.B; p.Ho = ph.out; p.Wo = pw.out; p.C = a.C;
    p.pad_t = ph.before; p.pad_l = pw.before;
    p.tiles_x = (pw.out + tw - 1) / tw;
    p.tiles_y = (ph.out + th - 1) / th;
    p.cblocks = a.C / cb;
    p.spatial_tiles = a.B * p.tiles_x * p.tiles_y;
    int sms = gemm_num_sms();
    if (sms <= 0) sms = 148;
    int occ = 1;
    if (a.stride == 1 && a.dil == 1) occ = cb == 64 ? occupancy_t<1, 1, 64, kTW1, kTH1>() : occupancy_t<1, 1, 32, kTW1, kTH1>();
    else if (a.stride == 1) occ = cb == 64 ? occupancy_t<1, 2, 64, kTW1, kTH1>() : occupancy_t<1, 2, 32, kTW1, kTH1>();
    else occ = cb == 64 ? occupancy_t<2, 1, 64, kTW2, kTH2>() : occupancy_t<2, 1, 32, kTW2, kTH2>();
    // persistent grid = exactly the resident CTAs (one wave), a multiple of cblocks so that each CTA keeps
    // one channel block (its taps stay in registers)
    pl->max_ctas = occ * sms;
    int per_cb = pl->max_ctas / p.cblocks;
    if (per_cb > p.spatial_tiles) per_cb = p.spatial_tiles;
    if (per_cb < 1) per_cb = 1;
    pl->grid = per_cb * p.cblocks;
    const int iw = (tw - 1) * a.stride + 2 * a.dil + 1, ih = (th - 1) * a.stride + 2 * a.dil + 1;
    cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t strides[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.W * a.C * 2, (cuuint64_t)a.H * a.W * a.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)iw, (cuuint32_t)ih, 1};
    if (!encode_tiled_bf16(&pl->map_in, 4, a.in, dims, strides, box, /*swizzle128=*/0, err, errlen)) {
        free(pl);
        return nullptr;
    }
    return pl;
}

void dw_plan_destroy(DwPlan *p) { free(p); }

cudaError_t dw_plan_launch(const DwPlan *pl, cudaStream_t s) {
    if (pl->stride == 1 && pl->dil == 1)
        return pl->cb == 64 ? launch_dw_t<1, 1, 64, kTW1, kTH1>(pl, s) : launch_dw_t<1, 1, 32, kTW1, kTH1>(pl, s);
    if (pl->stride == 1 && pl->dil == 2)
        return pl->cb == 64 ? launch_dw_t<1, 2, 64, kTW1, kTH1>(pl, s) : launch_dw_t<1, 2, 32, kTW1, kTH1>(pl, s);
    if (pl->stride == 2 && pl->dil == 1)
        return pl->cb == 64 ? launch_dw_t<2, 1, 64, kTW2, kTH2>(pl, s) : launch_dw_t<2, 1, 32, kTW2, kTH2>(pl, s);
    return cudaErrorInvalidValue;
}

}  // namespace lwop
