// First layer on the tensor cores: 3x3 stride-2 conv 3 -> 32 + folded batch norm + ReLU
// (reference conv(inputs, 32, stride=2, bias=False), src/lightweight_openpose.py:10 via src/net_factory.py:4-12)
// as an implicit GEMM  out[pixel][32] = A[pixel][K] * W[32][K]^T  with K = 27 taps*channels.
//
// The 3-channel NHWC frame cannot be fetched by TMA as an operand (12- or 3-byte pixels), and on the CUDA cores
// the layer costs 864 FMAs per output pixel (fp32-FMA-bound, 0.4-0.5 ms per 256 frames).  Here 8 producer warps
// assemble the im2col rows in shared memory, directly in the swizzled K-major layout tcgen05.mma reads:
//   uint8 frames   A = the raw 0..255 values as bf16 (exact); the reference's `/ 255.` is one fp32 multiply of the
//                  accumulator in the epilogue (K padded 27 -> 32, 64-byte rows, 64B swizzle)
//   float32 frames A = (hi, lo) bf16 split of every value (x = hi + lo to 2^-17), W repeated for both halves
//                  (K = 2 x 32, 128-byte rows, 128B swizzle)
// Tile = up to 128 consecutive pixels of one output row.  The producers stream the three input row segments of
// the next kC1Depth tiles into a shared-memory ring with cp.async (4-byte copies, zero-filled outside the frame),
// so enough loads are in flight to cover the DRAM latency from one CTA per SM; one thread issues the MMAs into one
// of four TMEM accumulator stages; 4 epilogue warps scale, add the bias, apply ReLU, pack to bf16 and store 64 B
// per pixel.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "lwop_kernels.cuh"

namespace lwop {

namespace {

constexpr int kC1Threads = 512;                 // warps 0-3 control (1 = MMA, 2 = TMEM alloc), 4-11 producers, 12-15 epilogue
constexpr int kC1ProdWarp0 = 4, kC1ProdThreads = 256;
constexpr int kC1EpiWarp0 = 12;
constexpr int kC1Stages = 4;
constexpr int kC1MaxTileW = 128;
constexpr int kC1Depth = 8;                     // input staging ring: tiles in flight

struct Conv1Params {
    const void *in;                  // [B][H][W][3] uint8 or float32
    __nv_bfloat16 *out;              // [B][Ho][Wo][32]
    const __nv_bfloat16 *w;          // [32][KTOT] packed (see above)
    const float *bias;               // [32]
    int B, H, W, Ho, Wo, pad_t, pad_l;
    int tiles_per_row, tile_w, num_tiles;
    float scale;                     // 1/255 for uint8 frames (applied to the fp32 accumulator), 1 for float32 frames
};

template <bool U8>
struct C1Cfg {
    static constexpr int KTOT = U8 ? 32 : 64;
    static constexpr int KSWZ = KTOT * 2;                      // bytes per operand row: 64 / 128
    static constexpr int kABytes = 128 * KSWZ;
    static constexpr int kBBytes = 32 * KSWZ;
    static constexpr int kPix = U8 ? 3 : 12;                   // bytes per input pixel
    static constexpr int kSegBytes = (2 * kC1MaxTileW + 1) * kPix;          // one input row segment of a tile
    static constexpr int kSegWords = (kSegBytes + 3 + 3) / 4;  // + up to 3 bytes of misalignment
    static constexpr int kWordsPerThread = (3 * kSegWords + kC1ProdThreads - 1) / kC1ProdThreads;
    static constexpr int kStageBytes = ((3 * kSegWords * 4 + 127) / 128) * 128;
    static constexpr int kAOff = 0;
    static constexpr int kBOff = kAOff + kC1Stages * kABytes;
    static constexpr int kStgOff = kBOff + ((kBBytes + 1023) / 1024) * 1024;
    static constexpr int kBarOff = kStgOff + kC1Depth * kStageBytes;
    static constexpr int kTotal = kBarOff + 4 * kC1Stages * 8 + 16 + 32 * 4;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// exact bf16 of a small non-negative integer (<= 8 significant bits): the top half of its float32 pattern
__device__ __forceinline__ uint32_t u8_to_bf16_bits(uint32_t v) { return __float_as_uint((float)v) >> 16; }

template <bool U8>
__global__ void __launch_bounds__(kC1Threads, 1) conv1_tc_kernel(const Conv1Params p) {
    using L = C1Cfg<U8>;
    constexpr int KSWZ = L::KSWZ;
    constexpr uint32_t kIdesc = make_idesc_bf16(128, 32);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("lwop: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + L::kBarOff);
    uint64_t *a_empty = a_full + kC1Stages;
    uint64_t *t_full = a_empty + kC1Stages;
    uint64_t *t_empty = t_full + kC1Stages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(t_empty + kC1Stages);
    float *bias_s = reinterpret_cast<float *>(tmem_ptr + 4);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kC1Stages; ++i) {
            mbar_init(&a_full[i], 8);        // one arrival per producer warp
            mbar_init(&a_empty[i], 1);
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 4);       // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<128>(tmem_ptr);
    if (threadIdx.x < 32) bias_s[threadIdx.x] = p.bias[threadIdx.x];
    // weights [32][KTOT] -> K-major swizzled operand tile (ordinary stores, made visible to the tensor core below)
    for (int i = threadIdx.x; i < 32 * (KSWZ / 16); i += kC1Threads) {
        const int n = i / (KSWZ / 16), j = i - n * (KSWZ / 16);
        const uint4 v = *reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(p.w) + n * KSWZ + j * 16);
        const int swz = KSWZ == 128 ? (n & 7) : ((n >> 1) & 3);
        *reinterpret_cast<uint4 *>(smem + L::kBOff + n * KSWZ + ((j ^ swz) << 4)) = v;
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto tile_coords = [&](int tile, int &b, int &oy, int &ox0, int &tw) {
        const int tr = tile % p.tiles_per_row;
        const int row = tile / p.tiles_per_row;
        oy = row % p.Ho;
        b = row / p.Ho;
        ox0 = tr * p.tile_w;
        tw = min(p.tile_w, p.Wo - ox0);
    };

    if (warp == 1) {
        // ===================== MMA issuer =====================
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int s = it % kC1Stages;
            const uint32_t ph = (uint32_t)((it / kC1Stages) & 1);
            mbar_wait(&a_full[s], ph);
            fence_proxy_async();             // producers wrote the A stage with ordinary stores
            mbar_wait(&t_empty[s], ph ^ 1);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(smem + L::kAOff + s * L::kABytes), b_addr = smem_u32(smem + L::kBOff);
            if (elect_one()) {
                const uint64_t da = make_kmajor_desc<KSWZ>(a_addr), db = make_kmajor_desc<KSWZ>(b_addr);
#pragma unroll
                for (int k = 0; k < L::KTOT / 16; ++k)
                    umma_bf16_ss(tmem_base + (uint32_t)(s * 32), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc,
                                 k != 0 ? 1u : 0u);
                umma_commit(&a_empty[s]);
                umma_commit(&t_full[s]);
            }
            __syncwarp();
        }
    } else if (warp >= kC1ProdWarp0 && warp < kC1ProdWarp0 + 8) {
        // ===================== producers: input rows -> im2col operand rows =====================
        const int t = threadIdx.x - kC1ProdWarp0 * 32;
        const int m = t >> 1, hh = t & 1;                      // operand row (pixel of the tile); K half: taps 0..15 / 16..26
        const long long total_words = ((long long)p.B * p.H * p.W * L::kPix + 3) / 4;
        const uint32_t *in_words = reinterpret_cast<const uint32_t *>(p.in);
        // Stage 1 (global -> shared, asynchronous): this thread's words of a tile's three input row segments.  Rows
        // outside the frame and words past the end of the buffer are zero-filled by the copy itself (src-size 0);
        // the one column that can fall outside the row ('same' padding at x = -1 or x = W) is masked by the builder.
        auto stage_tile = [&](int tile, int slot) {
            int b, oy, ox0, tw;
            tile_coords(tile, b, oy, ox0, tw);
            const int xs = 2 * ox0 - p.pad_l;                  // first input column of the segment (may be -1)
            const uint32_t dst0 = smem_u32(smem + L::kStgOff + slot * L::kStageBytes);
#pragma unroll
            for (int i = 0; i < L::kWordsPerThread; ++i) {
                const int wi = t + i * kC1ProdThreads;         // word of the staging buffer: row (wi / kSegWords)
                const int dy = wi / L::kSegWords, wo = wi - dy * L::kSegWords;
                if (dy < 3 && wo * 4 < (2 * tw + 1) * L::kPix + 3) {     // only the part of the segment this tile reads
                    const int iy = 2 * oy - p.pad_t + dy;
                    const long long byte0 = (((long long)b * p.H + iy) * p.W + xs) * L::kPix;   // segment start
                    const long long w0 = (byte0 >> 2) + wo;    // floor division keeps the misalignment in [0, 4)
                    const bool ok = iy >= 0 && iy < p.H && w0 >= 0 && w0 < total_words;
                    const uint32_t *src = in_words + (ok ? w0 : 0);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + (uint32_t)wi * 4u), "l"(src),
                                 "r"(ok ? 4 : 0)
                                 : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // Stage 2 (shared -> operand row): 16 (or 11) of the row's 27 taps, all offsets compile-time constants.
        auto build = [&](auto HH, const unsigned char *stg, const int (&mis)[3], unsigned char *arow, int swz, bool ledge,
                         bool redge) {
            constexpr int H_ = decltype(HH)::value;
            const unsigned char *base[3];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) base[dy] = stg + dy * L::kSegWords * 4 + mis[dy] + 2 * m * L::kPix;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
                float x[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int kk = H_ * 16 + e + q;             // tap*3 + c (>= 27: zero padding)
                    x[q] = 0.0f;
                    if (kk < 27) {
                        const int tap = kk / 3, c = kk - tap * 3, dy = tap / 3, dx = tap - dy * 3;
                        if (U8) x[q] = (float)base[dy][dx * 3 + c];
                        else x[q] = *reinterpret_cast<const float *>(base[dy] + dx * 12 + c * 4);
                        if (dx == 0 && ledge) x[q] = 0.0f;       // column -1 of the frame ('same' padding)
                        if (dx == 2 && redge) x[q] = 0.0f;       // column W of the frame
                    }
                }
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(x[0], x[1]);      // exact for 0..255
                hi[e / 2] = *reinterpret_cast<const uint32_t *>(&h2);
                if (!U8) {
                    const float2 hf = __bfloat1622float2(h2);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(x[0] - hf.x, x[1] - hf.y);
                    lo[e / 2] = *reinterpret_cast<const uint32_t *>(&l2);
                }
            }
            // K chunks (16 bytes = 8 taps): this thread owns chunks {2H, 2H+1} of the hi half and, for float32
            // frames, the same two chunks of the lo half (chunks 4..7)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int chunk = 2 * H_ + j;
                *reinterpret_cast<uint4 *>(arow + ((chunk ^ swz) << 4)) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                if (!U8)
                    *reinterpret_cast<uint4 *>(arow + (((4 + chunk) ^ swz) << 4)) =
                        make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
        };
        int it = 0;
        int tile = blockIdx.x;
        // prologue: kC1Depth - 1 tiles in flight (empty groups keep the group count uniform at the tail)
        for (int d = 0; d < kC1Depth - 1; ++d) {
            const int tl = tile + d * (int)gridDim.x;
            if (tl < p.num_tiles) stage_tile(tl, d);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int s = it % kC1Stages;
            const uint32_t ph = (uint32_t)((it / kC1Stages) & 1);
            const int slot = it % kC1Depth;
            asm volatile("cp.async.wait_group %0;" ::"n"(kC1Depth - 2) : "memory");   // this tile's copies have landed
            named_bar_sync(1, kC1ProdThreads);                  // ... for every producer thread; slot (it-1) is drained
            {   // refill the slot drained in the previous iteration with the tile kC1Depth - 1 ahead
                const int tl = tile + (kC1Depth - 1) * (int)gridDim.x;
                if (tl < p.num_tiles) stage_tile(tl, (it + kC1Depth - 1) % kC1Depth);
                else asm volatile("cp.async.commit_group;" ::: "memory");
            }
            const unsigned char *stg = smem + L::kStgOff + slot * L::kStageBytes;
            int b, oy, ox0, tw;
            tile_coords(tile, b, oy, ox0, tw);
            const int xs = 2 * ox0 - p.pad_l;
            mbar_wait(&a_empty[s], ph ^ 1);
            if (m < tw) {
                // each staged segment starts at byte (its global byte address & 3) of its first word
                int mis[3];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const long long byte0 = (((long long)b * p.H + (2 * oy - p.pad_t + dy)) * p.W + xs) * L::kPix;
                    mis[dy] = (int)(byte0 & 3);
                }
                unsigned char *arow = smem + L::kAOff + s * L::kABytes + m * KSWZ;
                const int swz = KSWZ == 128 ? (m & 7) : ((m >> 1) & 3);
                const int ix0 = xs + 2 * m;                      // input column of tap dx = 0
                const bool ledge = ix0 < 0, redge = ix0 + 2 >= p.W;
                if (hh == 0) build(std::integral_constant<int, 0>{}, stg, mis, arow, swz, ledge, redge);
                else build(std::integral_constant<int, 1>{}, stg, mis, arow, swz, ledge, redge);
            }
            __syncwarp();                                       // orders the lanes' stores before lane 0's release
            if (lane == 0) mbar_arrive(&a_full[s]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp >= kC1EpiWarp0) {
        // ===================== epilogue =====================
        const int q = warp - kC1EpiWarp0;                      // TMEM lane quarter == warp % 4
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int s = it % kC1Stages;
            const uint32_t ph = (uint32_t)((it / kC1Stages) & 1);
            int b, oy, ox0, tw;
            tile_coords(tile, b, oy, ox0, tw);
            mbar_wait(&t_full[s], ph);
            tcgen05_fence_after();
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 32), r);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[s]);
            const int m = q * 32 + lane;
            if (m < tw) {
                uint4 *op = reinterpret_cast<uint4 *>(p.out + ((((long long)b * p.Ho + oy) * p.Wo + ox0 + m) << 5));
                const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s);
                const float2 sc = make_float2(p.scale, p.scale);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                    const float2 s0 = ffma2(make_float2(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1])), sc, make_float2(b0.x, b0.y));
                    const float2 s1 = ffma2(make_float2(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3])), sc, make_float2(b0.z, b0.w));
                    const float2 s2 = ffma2(make_float2(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5])), sc, make_float2(b1.x, b1.y));
                    const float2 s3 = ffma2(make_float2(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7])), sc, make_float2(b1.z, b1.w));
                    uint4 o;
                    o.x = pack_relu_bf16x2(s0.x, s0.y); o.y = pack_relu_bf16x2(s1.x, s1.y);
                    o.z = pack_relu_bf16x2(s2.x, s2.y); o.w = pack_relu_bf16x2(s3.x, s3.y);
                    op[j] = o;
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<128>(tmem_base);
    }
}

uint16_t host_bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);      // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

}  // namespace

// Packs the folded first-layer weights [32][27] fp32 into the two bf16 operand tables:
//   w_u8  [32][32]: W, K padded with zeros (the `/ 255.` is applied to the accumulator)
//   w_f32 [32][64]: W at k = 0..26 (pairs with the hi halves) and again at k = 32..58 (lo halves)
void conv1_tc_pack_weights(const float *w_n27, uint16_t *w_u8, uint16_t *w_f32) {
    for (int n = 0; n < 32; ++n) {
        for (int k = 0; k < 32; ++k) w_u8[n * 32 + k] = k < 27 ? host_bf16_rn(w_n27[n * 27 + k]) : 0;
        for (int k = 0; k < 64; ++k) {
            const int kk = k & 31;
            w_f32[n * 64 + k] = kk < 27 ? host_bf16_rn(w_n27[n * 27 + kk]) : 0;
        }
    }
}

cudaError_t launch_conv1_tc(const void *in, int in_is_u8, void *out, const void *w_u8_dev, const void *w_f32_dev,
                            const float *bias_dev, int B, int H, int W, cudaStream_t s) {
    SamePad ph = same_pad(H, 3, 2, 1), pw = same_pad(W, 3, 2, 1);
    Conv1Params p{};
    p.in = in; p.out = (__nv_bfloat16 *)out; p.bias = bias_dev;
    p.w = (const __nv_bfloat16 *)(in_is_u8 ? w_u8_dev : w_f32_dev);
    p.B = B; p.H = H; p.W = W; p.Ho = ph.out; p.Wo = pw.out; p.pad_t = ph.before; p.pad_l = pw.before;
    p.tiles_per_row = (pw.out + kC1MaxTileW - 1) / kC1MaxTileW;
    p.tile_w = (pw.out + p.tiles_per_row - 1) / p.tiles_per_row;
    p.num_tiles = B * ph.out * p.tiles_per_row;
    p.scale = in_is_u8 ? (float)(1.0 / 255.0) : 1.0f;
    int sms = gemm_num_sms();
    if (sms <= 0) sms = 148;
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    cudaError_t e;
    if (in_is_u8) {
        e = cudaFuncSetAttribute(conv1_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1Cfg<true>::kTotal);
        if (e != cudaSuccess) return e;
        conv1_tc_kernel<true><<<grid, kC1Threads, C1Cfg<true>::kTotal, s>>>(p);
    } else {
        e = cudaFuncSetAttribute(conv1_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1Cfg<false>::kTotal);
        if (e != cudaSuccess) return e;
        conv1_tc_kernel<false><<<grid, kC1Threads, C1Cfg<false>::kTotal, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace lwop
