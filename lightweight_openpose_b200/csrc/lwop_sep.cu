// Fused separable convolution: depthwise 3x3 (stride 1, dilation 1/2) -> pointwise 1x1 on the
// tensor cores, one kernel, the depthwise result never leaves the SM.
//
//   out[pixel, n] = act( sum_c dw(x)[pixel, c] * Wpw[n, c] + bias[n] ) (+ residual)
//   dw(x)[pixel, c] = sum_{3x3 taps} x[pixel + tap*dil, c] * Wdw[tap, c]          ('same' zero padding)
//
// This is tf.layers.separable_conv2d as the reference uses it (src/net_factory.py:15-19 conv_dw with
// the folded batch norm + ReLU, :25-29 conv_dw_no_bn with ELU; nothing sits between the depthwise and
// the pointwise half) plus the tf.add of lightweight_openpose.py:27 as an optional residual.
//
// CTA tile = an 8 x 16 patch of output pixels (= the 128 rows of one tcgen05 MMA) x BLOCK_N output
// channels; K (= input channels) is walked in chunks of CB channels:
//   warp 3      TMA producer: per chunk ONE 4D box load of the input patch + halo (zero OOB fill is the
//               'same' padding) plus the chunk's 9 x CB depthwise taps into a slab ring
//   warp 0      TMA producer: the [BLOCK_N x CB] pointwise weight tiles into the B ring
//   warps 4-11  depthwise producers (CUDA cores): slab -> 9-tap fp32 FMA (packed FFMA2) -> bf16 ->
//               written straight into the A ring in the 128B/64B-swizzled K-major layout the MMA
//               expects (generic-proxy stores, released through the mbarrier; the MMA thread crosses to
//               the async proxy with one fence.proxy.async per chunk after its acquire)
//   warp 1      tcgen05.mma issuer (A from the dw warps, B from TMA), fp32 accumulators in TMEM
//   warp 2      TMEM allocator
//   warps 12-19 epilogue, two groups of 4 warps: tcgen05.ld -> bias / activation / residual -> bf16 ->
//               swizzled staging -> 4D TMA store (the patch is clipped at the image border by the store)
// Two accumulator schemes (template NH):
//   NH = 1  BLOCK_N <= 256 columns per tile, two accumulator stages; epilogue group g owns stage g (every
//           other tile), so the drain of tile i overlaps the main loop of tile i+1.  Cout = 512 would need
//           two N tiles per pixel patch, i.e. the depthwise work done twice.
//   NH = 2  Cout = 512: ONE pass over the patch feeds two N = 256 MMAs per K chunk (all 512 TMEM columns,
//           one accumulator stage); group h drains half h.  The depthwise producers are the slower side of
//           this pipeline (~73.7k fp32 FMAs per chunk vs 1024 tensor cycles), and they run ahead through
//           the A ring while the epilogue drains, so the single accumulator stage costs little.
// Compared with the depthwise kernel followed by the 1x1 GEMM this removes one HBM write and one HBM
// read of the [B,H,W,C] intermediate per layer.
#include <stdlib.h>

#include "lwop_kernels.cuh"

namespace lwop {

namespace {

constexpr int kTW = 16, kTH = 8;                 // output patch (w x h) = 128 GEMM rows
constexpr int kDwWarp0 = 4, kDwWarps = 8, kDwThreads = kDwWarps * 32;
constexpr int kEpiWarp0 = 12, kEpiWarps = 8;               // two epilogue groups of 4 warps (one per TMEM lane quarter)
constexpr int kSepThreads = (kEpiWarp0 + kEpiWarps) * 32;   // 640

struct SepParams {
    const float *dw_w;              // [9][C] fp32
    const float *bias;              // [Cout_pad]
    const __nv_bfloat16 *residual;  // NHWC [B,H,W,ldr] or nullptr
    int B, H, W, C;
    int Cout_pad;
    int tiles_x, tiles_y, num_m_tiles, num_n_tiles;
    int k_blocks;                   // C / CB
    int act, ldr, out_col0;
    unsigned long long *prof;       // optional [16] stall counters (LWOP_SEP_PROF=1), nullptr otherwise
    unsigned park;                  // roles whose barrier waits park the warp (bit 0 TMA, 1 MMA, 2 depthwise, 3 epilogue)
    int wide64;                     // 64-column staging tiles + TMA stores in the epilogue where the shape allows (LWOP_SEP_WIDE64=0: 32-column ones)
};

// stall accounting (debugging aid): cycles one elected thread of each role spends in each barrier wait
enum { PF_TOTAL = 0, PF_TMA_SLAB_EMPTY, PF_TMA_AB_EMPTY, PF_MMA_TMEM_EMPTY, PF_MMA_B_FULL, PF_MMA_A_FULL,
       PF_DW_SLAB_FULL, PF_DW_AB_EMPTY, PF_DW_COMPUTE, PF_EPI_TMEM_FULL, PF_EPI_STORE_WAIT, PF_EPI_WORK, PF_EPI_LDTM, PF_EPI_MATH, PF_EPI_FENCE, PF_N };

#define PF_WAIT(slot, stmt)                                     \
    do {                                                        \
        if (pf) {                                               \
            const long long _t0 = clock64();                    \
            stmt;                                               \
            pf_acc[slot] += (unsigned long long)(clock64() - _t0); \
        } else {                                                \
            stmt;                                               \
        }                                                       \
    } while (0)

template <int BLOCK_N, int CB, int DIL, int SA, int SB, int SS, int NH = 1>
struct SepSmem {
    static constexpr int KSWZ = 2 * CB;                      // bytes per K-major operand row
    static constexpr int kABytes = 128 * KSWZ;
    static constexpr int kBBytes = BLOCK_N * KSWZ;
    static constexpr int IW = kTW + 2 * DIL, IH = kTH + 2 * DIL;
    static constexpr int kSlabPix = IW * IH * CB * 2;        // input patch + halo, bf16
    static constexpr int kTapBytes = 9 * CB * 4;             // this chunk's depthwise taps, fp32 [9][CB]
    static constexpr int kSlabTx = kSlabPix + kTapBytes;
    static constexpr int kSlabBytes = ((kSlabTx + 127) / 128) * 128;
    static constexpr int kAOff = 0;
    static constexpr int kBOff = kAOff + SA * kABytes;
    // 8 epilogue warps x 4 KB (32 rows x 64 bf16, 128B swizzle; the 32-column form uses the first 2 KB with a 64B
    // swizzle); the single-pass Cout = 512 form (NH == 2) only has the 32-column epilogue: 2 KB
    static constexpr bool kWide = NH == 1 && BLOCK_N >= 64;
    static constexpr int kStagingBytes = kWide ? 4096 : 2048;
    static constexpr int kStagingOff = kBOff + SB * kBBytes;
    static constexpr int kSlabOff = kStagingOff + kEpiWarps * kStagingBytes;
    static constexpr int kBarOff = kSlabOff + SS * kSlabBytes;
    static constexpr int kNumBars = 2 * SS + 2 * SA + 2 * SB + 4;
    static constexpr int kBiasOff = ((kBarOff + kNumBars * 8 + 8 + 15) / 16) * 16;
    static constexpr int kTotal = kBiasOff + 512 * 4;
    static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "operand stages must stay 1 KB aligned");
    static_assert(kSlabPix % 128 == 0, "tap block must stay 128-byte aligned for TMA");
    static_assert(kTotal <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void tma_store_4d(const void *desc, const void *smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// bf16 pair -> fp32 pair on the ALU pipe (PRMT), leaving the FMA pipe to the packed FMAs
__device__ __forceinline__ float2 unpack_bf16x2_prmt(uint32_t v) {
    uint32_t lo, hi;
    asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(v));
    asm("prmt.b32 %0, %1, 0, 0x3244;" : "=r"(hi) : "r"(v));
    return make_float2(__uint_as_float(lo), __uint_as_float(hi));
}
__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}

template <int BLOCK_N, int CB, int DIL, int SA, int SB, int SS, int NH, bool PROF>
__global__ void __launch_bounds__(kSepThreads, 1)
sep_conv_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_taps,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_out64, const SepParams p) {
    using L = SepSmem<BLOCK_N, CB, DIL, SA, SB, SS, NH>;
    constexpr int KSWZ = L::KSWZ;
    constexpr int kMmasPerBlock = CB / 16;
    constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(2 * BLOCK_N <= 512, "two accumulator stages / halves must fit in TMEM");
    static_assert(NH == 1 || NH == 2, "NH");
    constexpr uint32_t kIdesc = make_idesc_bf16(128, BLOCK_N);

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("lwop: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t *slab_full = reinterpret_cast<uint64_t *>(smem + L::kBarOff);
    uint64_t *slab_empty = slab_full + SS;
    uint64_t *a_full = slab_empty + SS;
    uint64_t *a_empty = a_full + SA;
    uint64_t *b_full = a_empty + SA;
    uint64_t *b_empty = b_full + SB;
    uint64_t *tmem_full = b_empty + SB;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *bias_s = reinterpret_cast<float *>(smem + L::kBiasOff);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    constexpr bool pf = PROF;
    unsigned long long pf_acc[PROF ? PF_N : 1];
#pragma unroll
    for (int i = 0; i < (PROF ? PF_N : 1); ++i) pf_acc[i] = 0;
    const long long pf_start = PROF ? clock64() : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_in);
        tma_prefetch_desc(&map_taps);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SS; ++i) {
            mbar_init(&slab_full[i], 1);
            mbar_init(&slab_empty[i], kDwWarps);
        }
        for (int i = 0; i < SA; ++i) {
            mbar_init(&a_full[i], kDwWarps);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < SB; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);             // one arrival per warp of the epilogue group
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr);
    for (int i = threadIdx.x; i < p.Cout_pad; i += kSepThreads) bias_s[i] = p.bias[i];
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();                  // the previous kernel of the chain has completed: its output may be read, our output written
    pdl_launch_dependents();     // the next kernel may start its prologue while this one runs

    if (warp == 0) {
        // ===================== TMA producer: pointwise weight tiles (B ring) =====================
        if (lane == 0) {
            int sb = 0;
            uint32_t bph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        PF_WAIT(PF_TMA_AB_EMPTY, mbar_wait_sel(&b_empty[sb], bph ^ 1, p.park & 1u));
                        mbar_arrive_expect_tx(&b_full[sb], L::kBBytes);
                        tma_load_2d(smem + L::kBOff + sb * L::kBBytes, &map_b, &b_full[sb], kb * CB,
                                    (n_tile * NH + h) * BLOCK_N);
                        if (++sb == SB) { sb = 0; bph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== TMA producer: input slabs + depthwise taps (own thread, so that a full B ring
        // never delays the loads the depthwise warps are waiting for) =====================
        if (lane == 0) {
            int ss = 0;
            uint32_t sph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.num_n_tiles;
                const int b = m_tile / tiles_per_img;
                const int t2 = m_tile - b * tiles_per_img;
                const int ty = t2 / p.tiles_x, tx = t2 - ty * p.tiles_x;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    PF_WAIT(PF_TMA_SLAB_EMPTY, mbar_wait_sel(&slab_empty[ss], sph ^ 1, p.park & 1u));
                    unsigned char *slab = smem + L::kSlabOff + ss * L::kSlabBytes;
                    mbar_arrive_expect_tx(&slab_full[ss], L::kSlabTx);
                    tma_load_4d(slab, &map_in, &slab_full[ss], kb * CB, tx * kTW - DIL, ty * kTH - DIL, b);
                    tma_load_2d(slab + L::kSlabPix, &map_taps, &slab_full[ss], kb * CB, 0);
                    if (++ss == SS) { ss = 0; sph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // whole warp walks the warp-uniform schedule and waits; one elected lane issues (descriptors stay uniform)
        {
            int sa = 0, sb = 0;
            uint32_t aph = 0, bph = 0;
            int local_tile = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
                // NH = 1: accumulator stage = local_tile & 1 (used every other tile); NH = 2: half h, used every tile
                const int acc = NH == 1 ? (local_tile & 1) : 0;
                const uint32_t acc_phase = NH == 1 ? (uint32_t)((local_tile >> 1) & 1) : (uint32_t)(local_tile & 1);
                if (NH == 1) {
                    PF_WAIT(PF_MMA_TMEM_EMPTY, mbar_wait_sel(&tmem_empty[acc], acc_phase ^ 1, p.park & 2u));
                    tcgen05_fence_after();
                }
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    PF_WAIT(PF_MMA_A_FULL, mbar_wait_sel(&a_full[sa], aph, p.park & 2u));
                    // the A stage was written with ordinary (generic-proxy) stores released through a_full;
                    // this fence makes them visible to the tensor core's async-proxy reads issued below
                    fence_proxy_async();
                    const uint32_t a_addr = smem_u32(smem + L::kAOff + sa * L::kABytes);
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        if (NH == 2 && kb == 0) PF_WAIT(PF_MMA_TMEM_EMPTY, mbar_wait_sel(&tmem_empty[h], acc_phase ^ 1, p.park & 2u));
                        PF_WAIT(PF_MMA_B_FULL, mbar_wait_sel(&b_full[sb], bph, p.park & 2u));
                        tcgen05_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)((NH == 1 ? acc : h) * BLOCK_N);
                        const uint32_t b_addr = smem_u32(smem + L::kBOff + sb * L::kBBytes);
                        if (elect_one()) {
                            const uint64_t da = make_kmajor_desc<KSWZ>(a_addr);
                            const uint64_t db = make_kmajor_desc<KSWZ>(b_addr);
#pragma unroll
                            for (int k = 0; k < kMmasPerBlock; ++k)
                                umma_bf16_ss(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc,
                                             (kb | k) != 0 ? 1u : 0u);
                            umma_commit(&b_empty[sb]);
                            if (h == NH - 1) umma_commit(&a_empty[sa]);
                            if (kb == p.k_blocks - 1) {
                                if (NH == 1) umma_commit(&tmem_full[acc]);
                                else if (h == NH - 1) { umma_commit(&tmem_full[0]); umma_commit(&tmem_full[1]); }
                            }
                        }
                        __syncwarp();
                        if (++sb == SB) { sb = 0; bph ^= 1; }
                    }
                    if (++sa == SA) { sa = 0; aph ^= 1; }
                }
            }
        }
    } else if (warp >= kDwWarp0 && warp < kDwWarp0 + kDwWarps) {
        // ===================== depthwise producers =====================
        int sa = 0, ss = 0;
        uint32_t aph = 0, sph = 0;
        if constexpr (CB == 64) {
            // One warp = one 4 x 4 block of output pixels x all 64 channels of the chunk; lane = channel pair.
            // Every LDS.32 / STS.32 of a warp touches one full 128-byte pixel row (one shared-memory wavefront),
            // and a (4+2d)^2 input block feeds 16 outputs: 2.25 (d=1) loads per output instead of 9.
            constexpr int IB = 4 + 2 * DIL;                     // input block edge
            const int wq = warp - kDwWarp0;
            const int bx = wq & 3, by = wq >> 2;
            const uint32_t slab_off = (uint32_t)((((by * 4) * L::IW + bx * 4) * CB + lane * 2) * 2);
            const uint32_t tap_off = (uint32_t)(L::kSlabPix + lane * 8);
            // A row m = (by*4 + oy)*16 + bx*4 + ox; 128B swizzle phase m & 7 = ((bx & 1)*4 + ox)
            uint32_t a_col[4];
#pragma unroll
            for (int ox = 0; ox < 4; ++ox)
                a_col[ox] = (uint32_t)(((by * 4) * kTW + bx * 4 + ox) * 128 +
                                       ((((lane >> 2) ^ (((bx & 1) << 2) | ox)) << 4) | ((lane & 3) << 2)));
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    float2 acc[4][4];
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy)
#pragma unroll
                        for (int ox = 0; ox < 4; ++ox) acc[oy][ox] = make_float2(0.0f, 0.0f);
                    PF_WAIT(PF_DW_SLAB_FULL, mbar_wait_sel(&slab_full[ss], sph, p.park & 4u));
                    const long long pf_c0 = pf ? clock64() : 0;
                    const uint32_t sbase = smem_u32(smem + L::kSlabOff + ss * L::kSlabBytes);
                    const uint32_t tbase = sbase + slab_off;
                    float2 wt[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const uint2 v = lds_u64(sbase + tap_off + (uint32_t)(k * CB * 4));
                        wt[k] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
                    }
#pragma unroll
                    for (int iy = 0; iy < IB; ++iy) {
                        float2 v[IB];
#pragma unroll
                        for (int ix = 0; ix < IB; ++ix) {
                            uint32_t raw;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw)
                                         : "r"(tbase + (uint32_t)((iy * L::IW + ix) * CB * 2)));
                            v[ix] = unpack_bf16x2_prmt(raw);
                        }
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const int oy = iy - dy * DIL;
                            if (oy >= 0 && oy < 4) {
#pragma unroll
                                for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                                    for (int dx = 0; dx < 3; ++dx)
                                        acc[oy][ox] = ffma2(v[ox + dx * DIL], wt[dy * 3 + dx], acc[oy][ox]);
                            }
                        }
                    }
                    // The slab is handed back to the TMA producer only AFTER the operand stores below: they consume every
                    // accumulator, hence every load of the slab has returned its data by then.  Arriving right after the
                    // loop is a write-after-read race -- nothing orders the barrier arrive (no register dependence) behind
                    // shared-memory loads still in flight, and a refill that lands first feeds the last input rows of the
                    // block with the NEXT chunk's pixels (seen as run-to-run differences in the layer with the slowest
                    // epilogue, whose producer is always waiting for a free slab).
                    if (pf) pf_acc[PF_DW_COMPUTE] += (unsigned long long)(clock64() - pf_c0);
                    PF_WAIT(PF_DW_AB_EMPTY, mbar_wait_sel(&a_empty[sa], aph ^ 1, p.park & 4u));   // the MMAs that read this A stage retired
                    const uint32_t abase = smem_u32(smem + L::kAOff + sa * L::kABytes);
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy)
#pragma unroll
                        for (int ox = 0; ox < 4; ++ox)
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(abase + a_col[ox] + (uint32_t)(oy * kTW * 128)),
                                         "r"(pack_bf16x2(acc[oy][ox].x, acc[oy][ox].y))
                                         : "memory");
                    // every writer makes ITS OWN generic-proxy stores visible to the async proxy (the tensor core reads the
                    // operand through it) before the release; a fence on the consumer side alone is not enough -- it left
                    // rare stale operand reads (1-2 images per 256 with 1e-3 differences from run to run)
                    fence_proxy_async();
                    __syncwarp();                                          // orders the lanes' stores before lane 0's release
                    if (lane == 0) {
                        mbar_arrive(&a_full[sa]);
                        mbar_arrive(&slab_empty[ss]);                      // slab drained by this warp (see above)
                    }
                    if (++ss == SS) { ss = 0; sph ^= 1; }
                    if (++sa == SA) { sa = 0; aph ^= 1; }
                }
            }
        } else {
        constexpr int TPP = CB / 4;                    // threads per pixel (4 channels each)
        constexpr int SLOTS = kDwThreads / TPP;        // pixel columns in flight
        constexpr int ROWG = SLOTS / kTW;              // row groups
        constexpr int R = kTH / ROWG;                  // output rows per thread
        constexpr int IRN = R + 2 * DIL;               // input rows one thread walks
        static_assert(SLOTS % kTW == 0 && ROWG >= 1 && kTH % ROWG == 0, "bad depthwise mapping");
        const int t = threadIdx.x - kDwWarp0 * 32;
        const int c4 = t % TPP, slot = t / TPP;
        const int col = slot % kTW, rg = slot / kTW;
        // this thread's 8 bytes inside a swizzled operand row: 16-byte chunk (c4 >> 1) XOR the row's swizzle phase;
        // row m = (rg*R + r)*16 + col, so the phase depends on col only (128B: m & 7, 64B: (m >> 1) & 3)
        const int swz = KSWZ == 128 ? (col & 7) : ((col >> 1) & 3);
        const uint32_t a_off = (uint32_t)(((rg * R) * kTW + col) * KSWZ + ((((c4 >> 1) ^ swz) << 4) | ((c4 & 1) << 3)));
        const uint32_t slab_off = (uint32_t)((((rg * R) * L::IW + col) * CB + c4 * 4) * 2);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                float2 acc[R][2];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r][0] = make_float2(0.0f, 0.0f);
                    acc[r][1] = make_float2(0.0f, 0.0f);
                }
                PF_WAIT(PF_DW_SLAB_FULL, mbar_wait_sel(&slab_full[ss], sph, p.park & 4u));
                const long long pf_c0 = pf ? clock64() : 0;
                const uint32_t sbase = smem_u32(smem + L::kSlabOff + ss * L::kSlabBytes);
                const uint32_t tbase = sbase + slab_off;
                float2 wt[9][2];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const uint4 v = lds_u128(sbase + (uint32_t)(L::kSlabPix + (k * CB + c4 * 4) * 4));
                    wt[k][0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
                    wt[k][1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
                }
#pragma unroll
                for (int ir = 0; ir < IRN; ++ir) {
                    float2 v[3][2];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const uint2 raw = lds_u64(tbase + (uint32_t)((ir * L::IW + dx * DIL) * CB * 2));
                        v[dx][0] = unpack_bf16x2_prmt(raw.x);
                        v[dx][1] = unpack_bf16x2_prmt(raw.y);
                    }
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const int r = ir - dy * DIL;
                        if (r >= 0 && r < R) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                acc[r][0] = ffma2(v[dx][0], wt[dy * 3 + dx][0], acc[r][0]);
                                acc[r][1] = ffma2(v[dx][1], wt[dy * 3 + dx][1], acc[r][1]);
                            }
                        }
                    }
                }
                if (pf) pf_acc[PF_DW_COMPUTE] += (unsigned long long)(clock64() - pf_c0);
                PF_WAIT(PF_DW_AB_EMPTY, mbar_wait_sel(&a_empty[sa], aph ^ 1, p.park & 4u));   // the MMAs that read this A stage retired
                const uint32_t abase = smem_u32(smem + L::kAOff + sa * L::kABytes) + a_off;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t lo = pack_bf16x2(acc[r][0].x, acc[r][0].y), hi = pack_bf16x2(acc[r][1].x, acc[r][1].y);
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(abase + (uint32_t)(r * kTW * KSWZ)), "r"(lo),
                                 "r"(hi)
                                 : "memory");
                }
                fence_proxy_async();                                   // writer-side proxy fence (see the 64-channel form above)
                __syncwarp();                                          // orders the lanes' stores before lane 0's release
                if (lane == 0) {
                    mbar_arrive(&a_full[sa]);
                    mbar_arrive(&slab_empty[ss]);                      // slab released after the stores that consumed it (see above)
                }
                if (++ss == SS) { ss = 0; sph ^= 1; }
                if (++sa == SA) { sa = 0; aph ^= 1; }
            }
        }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue =====================
        const int q = (warp - kEpiWarp0) & 3;                 // TMEM lane quarter == warp % 4
        const int grp = (warp - kEpiWarp0) >> 2;              // NH = 1: accumulator stage; NH = 2: column half
        unsigned char *stg = smem + L::kStagingOff + (grp * 4 + q) * L::kStagingBytes;
        int local_tile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            if (NH == 1 && (local_tile & 1) != grp) continue;
            const uint32_t acc_phase = NH == 1 ? (uint32_t)((local_tile >> 1) & 1) : (uint32_t)(local_tile & 1);
            const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
            const int b = m_tile / tiles_per_img;
            const int t2 = m_tile - b * tiles_per_img;
            const int ty = t2 / p.tiles_x, tx = t2 - ty * p.tiles_x;
            const int n0 = (NH == 1 ? n_tile : n_tile * NH + grp) * BLOCK_N;
            const int py = ty * kTH + q * 2 + (lane >> 4), px = tx * kTW + (lane & 15);
            const bool row_ok = py < p.H && px < p.W;
            const size_t pix = ((size_t)b * p.H + py) * p.W + px;
            PF_WAIT(PF_EPI_TMEM_FULL, mbar_wait_sel(&tmem_full[grp], acc_phase, p.park & 8u));
            const long long pf_e0 = pf ? clock64() : 0;
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * BLOCK_N);
            // residual rows (tf.add) are fetched one 32-column group ahead so their latency hides behind the previous group
            // every lane reads a residual row (lanes outside the map read the clamped border pixel; their outputs are
            // clipped by the TMA store): the loads stay warp-uniform, so the .sync.aligned tensor-memory loads below never
            // run in a diverged warp
            const size_t pixc = ((size_t)b * p.H + min(py, p.H - 1)) * p.W + min(px, p.W - 1);
            const uint4 *rbase = p.residual ? reinterpret_cast<const uint4 *>(p.residual + pixc * p.ldr + n0) : nullptr;
            uint4 rnext[4] = {};
            if (rbase) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rnext[j] = rbase[j];
            }
            if constexpr (L::kWide) {
                if (p.wide64 && !p.residual && (p.act == LWOP_ACT_RELU || p.act == LWOP_ACT_ELU)) {
                    // 64 output channels per iteration: ONE 128B-swizzled staging tile and ONE TMA store per warp -- half
                    // the load / fence / store round trips of the 32-column form, which is what the epilogue-bound
                    // layers (first layer: 1.1 GB of output; the ELU layers) spend their time on.
                    const bool elu = p.act == LWOP_ACT_ELU;
#pragma unroll 1
                    for (int g = 0; g < BLOCK_N / 64; ++g) {
                        uint32_t ra[32], rb[32];
                        tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64), ra);
                        tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64 + 32), rb);
                        tmem_ld_wait();
                        if (g == BLOCK_N / 64 - 1) {              // accumulator drained: hand it back to the MMA warp
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[grp]);
                        }
                        if (lane == 0) PF_WAIT(PF_EPI_STORE_WAIT, bulk_wait_read_all());
                        __syncwarp();
                        const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s + n0 + g * 64);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t *r = j < 4 ? &ra[j * 8] : &rb[(j - 4) * 8];
                            const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                            float2 s0 = fadd2(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), make_float2(b0.x, b0.y));
                            float2 s1 = fadd2(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), make_float2(b0.z, b0.w));
                            float2 s2 = fadd2(make_float2(__uint_as_float(r[4]), __uint_as_float(r[5])), make_float2(b1.x, b1.y));
                            float2 s3 = fadd2(make_float2(__uint_as_float(r[6]), __uint_as_float(r[7])), make_float2(b1.z, b1.w));
                            uint4 o;
                            if (elu) {
                                s0 = elu2_fast(s0); s1 = elu2_fast(s1); s2 = elu2_fast(s2); s3 = elu2_fast(s3);
                                o.x = pack_bf16x2(s0.x, s0.y); o.y = pack_bf16x2(s1.x, s1.y);
                                o.z = pack_bf16x2(s2.x, s2.y); o.w = pack_bf16x2(s3.x, s3.y);
                            } else {
                                o.x = pack_relu_bf16x2(s0.x, s0.y); o.y = pack_relu_bf16x2(s1.x, s1.y);
                                o.z = pack_relu_bf16x2(s2.x, s2.y); o.w = pack_relu_bf16x2(s3.x, s3.y);
                            }
                            // 128B swizzle: 16-byte chunk index XOR (row & 7)
                            *reinterpret_cast<uint4 *>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_4d(&map_out64, stg, p.out_col0 + n0 + g * 64, tx * kTW, ty * kTH + q * 2, b);
                            bulk_commit_group();
                        }
                    }
                    if (pf) pf_acc[PF_EPI_WORK] += (unsigned long long)(clock64() - pf_e0);
                    continue;
                }
            }
#pragma unroll 1
            for (int g = 0; g < BLOCK_N / 32; ++g) {
                uint32_t r0[32];
                tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 32), r0);
                uint4 rr[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) rr[j] = rnext[j];
                if (rbase && g + 1 < BLOCK_N / 32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) rnext[j] = rbase[(g + 1) * 4 + j];
                }
                tmem_ld_wait();
                if (g == BLOCK_N / 32 - 1) {                  // accumulator drained: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[grp]);
                }
                if (lane == 0) PF_WAIT(PF_EPI_STORE_WAIT, bulk_wait_read_all());   // previous store has finished reading the staging tile
                __syncwarp();
                const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s + n0 + g * 32);
                if (p.act == LWOP_ACT_RELU && !p.residual) {
                    // conv_dw layers (folded BN + ReLU): packed bias add, then ReLU + bf16 pack in one instruction
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                        const float2 s0 = fadd2(make_float2(__uint_as_float(r0[j * 8 + 0]), __uint_as_float(r0[j * 8 + 1])),
                                                make_float2(b0.x, b0.y));
                        const float2 s1 = fadd2(make_float2(__uint_as_float(r0[j * 8 + 2]), __uint_as_float(r0[j * 8 + 3])),
                                                make_float2(b0.z, b0.w));
                        const float2 s2 = fadd2(make_float2(__uint_as_float(r0[j * 8 + 4]), __uint_as_float(r0[j * 8 + 5])),
                                                make_float2(b1.x, b1.y));
                        const float2 s3 = fadd2(make_float2(__uint_as_float(r0[j * 8 + 6]), __uint_as_float(r0[j * 8 + 7])),
                                                make_float2(b1.z, b1.w));
                        uint4 o;
                        o.x = pack_relu_bf16x2(s0.x, s0.y); o.y = pack_relu_bf16x2(s1.x, s1.y);
                        o.z = pack_relu_bf16x2(s2.x, s2.y); o.w = pack_relu_bf16x2(s3.x, s3.y);
                        *reinterpret_cast<uint4 *>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o;
                    }
                } else if (p.act == LWOP_ACT_ELU) {
                    // conv_dw_no_bn layers (bias + ELU, the last one + tf.add): packed bias add, branch-free packed ELU
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                        float2 s0 = elu2_fast(fadd2(make_float2(__uint_as_float(r0[j * 8 + 0]), __uint_as_float(r0[j * 8 + 1])),
                                                    make_float2(b0.x, b0.y)));
                        float2 s1 = elu2_fast(fadd2(make_float2(__uint_as_float(r0[j * 8 + 2]), __uint_as_float(r0[j * 8 + 3])),
                                                    make_float2(b0.z, b0.w)));
                        float2 s2 = elu2_fast(fadd2(make_float2(__uint_as_float(r0[j * 8 + 4]), __uint_as_float(r0[j * 8 + 5])),
                                                    make_float2(b1.x, b1.y)));
                        float2 s3 = elu2_fast(fadd2(make_float2(__uint_as_float(r0[j * 8 + 6]), __uint_as_float(r0[j * 8 + 7])),
                                                    make_float2(b1.z, b1.w)));
                        if (rbase) {
                            s0 = fadd2(s0, make_float2(bf16_lo(rr[j].x), bf16_hi(rr[j].x)));
                            s1 = fadd2(s1, make_float2(bf16_lo(rr[j].y), bf16_hi(rr[j].y)));
                            s2 = fadd2(s2, make_float2(bf16_lo(rr[j].z), bf16_hi(rr[j].z)));
                            s3 = fadd2(s3, make_float2(bf16_lo(rr[j].w), bf16_hi(rr[j].w)));
                        }
                        uint4 o;
                        o.x = pack_bf16x2(s0.x, s0.y); o.y = pack_bf16x2(s1.x, s1.y);
                        o.z = pack_bf16x2(s2.x, s2.y); o.w = pack_bf16x2(s3.x, s3.y);
                        *reinterpret_cast<uint4 *>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o;
                    }
                } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {                 // 8 columns = one 16-byte chunk
                    float v[8];
                    const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = apply_act_fast(__uint_as_float(r0[j * 8 + e]) + bb[e], p.act);
                    if (rbase) {
                        v[0] += bf16_lo(rr[j].x); v[1] += bf16_hi(rr[j].x); v[2] += bf16_lo(rr[j].y); v[3] += bf16_hi(rr[j].y);
                        v[4] += bf16_lo(rr[j].z); v[5] += bf16_hi(rr[j].z); v[6] += bf16_lo(rr[j].w); v[7] += bf16_hi(rr[j].w);
                    }
                    uint4 o;
                    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                    // 64B swizzle: 16-byte chunk index XOR ((row >> 1) & 3)
                    *reinterpret_cast<uint4 *>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o;
                }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    // box = (32 channels, 16 columns, 2 rows, 1 image): rows of the staging tile in the same order
                    tma_store_4d(&map_out, stg, p.out_col0 + n0 + g * 32, tx * kTW, ty * kTH + q * 2, b);
                    bulk_commit_group();
                }
            }
            if (pf) pf_acc[PF_EPI_WORK] += (unsigned long long)(clock64() - pf_e0);
        }
        if (lane == 0) bulk_wait_all();
    }
    if (pf && lane == 0 && (warp == 0 || warp == 1 || warp == 3 || warp == kDwWarp0 || warp == kEpiWarp0)) {
        if (warp == 0) pf_acc[PF_TOTAL] = (unsigned long long)(clock64() - pf_start);
#pragma unroll
        for (int i = 0; i < (PROF ? PF_N : 1); ++i)
            if (pf_acc[i]) atomicAdd(p.prof + i, pf_acc[i]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------
// CTA-pair variant with ONE tcgen05.mma per pair (cta_group::2) for Cout = 512: the two CTAs of a cluster process
// two DIFFERENT pixel patches (M = 256 across the pair) and split every [512 x 64] weight chunk between them:
// each CTA streams only 128 of the 256 rows of each N half, the tensor cores of both SMs read both halves.  This
// halves the weight traffic L2 -> SM per output row, which is what bounds the single-CTA form (every 128-pixel
// tile re-streams the 512 KB weight matrix).  Each CTA runs its own slab loads, depthwise producers and epilogue;
// the leader CTA's MMA thread issues for the pair once BOTH CTAs' A chunks are ready (each CTA's forwarder thread
// fences its depthwise warps' stores into the async proxy, then arrives on the leader's barrier).
// ---------------------------------------------------------------------------
template <int DIL, int SA_, int SB, int SS>
struct SepPair2Smem {
    static constexpr int SA = SA_;                              // >= kKGroup: a K group of 4 chunks is re-read by N half 1
    static constexpr int kABytes = 128 * 128;
    static constexpr int kBBytes = 128 * 128;                   // this CTA's 128 rows of one N half of a weight chunk
    static constexpr int IW = kTW + 2 * DIL, IH = kTH + 2 * DIL;
    static constexpr int kSlabPix = IW * IH * 64 * 2;
    static constexpr int kTapBytes = 9 * 64 * 4;
    static constexpr int kSlabTx = kSlabPix + kTapBytes;
    static constexpr int kSlabBytes = ((kSlabTx + 127) / 128) * 128;
    static constexpr int kAOff = 0;
    static constexpr int kBOff = kAOff + SA * kABytes;
    static constexpr int kStagingOff = kBOff + SB * kBBytes;
    static constexpr int kSlabOff = kStagingOff + kEpiWarps * 4096;
    static constexpr int kBarOff = kSlabOff + SS * kSlabBytes;
    static constexpr int kNumBars = 2 * SS + 3 * SA + 2 * SB + 4;
    static constexpr int kBiasOff = ((kBarOff + kNumBars * 8 + 8 + 15) / 16) * 16;
    static constexpr int kTotal = kBiasOff + 512 * 4;
    static_assert(kTotal <= 232448, "shared memory budget exceeded");
};

template <int DIL, int SA_, int SB, int SS, bool PROF>
__global__ void __launch_bounds__(kSepThreads, 1)
sep_pair2_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_taps,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                 const SepParams p) {
    using L = SepPair2Smem<DIL, SA_, SB, SS>;
    constexpr int KG = 4;                                             // chunks per K group
    constexpr int CB = 64, KSWZ = 128, BLOCK_N = 256, SA = L::SA;
    constexpr uint32_t kIdesc = make_idesc_bf16(256, BLOCK_N);        // M = 256 across the pair

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("lwop: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t *slab_full = reinterpret_cast<uint64_t *>(smem + L::kBarOff);
    uint64_t *slab_empty = slab_full + SS;
    uint64_t *a_ready = slab_empty + SS;          // [SA] local: depthwise warps -> forwarder
    uint64_t *a_full = a_ready + SA;              // [SA] leader: the forwarders of both CTAs
    uint64_t *a_empty = a_full + SA;              // [SA] local: multicast commit
    uint64_t *b_full = a_empty + SA;              // [SB] leader: bytes of both CTAs' loads
    uint64_t *b_empty = b_full + SB;              // [SB] local: multicast commit
    uint64_t *tmem_full = b_empty + SB;           // [2]  local: multicast commit (half h)
    uint64_t *tmem_empty = tmem_full + 2;         // [2]  leader: epilogue warps of both CTAs
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *bias_s = reinterpret_cast<float *>(smem + L::kBiasOff);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
    constexpr bool pf = PROF;
    unsigned long long pf_acc[PROF ? PF_N : 1];
#pragma unroll
    for (int i = 0; i < (PROF ? PF_N : 1); ++i) pf_acc[i] = 0;
    const long long pf_start = PROF ? clock64() : 0;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_pair_tiles = (p.num_m_tiles + 1) >> 1;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int kbn = p.k_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_in);
        tma_prefetch_desc(&map_taps);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SS; ++i) {
            mbar_init(&slab_full[i], 1);
            mbar_init(&slab_empty[i], kDwWarps);
        }
        for (int i = 0; i < SA; ++i) {
            mbar_init(&a_ready[i], kDwWarps);
            mbar_init(&a_full[i], 2);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < SB; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 16);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2sm<512>(tmem_ptr);
    for (int i = threadIdx.x; i < 512; i += kSepThreads) bias_s[i] = p.bias[i];
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();                  // the previous kernel of the chain has completed: its output may be read, our output written
    pdl_launch_dependents();     // the next kernel may start its prologue while this one runs

    // this CTA's patch of pair tile pt (the last pair may hold a dummy patch: b == B, loads zero-fill, stores clip)
    auto my_tile = [&](int pt, int &b, int &ty, int &tx) {
        const int tile = pt * 2 + rank;
        b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        ty = t2 / p.tiles_x;
        tx = t2 - ty * p.tiles_x;
    };

    if (warp == 0) {
        // ===================== TMA: this CTA's 128 rows of each N half of every weight chunk =====================
        if (lane == 0) {
            int sb = 0;
            uint32_t bph = 0;
            for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs)
                for (int g = 0; g < kbn; g += KG)
                    for (int h = 0; h < 2; ++h)
                        for (int kb = g; kb < g + KG; ++kb) {
                            PF_WAIT(PF_TMA_AB_EMPTY, mbar_wait_sel(&b_empty[sb], bph ^ 1, p.park & 1u));
                            if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * L::kBBytes);
                            tma_load_2d_2sm(smem + L::kBOff + sb * L::kBBytes, &map_b, &b_full[sb], kb * CB, h * BLOCK_N + rank * 128);
                            if (++sb == SB) { sb = 0; bph ^= 1; }
                        }
        }
    } else if (warp == 3) {
        // ===================== TMA: slabs + taps of this CTA's patch =====================
        if (lane == 0) {
            int ss = 0;
            uint32_t sph = 0;
            for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
                int b, ty, tx;
                my_tile(pt, b, ty, tx);
                for (int kb = 0; kb < kbn; ++kb) {
                    PF_WAIT(PF_TMA_SLAB_EMPTY, mbar_wait_sel(&slab_empty[ss], sph ^ 1, p.park & 1u));
                    unsigned char *slab = smem + L::kSlabOff + ss * L::kSlabBytes;
                    mbar_arrive_expect_tx(&slab_full[ss], L::kSlabTx);
                    tma_load_4d(slab, &map_in, &slab_full[ss], kb * CB, tx * kTW - DIL, ty * kTH - DIL, b);
                    tma_load_2d(slab + L::kSlabPix, &map_taps, &slab_full[ss], kb * CB, 0);
                    if (++ss == SS) { ss = 0; sph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== forwarder: own A chunk ready -> visible to the async proxy -> leader's barrier =====================
        if (lane == 0) {
            int sa = 0;
            uint32_t aph = 0;
            for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs)
                for (int kb = 0; kb < kbn; ++kb) {
                    mbar_wait_sel(&a_ready[sa], aph, p.park & 1u);
                    fence_proxy_async();
                    mbar_arrive_leader(&a_full[sa]);
                    if (++sa == SA) { sa = 0; aph ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, for the pair) =====================
        if (leader) {
            // K runs in groups of SA chunks; within a group N half 0 is issued for all SA chunks, then N half 1 re-reads
            // the same A stages.  Half 0 of a tile therefore completes SA*4 MMAs before half 1 and half 0 of the next
            // tile starts while half 1 is still being drained: the epilogue overlaps the MMAs although the two halves
            // fill all 512 TMEM columns.
            int sb = 0, sa0 = 0;                    // sa0 / aph0: A ring position of the first chunk of the current K group
            uint32_t aph0 = 0, bph = 0;
            int local_tile = 0;
            for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs, ++local_tile) {
                const uint32_t acc_phase = (uint32_t)(local_tile & 1);
                for (int g = 0; g < kbn; g += KG) {
                    for (int h = 0; h < 2; ++h) {
                        for (int c = 0; c < KG; ++c) {
                            const int kb = g + c;
                            int sa = sa0 + c;
                            uint32_t aph = aph0;
                            if (sa >= SA) { sa -= SA; aph ^= 1; }
                            if (kb == 0) PF_WAIT(PF_MMA_TMEM_EMPTY, mbar_wait_sel(&tmem_empty[h], acc_phase ^ 1, p.park & 2u));
                            if (h == 0) PF_WAIT(PF_MMA_A_FULL, mbar_wait_sel(&a_full[sa], aph, p.park & 2u));
                            PF_WAIT(PF_MMA_B_FULL, mbar_wait_sel(&b_full[sb], bph, p.park & 2u));
                            tcgen05_fence_after();
                            const uint32_t a_addr = smem_u32(smem + L::kAOff + sa * L::kABytes);
                            const uint32_t b_addr = smem_u32(smem + L::kBOff + sb * L::kBBytes);
                            if (elect_one()) {
                                const uint64_t da = make_kmajor_desc<KSWZ>(a_addr);
                                const uint64_t db = make_kmajor_desc<KSWZ>(b_addr);
#pragma unroll
                                for (int k = 0; k < CB / 16; ++k)
                                    umma_bf16_ss_2sm(tmem_base + (uint32_t)(h * BLOCK_N), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2),
                                                     kIdesc, (kb | k) != 0 ? 1u : 0u);
                                umma_commit_2sm(&b_empty[sb], 3);
                                if (h == 1) umma_commit_2sm(&a_empty[sa], 3);
                                if (kb == kbn - 1) umma_commit_2sm(&tmem_full[h], 3);
                            }
                            __syncwarp();
                            if (++sb == SB) { sb = 0; bph ^= 1; }
                        }
                    }
                    sa0 += KG;
                    if (sa0 >= SA) { sa0 -= SA; aph0 ^= 1; }
                }
            }
        }
    } else if (warp >= kDwWarp0 && warp < kDwWarp0 + kDwWarps) {
        // ===================== depthwise producers (own patch, every chunk) =====================
        constexpr int IB = 4 + 2 * DIL;
        const int wq = warp - kDwWarp0;
        const int bx = wq & 3, by = wq >> 2;
        const uint32_t slab_off = (uint32_t)((((by * 4) * L::IW + bx * 4) * CB + lane * 2) * 2);
        const uint32_t tap_off = (uint32_t)(L::kSlabPix + lane * 8);
        uint32_t a_col[4];
#pragma unroll
        for (int ox = 0; ox < 4; ++ox)
            a_col[ox] = (uint32_t)(((by * 4) * kTW + bx * 4 + ox) * 128 +
                                   ((((lane >> 2) ^ (((bx & 1) << 2) | ox)) << 4) | ((lane & 3) << 2)));
        int sa = 0, ss = 0;
        uint32_t aph = 0, sph = 0;
        for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
            for (int kb = 0; kb < kbn; ++kb) {
                float2 acc[4][4];
#pragma unroll
                for (int oy = 0; oy < 4; ++oy)
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox) acc[oy][ox] = make_float2(0.0f, 0.0f);
                PF_WAIT(PF_DW_SLAB_FULL, mbar_wait_sel(&slab_full[ss], sph, p.park & 4u));
                const long long pf_c0 = pf ? clock64() : 0;
                const uint32_t sbase = smem_u32(smem + L::kSlabOff + ss * L::kSlabBytes);
                const uint32_t tbase = sbase + slab_off;
                float2 wt[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const uint2 v = lds_u64(sbase + tap_off + (uint32_t)(k * CB * 4));
                    wt[k] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
                }
#pragma unroll
                for (int iy = 0; iy < IB; ++iy) {
                    float2 v[IB];
#pragma unroll
                    for (int ix = 0; ix < IB; ++ix) {
                        uint32_t raw;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw) : "r"(tbase + (uint32_t)((iy * L::IW + ix) * CB * 2)));
                        v[ix] = unpack_bf16x2_prmt(raw);
                    }
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const int oy = iy - dy * DIL;
                        if (oy >= 0 && oy < 4) {
#pragma unroll
                            for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx)
                                    acc[oy][ox] = ffma2(v[ox + dx * DIL], wt[dy * 3 + dx], acc[oy][ox]);
                        }
                    }
                }
                // (the slab is released after the operand stores that consume every load of it: see sep_conv_kernel)
                if (pf) pf_acc[PF_DW_COMPUTE] += (unsigned long long)(clock64() - pf_c0);
                PF_WAIT(PF_DW_AB_EMPTY, mbar_wait_sel(&a_empty[sa], aph ^ 1, p.park & 4u));
                const uint32_t abase = smem_u32(smem + L::kAOff + sa * L::kABytes);
#pragma unroll
                for (int oy = 0; oy < 4; ++oy)
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox)
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(abase + a_col[ox] + (uint32_t)(oy * kTW * 128)),
                                     "r"(pack_bf16x2(acc[oy][ox].x, acc[oy][ox].y))
                                     : "memory");
                fence_proxy_async();                   // writer-side proxy fence: own stores -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_ready[sa]);
                    mbar_arrive(&slab_empty[ss]);
                }
                if (++ss == SS) { ss = 0; sph ^= 1; }
                if (++sa == SA) { sa = 0; aph ^= 1; }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue: group h drains N half h of this CTA's 128 rows =====================
        const int q = (warp - kEpiWarp0) & 3;
        const int grp = (warp - kEpiWarp0) >> 2;
        unsigned char *stg = smem + L::kStagingOff + (grp * 4 + q) * 4096;
        int local_tile = 0;
        for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs, ++local_tile) {
            const uint32_t acc_phase = (uint32_t)(local_tile & 1);
            int b, ty, tx;
            my_tile(pt, b, ty, tx);
            const int py = ty * kTH + q * 2 + (lane >> 4), px = tx * kTW + (lane & 15);
            const bool row_ok = b < p.B && py < p.H && px < p.W;
            const size_t pix = ((size_t)b * p.H + py) * p.W + px;
            const long long pf_e0 = pf ? clock64() : 0;
#pragma unroll 1
            for (int gg = 0; gg < 4; ++gg) {
                // N half h = gg / 2 is drained by all 8 warps (this warp: 128 of its 256 columns, 64 per iteration), half 0
                // first: it completes one K group earlier than half 1 and the next tile's MMAs for it start as soon as it is empty
                const int h = gg >> 1, g = gg & 1;
                const int n0 = h * BLOCK_N + grp * (BLOCK_N / 2) + g * 64;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0;
                if (g == 0) {
                    PF_WAIT(PF_EPI_TMEM_FULL, mbar_wait_sel(&tmem_full[h], acc_phase, p.park & 8u));
                    tcgen05_fence_after();
                }
                uint32_t r0[2][32];
                const long long pf_l0 = pf ? clock64() : 0;
                tmem_ld_32x32b_x32(taddr, r0[0]);
                tmem_ld_32x32b_x32(taddr + 32u, r0[1]);
                const uint4 *rp = (p.residual && row_ok) ? reinterpret_cast<const uint4 *>(p.residual + pix * p.ldr + n0) : nullptr;
                tmem_ld_wait();
                const long long pf_m0 = pf ? clock64() : 0;
                if (pf) pf_acc[PF_EPI_LDTM] += (unsigned long long)(pf_m0 - pf_l0);
                if (g == 1) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tmem_empty[h]);
                }
                uint4 o[2][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s + n0 + u * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                        const uint32_t *r = &r0[u][j * 8];
                        if (p.act == LWOP_ACT_RELU && !p.residual) {
                            const float2 s0 = fadd2(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), make_float2(b0.x, b0.y));
                            const float2 s1 = fadd2(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), make_float2(b0.z, b0.w));
                            const float2 s2 = fadd2(make_float2(__uint_as_float(r[4]), __uint_as_float(r[5])), make_float2(b1.x, b1.y));
                            const float2 s3 = fadd2(make_float2(__uint_as_float(r[6]), __uint_as_float(r[7])), make_float2(b1.z, b1.w));
                            o[u][j].x = pack_relu_bf16x2(s0.x, s0.y); o[u][j].y = pack_relu_bf16x2(s1.x, s1.y);
                            o[u][j].z = pack_relu_bf16x2(s2.x, s2.y); o[u][j].w = pack_relu_bf16x2(s3.x, s3.y);
                        } else {
                            float v[8];
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = apply_act_fast(__uint_as_float(r[e]) + bb[e], p.act);
                            if (rp) {
                                const uint4 rr = rp[u * 4 + j];
                                v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                                v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                            }
                            o[u][j].x = pack_bf16x2(v[0], v[1]); o[u][j].y = pack_bf16x2(v[2], v[3]);
                            o[u][j].z = pack_bf16x2(v[4], v[5]); o[u][j].w = pack_bf16x2(v[6], v[7]);
                        }
                    }
                }
                // the previous iteration's TMA stores must have finished READING the staging tiles; that latency hid behind the math
                if (pf) pf_acc[PF_EPI_MATH] += (unsigned long long)(clock64() - pf_m0);
                if (lane == 0) PF_WAIT(PF_EPI_STORE_WAIT, bulk_wait_read_all());
                __syncwarp();
                const long long pf_f0 = pf ? clock64() : 0;
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4 *>(stg + u * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o[u][j];
                fence_proxy_async();
                __syncwarp();
                if (pf) pf_acc[PF_EPI_FENCE] += (unsigned long long)(clock64() - pf_f0);
                if (lane == 0) {
                    tma_store_4d(&map_out, stg, p.out_col0 + n0, tx * kTW, ty * kTH + q * 2, b);
                    tma_store_4d(&map_out, stg + 2048, p.out_col0 + n0 + 32, tx * kTW, ty * kTH + q * 2, b);
                    bulk_commit_group();
                }
            }
            if (pf) pf_acc[PF_EPI_WORK] += (unsigned long long)(clock64() - pf_e0);
        }
        if (lane == 0) bulk_wait_all();
    }
    if (pf && lane == 0 && (warp == 0 || warp == 1 || warp == 3 || warp == kDwWarp0 || warp == kEpiWarp0)) {
        if (warp == 0) pf_acc[PF_TOTAL] = (unsigned long long)(clock64() - pf_start);
#pragma unroll
        for (int i = 0; i < (PROF ? PF_N : 1); ++i)
            if (pf_acc[i]) atomicAdd(p.prof + i, pf_acc[i]);
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm<512>(tmem_base);
    }
}

}  // namespace

struct SepPlan {
    CUtensorMap map_in, map_taps, map_b, map_out, map_out64;   // map_out64: (64 channels, 16, 2, 1) box, 128B swizzle (Cout_pad == 64)
    SepParams p;
    int block_n, cb, dil, nh;
    int pair;          // CTA-pair kernel (Cout = 512)
    int grid;
};

namespace {

template <int BLOCK_N, int CB, int DIL, int SA, int SB, int SS, int NH>
cudaError_t launch_sep_t(const SepPlan *pl, cudaStream_t s) {
    constexpr int smem = SepSmem<BLOCK_N, CB, DIL, SA, SB, SS, NH>::kTotal;
    if (pl->p.prof) {
        auto kfn = sep_conv_kernel<BLOCK_N, CB, DIL, SA, SB, SS, NH, true>;
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        return launch_chain(kfn, (unsigned)pl->grid, kSepThreads, smem, s, 1, pl->map_in, pl->map_taps, pl->map_b, pl->map_out, pl->map_out64, pl->p);
    }
    auto kfn = sep_conv_kernel<BLOCK_N, CB, DIL, SA, SB, SS, NH, false>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return launch_chain(kfn, (unsigned)pl->grid, kSepThreads, smem, s, 1, pl->map_in, pl->map_taps, pl->map_b, pl->map_out, pl->map_out64, pl->p);
}

}  // namespace

bool sep_plan_supported(int C, int Cout_pad, int stride, int dil) {
    if (stride != 1 || (dil != 1 && dil != 2)) return false;
    if (!(C == 32 || C % 64 == 0)) return false;
    if (C == 32) return Cout_pad == 64 && dil == 1;
    if (Cout_pad % 256 == 0) return true;
    return Cout_pad == 128 && dil == 1;
}

SepPlan *sep_plan_create(const SepConvDesc &d, char *err, size_t errlen) {
    if (!gemm_driver_ready(err, errlen)) return nullptr;
    if (!sep_plan_supported(d.C, d.Cout_pad, 1, d.dil)) {
        snprintf(err, errlen, "fused separable conv: unsupported C=%d Cout_pad=%d dil=%d", d.C, d.Cout_pad, d.dil);
        return nullptr;
    }
    if (!d.out || (d.ldo % 8) != 0 || (d.out_col0 % 8) != 0) {
        snprintf(err, errlen, "fused separable conv: output pitch / column offset must be multiples of 8");
        return nullptr;
    }
    SepPlan *pl = (SepPlan *)calloc(1, sizeof(SepPlan));
    pl->cb = d.C == 32 ? 32 : 64;
    pl->block_n = d.Cout_pad % 256 == 0 ? 256 : d.Cout_pad;
    pl->dil = d.dil;
    SepParams &p = pl->p;
    {
        static const char *pk = getenv("LWOP_SEP_PARK");
        p.park = pk ? (unsigned)atoi(pk) : 1u;
        static const char *w64 = getenv("LWOP_SEP_WIDE64");
        p.wide64 = w64 ? atoi(w64) : 1;
    }
    p.dw_w = d.dw_w;
    p.bias = d.bias;
    p.residual = (const __nv_bfloat16 *)d.residual;
    p.B = d.B; p.H = d.H; p.W = d.W; p.C = d.C;
    p.Cout_pad = d.Cout_pad;
    p.tiles_x = (d.W + kTW - 1) / kTW;
    p.tiles_y = (d.H + kTH - 1) / kTH;
    p.num_m_tiles = d.B * p.tiles_x * p.tiles_y;
    // Cout = 512: both N = 256 halves in one pass over the patch (LWOP_SEP_NH1=1 keeps two N tiles, for A/B timing)
    pl->nh = (d.Cout_pad == 512 && !getenv("LWOP_SEP_NH1")) ? 2 : 1;
    p.num_n_tiles = d.Cout_pad / (pl->block_n * pl->nh);
    p.k_blocks = d.C / pl->cb;
    p.act = d.act;
    p.ldr = d.ldr;
    p.out_col0 = d.out_col0;
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    int sms = gemm_num_sms();
    if (sms <= 0) sms = 148;
    pl->grid = tiles < sms ? tiles : sms;
    // CTA-pair kernel for Cout = 512 (LWOP_SEP_PAIR=0 keeps the single-CTA NH = 2 form)
    const char *pe = getenv("LWOP_SEP_PAIR");
    // Cout = 512: CTA pairs with cta_group::2 MMAs (sep_pair2_kernel) unless LWOP_SEP_PAIR=0 (single-CTA NH = 2 kernel)
    pl->pair = (pl->nh == 2 && (pe ? atoi(pe) != 0 : true) && p.k_blocks % 4 == 0 && p.num_m_tiles >= 2) ? 1 : 0;
    if (pl->pair) {
        const int pair_tiles = (p.num_m_tiles + 1) / 2;
        const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
        pl->grid = 2 * pairs;
    }
    {   // input patch + halo: box (CB channels, 16 + 2d columns, 8 + 2d rows, 1 image), zero OOB fill
        cuuint64_t dims[4] = {(cuuint64_t)d.C, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.C * 2, (cuuint64_t)d.W * d.C * 2, (cuuint64_t)d.H * d.W * d.C * 2};
        cuuint32_t box[4] = {(cuuint32_t)pl->cb, (cuuint32_t)(kTW + 2 * d.dil), (cuuint32_t)(kTH + 2 * d.dil), 1};
        if (!encode_tiled_bf16(&pl->map_in, 4, d.in, dims, strides, box, 0, err, errlen)) {
            free(pl);
            return nullptr;
        }
    }
    {   // depthwise taps [9][C] fp32: the chunk's (CB, 9) block rides along with the slab
        cuuint64_t dims[2] = {(cuuint64_t)d.C, 9};
        cuuint64_t strides[1] = {(cuuint64_t)d.C * 4};
        cuuint32_t box[2] = {(cuuint32_t)pl->cb, 9};
        if (!encode_tiled_f32(&pl->map_taps, 2, d.dw_w, dims, strides, box, err, errlen)) {
            free(pl);
            return nullptr;
        }
    }
    {   // pointwise weights [Cout_pad][C] bf16, K-major tiles of (CB, BLOCK_N)
        cuuint64_t dims[2] = {(cuuint64_t)d.C, (cuuint64_t)d.Cout_pad};
        cuuint64_t strides[1] = {(cuuint64_t)d.C * 2};
        cuuint32_t box[2] = {(cuuint32_t)pl->cb, (cuuint32_t)(pl->pair ? 128 : pl->block_n)};
        if (!encode_tiled_bf16(&pl->map_b, 2, d.w_packed, dims, strides, box, 2 * pl->cb, err, errlen)) {
            free(pl);
            return nullptr;
        }
    }
    {   // output NHWC [B,H,W,ldo]: per epilogue warp a (32 channels, 16 columns, 2 rows) box, clipped at the border
        cuuint64_t dims[4] = {(cuuint64_t)d.ldo, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.ldo * 2, (cuuint64_t)d.W * d.ldo * 2, (cuuint64_t)d.H * d.W * d.ldo * 2};
        cuuint32_t box[4] = {32, (cuuint32_t)kTW, 2, 1};
        if (!encode_tiled_bf16(&pl->map_out, 4, d.out, dims, strides, box, 64, err, errlen)) {
            free(pl);
            return nullptr;
        }
        pl->map_out64 = pl->map_out;
        if (pl->nh == 1 && pl->block_n >= 64) {
            cuuint32_t box64[4] = {64, (cuuint32_t)kTW, 2, 1};
            if (!encode_tiled_bf16(&pl->map_out64, 4, d.out, dims, strides, box64, 128, err, errlen)) {
                free(pl);
                return nullptr;
            }
        }
    }
    return pl;
}

void sep_plan_destroy(SepPlan *p) { free(p); }

static cudaError_t sep_plan_launch_inner(const SepPlan *pl, cudaStream_t s);

cudaError_t sep_plan_launch(const SepPlan *pl, cudaStream_t s) {
    static const bool prof = getenv("LWOP_SEP_PROF") != nullptr;
    if (!prof) return sep_plan_launch_inner(pl, s);
    // debugging aid: per-role stall cycles, averaged per CTA, printed to stderr (synchronises)
    static const char *names[PF_N] = {"total", "tma:slab_empty", "tma:ab_empty", "mma:tmem_empty", "mma:b_full",
                                      "mma:a_full", "dw:slab_full", "dw:ab_empty", "dw:compute", "epi:tmem_full",
                                      "epi:store_wait", "epi:work", "epi:ldtm", "epi:math", "epi:fence"};
    unsigned long long *d = nullptr, hbuf[PF_N];
    cudaMalloc(&d, PF_N * sizeof(unsigned long long));
    cudaMemsetAsync(d, 0, PF_N * sizeof(unsigned long long), s);
    SepPlan tmp = *pl;
    tmp.p.prof = d;
    cudaError_t e = sep_plan_launch_inner(&tmp, s);
    cudaStreamSynchronize(s);
    cudaMemcpy(hbuf, d, sizeof(hbuf), cudaMemcpyDeviceToHost);
    cudaFree(d);
    fprintf(stderr, "lwop sep prof: C=%d N=%d dil=%d HxW=%dx%d B=%d grid=%d |", pl->p.C, pl->p.Cout_pad, pl->dil, pl->p.H,
            pl->p.W, pl->p.B, pl->grid);
    for (int i = 0; i < PF_N; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)hbuf[i] / pl->grid);
    fprintf(stderr, "\n");
    return e;
}

template <int DIL, int SA, int SB, int SS>
static cudaError_t launch_sep_pair2(const SepPlan *pl, cudaStream_t s) {
    auto kfn = pl->p.prof ? sep_pair2_kernel<DIL, SA, SB, SS, true> : sep_pair2_kernel<DIL, SA, SB, SS, false>;
    constexpr int smem = SepPair2Smem<DIL, SA, SB, SS>::kTotal;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return launch_chain(kfn, (unsigned)pl->grid, kSepThreads, smem, s, 2, pl->map_in, pl->map_taps, pl->map_b, pl->map_out, pl->p);
}

static cudaError_t sep_plan_launch_inner(const SepPlan *pl, cudaStream_t s) {
    // ring depths (A, B, slab): (4,3,3) / (5,3,2) / (4,4,2) measured within 2% of each other
    if (pl->pair) return pl->dil == 2 ? launch_sep_pair2<2, 4, 3, 2>(pl, s) : launch_sep_pair2<1, 4, 3, 3>(pl, s);
    if (pl->cb == 32) return launch_sep_t<64, 32, 1, 2, 3, 6, 1>(pl, s);
    if (pl->block_n == 128) return launch_sep_t<128, 64, 1, 2, 4, 3, 1>(pl, s);
    if (pl->nh == 2) {
        if (pl->dil == 2) return launch_sep_t<256, 64, 2, 2, 3, 2, 2>(pl, s);
        return launch_sep_t<256, 64, 1, 2, 3, 3, 2>(pl, s);
    }
    if (pl->dil == 2) return launch_sep_t<256, 64, 2, 2, 2, 2, 1>(pl, s);
    return launch_sep_t<256, 64, 1, 2, 3, 2, 1>(pl, s);
}

}  // namespace lwop
