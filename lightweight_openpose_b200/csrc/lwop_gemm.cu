// bf16 implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a):
//   D[M, N] = act( sum_taps A_tap[M, Cin] * W_tap[N, Cin]^T + bias ) (+ residual)
// with M = B*H*W output pixels (stride-1 'same' convs), N = Cout.
//
//   * operands staged by TMA into 128B/64B-swizzled shared memory
//       - 1x1 convs: A is the [M, Cin] activation matrix (2D tiled map)
//       - 3x3 convs: A tiles come from the NHWC tensor through an im2col-mode
//         tensor map; the 9 filter taps are 9 K-slices addressed by (offW, offH),
//         'same' zero padding falls out of the map's bounding box + OOB fill
//   * tcgen05.mma (cta_group::1, M=128, N=BLOCK_N, K=16) issued by one thread,
//     fp32 accumulators double-buffered in TMEM
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
//     warps 4-7 epilogue (tcgen05.ld -> bias/activation/residual -> global)
//   * persistent over output tiles (grid = min(tiles, SMs))
//
// Reference ops covered: the pointwise half of tf.layers.separable_conv2d and
// tf.layers.conv2d with kernel 1 or 3 (src/net_factory.py:5,15,25), with the
// folded batch-norm / bias / ReLU / ELU / tf.add epilogues of :9-11,18-19,29,37.
#include <stdlib.h>
#include <string.h>

#include "lwop_kernels.cuh"

namespace lwop {

namespace {

constexpr int kBlockM = 128;
constexpr int kNumThreads = 384;        // 4 control warps + 2 epilogue groups of 4 warps
constexpr int kEpilogueWarp0 = 4;

struct GemmParams {
    const float *bias;              // [Cout_pad]
    const __nv_bfloat16 *residual;  // [M][ldr] or nullptr
    void *out;                      // bf16 or fp32, [M][ldo], column offset out_col0
    float *out2;                    // optional fp32 copy [M][Cout] (narrow heads)
    long long M;
    int H, W;                       // spatial size (im2col decomposition of m)
    int Cout, Cout_pad;
    int k_blocks_per_tap, taps, dil;
    int ldo, out_col0, ldr;
    int out_is_f32, act;
    int num_m_tiles, num_n_tiles;
    int use_tma_store;              // wide bf16 output through shared-memory staging + TMA store
    unsigned park;                  // roles whose barrier waits park the warp (bit 0 TMA, 1 MMA, 3 epilogue), see mbar_wait_park
    int B, tiles_x, tiles_y, num_tiles;   // halo kernel: 16 x 8 pixel tiles per image and in total
};

template <int BLOCK_N, int KSWZ, int STAGES, int MT>
struct SmemLayout {
    static constexpr int kAHalfBytes = kBlockM * KSWZ;           // one 128-row MMA operand
    static constexpr int kABytes = MT * kAHalfBytes;
    static constexpr int kBBytes = BLOCK_N * KSWZ;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kBiasOffset = kBarOffset + 256;
    // epilogue staging for TMA stores: 4 warps x (32 rows x 128 B), 1024-byte aligned (128B swizzle)
    static constexpr int kStagingOffset = ((kBiasOffset + 512 * 4 + 1023) / 1024) * 1024;
    static constexpr int kStagingBytes = BLOCK_N >= 64 ? 8 * 4096 : 0;   // one 4 KB tile per epilogue warp
    // dynamic shared memory of a kernel without static __shared__ starts at the CTA's window base
    // (1024-byte aligned); the kernel traps if that ever fails to hold
    static constexpr int kTotal = kStagingOffset + kStagingBytes;
};

// MT = 128-row MMA tiles per CTA tile (1 or 2): with MT = 2 one weight (B) stage feeds two MMAs, which
// cuts the L2 -> SM traffic per flop of the N = 128 convolutions by a quarter (they are L2-bound).
template <int BLOCK_N, int KSWZ, bool IM2COL, int STAGES, int CLUSTER, int MT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_out, const GemmParams p) {
    using L = SmemLayout<BLOCK_N, KSWZ, STAGES, MT>;
    constexpr int BLOCK_K = KSWZ / 2;                 // bf16 elements per swizzle row
    constexpr int kMmasPerBlock = BLOCK_K / 16;
    constexpr int kTileM = kBlockM * MT;
    constexpr int kAccCols = MT * BLOCK_N;            // TMEM columns of one accumulator stage
    static_assert(2 * kAccCols <= 512, "accumulators do not fit in TMEM");
    constexpr uint32_t kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128
                                   : (2 * kAccCols <= 256) ? 256 : 512;
    constexpr uint32_t kIdesc = make_idesc_bf16(kBlockM, BLOCK_N);

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("lwop: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *bias_s = reinterpret_cast<float *>(smem + L::kBiasOffset);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
    const int k_blocks = p.taps * p.k_blocks_per_tap;
    // A cluster of CLUSTER CTAs works on CLUSTER consecutive M tiles of the same N tile in lock step:
    // every CTA fetches 1/CLUSTER of the weight (B) tile and multicasts it to its peers, so the L2 -> SM
    // traffic for B drops by CLUSTER.  m tiles past the end are dummies (zero-filled loads, clipped stores).
    const int rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / CLUSTER, num_clusters = gridDim.x / CLUSTER;
    const int num_m_groups = (p.num_m_tiles + CLUSTER - 1) / CLUSTER;
    const int num_tiles = num_m_groups * p.num_n_tiles;          // cluster-level work items
    constexpr uint16_t kMask = (uint16_t)((1u << CLUSTER) - 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (p.use_tma_store) tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], CLUSTER);         // one release per CTA that multicasts into this stage
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);            // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr);
    for (int i = threadIdx.x; i < p.Cout_pad; i += kNumThreads) bias_s[i] = p.bias[i];
    tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();               // peers' barriers are initialised before anyone signals them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();                  // the previous kernel of the chain has completed: its output may be read, our output written
    pdl_launch_dependents();     // the next kernel may start its prologue while this one runs

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_group = tile / p.num_n_tiles, n_tile = tile - m_group * p.num_n_tiles;
                const int m_tile = m_group * CLUSTER + rank;
                const long long m0 = (long long)m_tile * kTileM;
                int pn[MT], ph[MT], pw[MT];
#pragma unroll
                for (int hm = 0; hm < MT; ++hm) {
                    pn[hm] = ph[hm] = pw[hm] = 0;
                    if (IM2COL) {
                        const long long mh = m0 + hm * kBlockM;
                        const int hw = p.H * p.W;
                        pn[hm] = (int)(mh / hw);
                        const int rem = (int)(mh - (long long)pn[hm] * hw);
                        ph[hm] = rem / p.W;
                        pw[hm] = rem - ph[hm] * p.W;
                    }
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_sel(&empty_bar[stage], phase ^ 1, p.park & 1u);
                    unsigned char *sa = smem + stage * L::kStageBytes;
                    unsigned char *sb = sa + L::kABytes;
                    mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                    const int tap = kb / p.k_blocks_per_tap;
                    const int c0 = (kb - tap * p.k_blocks_per_tap) * BLOCK_K;
#pragma unroll
                    for (int hm = 0; hm < MT; ++hm) {
                        if (IM2COL) {
                            const int r = tap / 3, s = tap - r * 3;
                            tma_load_im2col_4d(sa + hm * L::kAHalfBytes, &map_a, &full_bar[stage], c0, pw[hm] - p.dil,
                                               ph[hm] - p.dil, pn[hm], (uint16_t)(s * p.dil), (uint16_t)(r * p.dil));
                        } else {
                            tma_load_2d(sa + hm * L::kAHalfBytes, &map_a, &full_bar[stage], c0,
                                        (int)(m0 + hm * kBlockM));
                        }
                    }
                    if (CLUSTER > 1) {
                        constexpr int kSliceRows = BLOCK_N / CLUSTER;
                        tma_load_2d_multicast(sb + rank * kSliceRows * KSWZ, &map_b, &full_bar[stage], kb * BLOCK_K,
                                              n_tile * BLOCK_N + rank * kSliceRows, kMask);
                    } else {
                        tma_load_2d(sb, &map_b, &full_bar[stage], kb * BLOCK_K, n_tile * BLOCK_N);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the (warp-uniform) tile schedule and waits on the barriers; one elected lane issues
        // the MMAs and commits.  Uniform control flow keeps descriptors and barrier addresses in uniform registers,
        // so consecutive tcgen05.mma instructions issue back to back.
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            mbar_wait_sel(&tmem_empty[acc], acc_phase ^ 1, p.park & 2u);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait_sel(&full_bar[stage], phase, p.park & 2u);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                const uint32_t sb = sa + L::kABytes;
                if (elect_one()) {
                    const uint64_t db = make_kmajor_desc<KSWZ>(sb);
#pragma unroll
                    for (int hm = 0; hm < MT; ++hm) {
                        const uint64_t da = make_kmajor_desc<KSWZ>(sa + hm * L::kAHalfBytes);
#pragma unroll
                        for (int k = 0; k < kMmasPerBlock; ++k) {
                            // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the >>4 address field
                            umma_bf16_ss(d_tmem + (uint32_t)(hm * BLOCK_N), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2),
                                         kIdesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    // frees the smem slot when these MMAs retire -- in every CTA that writes into it
                    if (CLUSTER > 1) umma_commit_multicast(&empty_bar[stage], kMask);
                    else umma_commit(&empty_bar[stage]);
                    if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);      // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpilogueWarp0) {
        // ===================== epilogue =====================
        // Two epilogue groups of 4 warps; group g owns accumulator stage g and therefore every other tile
        // of this CTA, so the drain of tile i overlaps the drain of tile i+1 as well as its MMAs.
        const int q = (warp - kEpilogueWarp0) & 3;           // TMEM lane quarter == warp % 4
        const int grp = (warp - kEpilogueWarp0) >> 2;
        const int acc = grp;
        uint32_t acc_phase = 0;
        const bool narrow = (p.Cout != p.Cout_pad) || p.out_is_f32 || (p.out_col0 & 7) || (p.ldo & 7);
        unsigned char *stg = smem + L::kStagingOffset + (grp * 4 + q) * 4096;
        int local_tile = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local_tile) {
            if ((local_tile & 1) != grp) continue;
            const int m_group = tile / p.num_n_tiles, n_tile = tile - m_group * p.num_n_tiles;
            const int m_tile = m_group * CLUSTER + rank;
            const int n0 = n_tile * BLOCK_N;
            mbar_wait_sel(&tmem_full[acc], acc_phase, p.park & 8u);
            tcgen05_fence_after();
#pragma unroll 1
          for (int hm = 0; hm < MT; ++hm) {
            const long long m = (long long)m_tile * kTileM + hm * kBlockM + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + hm * BLOCK_N);
            const bool row_ok = m < p.M;
            if (BLOCK_N >= 64 && p.use_tma_store) {
                // ---- wide bf16 output: 64 columns at a time -> swizzled staging -> one TMA store per warp ----
#pragma unroll 1
                for (int g = 0; g < BLOCK_N / 64; ++g) {
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64), r0);
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64 + 32), r1);
                    tmem_ld_wait();
                    if (g == BLOCK_N / 64 - 1 && hm == MT - 1) {   // accumulator drained: hand it back to the MMA warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    }
                    if (lane == 0) bulk_wait_read_all();     // previous store has finished reading the staging tile
                    __syncwarp();
                    const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s + n0 + g * 64);
                    const uint4 *rp = (p.residual && row_ok)
                                          ? reinterpret_cast<const uint4 *>(p.residual + m * p.ldr + n0 + g * 64)
                                          : nullptr;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {            // 8 columns = one 16-byte chunk
                        float v[8];
                        const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = j * 8 + e;
                            const uint32_t raw = c < 32 ? r0[c] : r1[c - 32];
                            v[e] = apply_act_fast(__uint_as_float(raw) + bb[e], p.act);
                        }
                        if (rp) {
                            const uint4 rr = rp[j];
                            v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                            v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                        }
                        uint4 o;
                        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                        // 128B swizzle: 16-byte chunk index XOR (row & 7)
                        *reinterpret_cast<uint4 *>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_out, stg, p.out_col0 + n0 + g * 64, m_tile * kTileM + hm * kBlockM + q * 32);
                        bulk_commit_group();
                    }
                }
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(taddr + (uint32_t)c, r);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = apply_act_fast(__uint_as_float(r[j]) + bias_s[n0 + c + j], p.act);
                if (row_ok) {
                    if (p.residual) {
                        const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + m * p.ldr + n0 + c);
                        const uint4 r0 = rp[0], r1 = rp[1];
                        v[0] += bf16_lo(r0.x); v[1] += bf16_hi(r0.x); v[2] += bf16_lo(r0.y); v[3] += bf16_hi(r0.y);
                        v[4] += bf16_lo(r0.z); v[5] += bf16_hi(r0.z); v[6] += bf16_lo(r0.w); v[7] += bf16_hi(r0.w);
                        v[8] += bf16_lo(r1.x); v[9] += bf16_hi(r1.x); v[10] += bf16_lo(r1.y); v[11] += bf16_hi(r1.y);
                        v[12] += bf16_lo(r1.z); v[13] += bf16_hi(r1.z); v[14] += bf16_lo(r1.w); v[15] += bf16_hi(r1.w);
                    }
                    if (!narrow) {
                        uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + m * p.ldo +
                                                              p.out_col0 + n0 + c);
                        uint4 o0, o1;
                        o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]);
                        o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
                        o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]);
                        o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
                        op[0] = o0;
                        op[1] = o1;
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = n0 + c + j;
                            if (n < p.Cout) {
                                if (p.out_is_f32)
                                    reinterpret_cast<float *>(p.out)[m * p.ldo + p.out_col0 + n] = v[j];
                                else
                                    reinterpret_cast<__nv_bfloat16 *>(p.out)[m * p.ldo + p.out_col0 + n] =
                                        __float2bfloat16_rn(v[j]);
                                if (p.out2) p.out2[m * p.Cout + n] = v[j];
                            }
                        }
                    }
                }
            }
          }   // hm
            if (!(BLOCK_N >= 64 && p.use_tma_store)) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
            acc_phase ^= 1;                                  // this group's stage is used once per two tiles
        }
        if (lane == 0) bulk_wait_all();                      // outstanding TMA stores complete before exit
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();               // nobody leaves while a peer may still write/signal here
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the N = 128 convolutions: two CTAs of a cluster (one TPC) work on two
// adjacent M tiles with ONE tcgen05.mma of M = 256 per K step.  Each CTA stages its own A rows but only HALF of
// the weight tile (64 of the 128 output channels); the tensor cores of both SMs read both halves.  Per SM this
// halves the weight traffic L2 -> shared memory and the operand reads of B from shared memory -- the two things
// that bound the single-CTA kernel (profiles/r1c_ncu_full_summary.txt: shared-memory data pipe 76 %, tensor 62 %).
// Protocol (as in CUTLASS' sm100 2-SM mainloop): the FULL barriers live in the leader CTA (rank 0) and count the
// bytes of both CTAs' TMA loads; the leader's MMA thread issues for the pair and commits (multicast) to the EMPTY
// and TMEM-FULL barriers of both CTAs; the epilogue warps of both CTAs arrive on the leader's TMEM-EMPTY barrier.
// ---------------------------------------------------------------------------
template <int STAGES, int MT>
struct SmemLayout2Sm {
    static constexpr int kAHalfBytes = kBlockM * 128;
    static constexpr int kABytes = MT * kAHalfBytes;
    static constexpr int kBBytes = 64 * 128;                     // this CTA's half of the [128 x 64] weight chunk
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kBiasOffset = kBarOffset + 256;
    static constexpr int kStagingOffset = ((kBiasOffset + 512 * 4 + 1023) / 1024) * 1024;
    static constexpr int kTotal = kStagingOffset + 8 * 4096;
    static_assert(kTotal <= 232448, "shared memory budget exceeded");
};

template <bool IM2COL, int STAGES, int MT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_conv2sm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_out, const GemmParams p) {
    using L = SmemLayout2Sm<STAGES, MT>;
    constexpr int BLOCK_N = 128, KSWZ = 128, BLOCK_K = 64;
    constexpr int kMmasPerBlock = BLOCK_K / 16;
    constexpr int kTileM = kBlockM * MT;
    constexpr int kAccCols = MT * BLOCK_N;
    constexpr uint32_t kTmemCols = 2 * kAccCols <= 256 ? 256 : 512;
    constexpr uint32_t kIdesc = make_idesc_bf16(256, BLOCK_N);        // M = 256 across the pair

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("lwop: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *bias_s = reinterpret_cast<float *>(smem + L::kBiasOffset);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
    const int k_blocks = p.taps * p.k_blocks_per_tap;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_pair_tiles = (p.num_m_tiles + 1) >> 1;             // two M tiles per work item (N = 128: one N tile)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);              // leader's producer arrival (+ the bytes of both CTAs)
            mbar_init(&empty_bar[i], 1);             // one multicast commit per use
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);             // multicast commit
            mbar_init(&tmem_empty[i], 8);            // leader only: 4 epilogue warps of each CTA
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2sm<kTmemCols>(tmem_ptr);
    for (int i = threadIdx.x; i < p.Cout_pad; i += kNumThreads) bias_s[i] = p.bias[i];
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                              // the peer's barriers exist before anyone signals them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();                  // the previous kernel of the chain has completed: its output may be read, our output written
    pdl_launch_dependents();     // the next kernel may start its prologue while this one runs

    if (warp == 0) {
        // ===================== TMA producer (both CTAs: own A rows, own half of B) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
                const int m_tile = pt * 2 + rank;
                const long long m0 = (long long)m_tile * kTileM;
                int pn[MT], ph[MT], pw[MT];
#pragma unroll
                for (int hm = 0; hm < MT; ++hm) {
                    pn[hm] = ph[hm] = pw[hm] = 0;
                    if (IM2COL) {
                        const long long mh = m0 + hm * kBlockM;
                        const int hw = p.H * p.W;
                        pn[hm] = (int)(mh / hw);
                        const int rem = (int)(mh - (long long)pn[hm] * hw);
                        ph[hm] = rem / p.W;
                        pw[hm] = rem - ph[hm] * p.W;
                    }
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_sel(&empty_bar[stage], phase ^ 1, p.park & 1u);
                    unsigned char *sa = smem + stage * L::kStageBytes;
                    unsigned char *sb = sa + L::kABytes;
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                    const int tap = kb / p.k_blocks_per_tap;
                    const int c0 = (kb - tap * p.k_blocks_per_tap) * BLOCK_K;
#pragma unroll
                    for (int hm = 0; hm < MT; ++hm) {
                        if (IM2COL) {
                            const int r = tap / 3, s = tap - r * 3;
                            tma_load_im2col_4d_2sm(sa + hm * L::kAHalfBytes, &map_a, &full_bar[stage], c0, pw[hm] - p.dil,
                                                   ph[hm] - p.dil, pn[hm], (uint16_t)(s * p.dil), (uint16_t)(r * p.dil));
                        } else {
                            tma_load_2d_2sm(sa + hm * L::kAHalfBytes, &map_a, &full_bar[stage], c0, (int)(m0 + hm * kBlockM));
                        }
                    }
                    tma_load_2d_2sm(sb, &map_b, &full_bar[stage], kb * BLOCK_K, rank * 64);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only, for the pair) =====================
        if (leader) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
                mbar_wait_sel(&tmem_empty[acc], acc_phase ^ 1, p.park & 2u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_sel(&full_bar[stage], phase, p.park & 2u);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t sb = sa + L::kABytes;
                    if (elect_one()) {
                        const uint64_t db = make_kmajor_desc<KSWZ>(sb);
#pragma unroll
                        for (int hm = 0; hm < MT; ++hm) {
                            const uint64_t da = make_kmajor_desc<KSWZ>(sa + hm * L::kAHalfBytes);
#pragma unroll
                            for (int k = 0; k < kMmasPerBlock; ++k)
                                umma_bf16_ss_2sm(d_tmem + (uint32_t)(hm * BLOCK_N), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2),
                                                 kIdesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_2sm(&empty_bar[stage], 3);                         // frees the slot in both CTAs
                        if (kb == k_blocks - 1) umma_commit_2sm(&tmem_full[acc], 3);   // accumulators of both CTAs complete
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= kEpilogueWarp0) {
        // ===================== epilogue (both CTAs, own 128-row halves) =====================
        const int q = (warp - kEpilogueWarp0) & 3;
        const int grp = (warp - kEpilogueWarp0) >> 2;
        const int acc = grp;
        uint32_t acc_phase = 0;
        unsigned char *stg = smem + L::kStagingOffset + (grp * 4 + q) * 4096;
        int local_tile = 0;
        for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs, ++local_tile) {
            if ((local_tile & 1) != grp) continue;
            const int m_tile = pt * 2 + rank;
            mbar_wait_sel(&tmem_full[acc], acc_phase, p.park & 8u);
            tcgen05_fence_after();
#pragma unroll 1
            for (int hm = 0; hm < MT; ++hm) {
                const long long m = (long long)m_tile * kTileM + hm * kBlockM + q * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + hm * BLOCK_N);
                const bool row_ok = m < p.M;
#pragma unroll 1
                for (int g = 0; g < BLOCK_N / 64; ++g) {
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64), r0);
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64 + 32), r1);
                    tmem_ld_wait();
                    if (g == BLOCK_N / 64 - 1 && hm == MT - 1) {   // drained: hand the stage back to the leader's MMA warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
                    }
                    if (lane == 0) bulk_wait_read_all();
                    __syncwarp();
                    const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s + g * 64);
                    const uint4 *rp = (p.residual && row_ok)
                                          ? reinterpret_cast<const uint4 *>(p.residual + m * p.ldr + g * 64)
                                          : nullptr;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v[8];
                        const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = j * 8 + e;
                            const uint32_t raw = c < 32 ? r0[c] : r1[c - 32];
                            v[e] = apply_act_fast(__uint_as_float(raw) + bb[e], p.act);
                        }
                        if (rp) {
                            const uint4 rr = rp[j];
                            v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                            v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                        }
                        uint4 o;
                        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                        *reinterpret_cast<uint4 *>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_out, stg, p.out_col0 + g * 64, m_tile * kTileM + hm * kBlockM + q * 32);
                        bulk_commit_group();
                    }
                }
            }
            acc_phase ^= 1;
        }
        if (lane == 0) bulk_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                              // nobody leaves while the peer may still read / signal here
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm<kTmemCols>(tmem_base);
    }
}


// ---------------------------------------------------------------------------
// Dense 3x3 convolution from ONE halo tile per K chunk (CTA pair, cta_group::2).
// The im2col form above fetches every input pixel nine times (once per tap) from L2 and writes it nine times into
// shared memory; this form loads the (16 + 2d) x (8 + 2d) pixel halo patch of a 16 x 8 output tile once per
// 64-channel chunk (one 4D tiled TMA box, zero fill outside the map = 'same' padding) and addresses all nine taps as
// SHIFTED VIEWS of that patch: with a tile 8 pixels wide the 128 MMA rows are 16 groups of 8 consecutive patch
// pixels, so a K-major SWIZZLE_128B descriptor whose start address is moved by (r d (8 + 2d) + s d) 128-byte rows and
// whose stride between 8-row groups (SBO) is the patch width (8 + 2d) * 128 bytes walks exactly the pixels tap
// (r, s) needs.  The hardware applies the 128-byte swizzle to absolute shared-memory address bits, so any
// 128-byte-aligned start row and any SBO are legal with base_offset = 0 (tools/ubench/desc_shift.cu).
// Patch traffic L2 -> SM per output pixel: 1.41x (d = 1) / 1.88x (d = 2) instead of 9x; the price is spatial tiling:
// 46 x 46 maps are covered by 3 x 6 tiles of 16 x 8 (8.9 % of the MMA rows fall outside the map and are clipped by
// the TMA store).
// Roles: warp 0 patch (A) producer, warp 3 weight (B) producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM
// allocation, warps 4-11 two epilogue groups.  Loop order per work item (MT tiles per CTA): chunk, tap, tile -- one
// weight stage feeds the MMAs of both tiles.
// ---------------------------------------------------------------------------
template <int DIL, int MT, int SB>
struct HaloSmem {
    static constexpr int HW = 8 + 2 * DIL, HH = 16 + 2 * DIL;
    static constexpr int kARows = HW * HH;
    static constexpr int kATx = kARows * 128;                         // bytes one patch load delivers
    static constexpr int kABytes = ((kATx + 1023) / 1024) * 1024;
    static constexpr int SA = 2 * MT;                                 // patches of the current and the next chunk
    static constexpr int kBBytes = 64 * 128;                          // this CTA's half of a [128 x 64] weight chunk
    static constexpr int kBOffset = SA * kABytes;
    static constexpr int kBarOffset = kBOffset + SB * kBBytes;
    static constexpr int kNumBars = 2 * SA + 2 * SB + 4;
    static constexpr int kBiasOffset = kBarOffset + ((kNumBars * 8 + 8 + 15) / 16) * 16;
    static constexpr int kStagingOffset = ((kBiasOffset + 128 * 4 + 1023) / 1024) * 1024;
    static constexpr int kTotal = kStagingOffset + 8 * 4096;
    static_assert(kTotal <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void tma_load_4d_2sm(void *smem_dst, const void *desc, uint64_t *bar, int c0, int c1, int c2,
                                                int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void *desc, const void *smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// K-major SWIZZLE_128B descriptor with an explicit stride between 8-row groups
__device__ __forceinline__ uint64_t make_kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

template <int DIL, int MT, int SB>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_out, const GemmParams p) {
    using L = HaloSmem<DIL, MT, SB>;
    constexpr int BLOCK_N = 128, SA = L::SA;
    constexpr int kAccCols = MT * BLOCK_N;
    constexpr uint32_t kTmemCols = 2 * kAccCols <= 256 ? 256 : 512;
    constexpr uint32_t kIdesc = make_idesc_bf16(256, BLOCK_N);        // M = 256 across the pair
    constexpr uint32_t kSbo = L::HW * 128;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("lwop: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *a_empty = a_full + SA;
    uint64_t *b_full = a_empty + SA;
    uint64_t *b_empty = b_full + SB;
    uint64_t *tmem_full = b_empty + SB;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *bias_s = reinterpret_cast<float *>(smem + L::kBiasOffset);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int chunks = p.k_blocks_per_tap;                            // 64-channel chunks of Cin
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_items = (p.num_tiles + 2 * MT - 1) / (2 * MT);
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2sm<kTmemCols>(tmem_ptr);
    for (int i = threadIdx.x; i < 128; i += kNumThreads) bias_s[i] = p.bias[i];
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    pdl_launch_dependents();

    // tile t -> (image, first row, first column); tiles past the end name image B: every load is zero fill, every
    // store is clipped
    auto tile_origin = [&](int t, int &n, int &y0, int &x0) {
        n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int ty = rem / p.tiles_x;
        y0 = ty * 16;
        x0 = (rem - ty * p.tiles_x) * 8;
        if (t >= p.num_tiles) { n = p.B; y0 = 0; x0 = 0; }
    };

    if (warp == 0) {
        // ===================== patch (A) producer: both CTAs, own tiles =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = pair_id; it < num_items; it += num_pairs) {
                for (int c = 0; c < chunks; ++c) {
#pragma unroll
                    for (int hm = 0; hm < MT; ++hm) {
                        int n, y0, x0;
                        tile_origin((it * 2 + rank) * MT + hm, n, y0, x0);
                        mbar_wait_sel(&a_empty[stage], phase ^ 1, p.park & 1u);
                        if (leader) mbar_arrive_expect_tx(&a_full[stage], 2 * L::kATx);
                        tma_load_4d_2sm(smem + stage * L::kABytes, &map_a, &a_full[stage], c * 64, x0 - DIL, y0 - DIL, n);
                        if (++stage == SA) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== weight (B) producer: both CTAs, own half of the 128 output channels =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = pair_id; it < num_items; it += num_pairs) {
                for (int c = 0; c < chunks; ++c) {
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait_sel(&b_empty[stage], phase ^ 1, p.park & 1u);
                        if (leader) mbar_arrive_expect_tx(&b_full[stage], 2 * L::kBBytes);
                        tma_load_2d_2sm(smem + L::kBOffset + stage * L::kBBytes, &map_b, &b_full[stage],
                                        (tap * chunks + c) * 64, rank * 64);
                        if (++stage == SB) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only, for the pair) =====================
        if (leader) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int it = pair_id; it < num_items; it += num_pairs) {
                mbar_wait_sel(&tmem_empty[acc], acc_phase ^ 1, p.park & 2u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
                for (int c = 0; c < chunks; ++c) {
                    // the MT patches of this chunk sit in stages sa .. sa + MT - 1 (SA is a multiple of MT)
#pragma unroll
                    for (int hm = 0; hm < MT; ++hm) mbar_wait_sel(&a_full[sa + hm], pa, p.park & 2u);
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_u32(smem + sa * L::kABytes);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait_sel(&b_full[sb], pb, p.park & 2u);
                        tcgen05_fence_after();
                        const int r = tap / 3, s = tap - r * 3;
                        const uint32_t shift = (uint32_t)((r * DIL * L::HW + s * DIL) * 128);
                        if (elect_one()) {
                            const uint64_t db = make_kmajor_desc<128>(smem_u32(smem + L::kBOffset + sb * L::kBBytes));
#pragma unroll
                            for (int hm = 0; hm < MT; ++hm) {
                                const uint64_t da = make_kmajor_desc_sbo(a0 + (uint32_t)(hm * L::kABytes) + shift, kSbo);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_ss_2sm(d_tmem + (uint32_t)(hm * BLOCK_N), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2),
                                                     kIdesc, (c | tap | k) != 0 ? 1u : 0u);
                            }
                            umma_commit_2sm(&b_empty[sb], 3);
                            if (tap == 8) {
#pragma unroll
                                for (int hm = 0; hm < MT; ++hm) umma_commit_2sm(&a_empty[sa + hm], 3);
                                if (c == chunks - 1) umma_commit_2sm(&tmem_full[acc], 3);
                            }
                        }
                        __syncwarp();
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                    sa += MT;
                    if (sa == SA) { sa = 0; pa ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= kEpilogueWarp0) {
        // ===================== epilogue (both CTAs): warp q of a group owns tile rows 4q .. 4q + 3 =====================
        const int q = (warp - kEpilogueWarp0) & 3;
        const int grp = (warp - kEpilogueWarp0) >> 2;
        const int acc = grp;
        uint32_t acc_phase = 0;
        unsigned char *stg = smem + L::kStagingOffset + (grp * 4 + q) * 4096;
        int local = 0;
        for (int it = pair_id; it < num_items; it += num_pairs, ++local) {
            if ((local & 1) != grp) continue;
            mbar_wait_sel(&tmem_full[acc], acc_phase, p.park & 8u);
            tcgen05_fence_after();
#pragma unroll 1
            for (int hm = 0; hm < MT; ++hm) {
                int n, y0, x0;
                tile_origin((it * 2 + rank) * MT + hm, n, y0, x0);
                const int y = y0 + q * 4 + (lane >> 3), x = x0 + (lane & 7);
                const bool px_ok = n < p.B && y < p.H && x < p.W;
                const long long m = ((long long)n * p.H + y) * p.W + x;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + hm * BLOCK_N);
#pragma unroll 1
                for (int g = 0; g < BLOCK_N / 64; ++g) {
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64), r0);
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 64 + 32), r1);
                    tmem_ld_wait();
                    if (g == BLOCK_N / 64 - 1 && hm == MT - 1) {   // drained: hand the stage back to the leader's MMA warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
                    }
                    if (lane == 0) bulk_wait_read_all();
                    __syncwarp();
                    const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s + g * 64);
                    const uint4 *rp = (p.residual && px_ok)
                                          ? reinterpret_cast<const uint4 *>(p.residual + m * p.ldr + g * 64)
                                          : nullptr;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v[8];
                        const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = j * 8 + e;
                            const uint32_t raw = c < 32 ? r0[c] : r1[c - 32];
                            v[e] = apply_act_fast(__uint_as_float(raw) + bb[e], p.act);
                        }
                        if (rp) {
                            const uint4 rr = rp[j];
                            v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                            v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                        }
                        uint4 o;
                        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                        *reinterpret_cast<uint4 *>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        // box (64 channels, 8 columns, 4 rows, 1 image); pixels outside the map are clipped
                        tma_store_4d(&map_out, stg, p.out_col0 + g * 64, x0, y0 + q * 4, n);
                        bulk_commit_group();
                    }
                }
            }
            acc_phase ^= 1;
        }
        if (lane == 0) bulk_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm<kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
int g_num_sms = 0;

}  // namespace

// libcuda is resolved at run time through the runtime's driver entry points so
// that the shared library still loads (and exports its symbols) on a machine
// without a GPU driver.
bool gemm_driver_ready(char *err, size_t errlen) {
    if (g_encode_tiled && g_encode_im2col) return true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        return false;
    }
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        snprintf(err, errlen, "cuTensorMapEncodeIm2col not available: %s", cudaGetErrorString(e));
        return false;
    }
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    return true;
}

int gemm_num_sms() { return g_num_sms; }
bool pdl_enabled() {
    static const bool on = !(getenv("LWOP_PDL") && atoi(getenv("LWOP_PDL")) == 0);
    return on;
}

bool encode_tiled_bf16(CUtensorMap *map, int rank, const void *base, const cuuint64_t *dims,
                       const cuuint64_t *strides_bytes, const cuuint32_t *box, int swizzle_bytes, char *err,
                       size_t errlen) {
    if (!gemm_driver_ready(err, errlen)) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), dims,
                                strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled (rank %d) failed: CUresult %d", rank, (int)r);
        return false;
    }
    return true;
}

bool encode_tiled_f32(CUtensorMap *map, int rank, const void *base, const cuuint64_t *dims,
                      const cuuint64_t *strides_bytes, const cuuint32_t *box, char *err, size_t errlen) {
    if (!gemm_driver_ready(err, errlen)) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), dims,
                                strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled (fp32, rank %d) failed: CUresult %d", rank, (int)r);
        return false;
    }
    return true;
}

struct GemmPlan {
    CUtensorMap map_a, map_b, map_out;
    GemmParams p;
    int block_n, kswz, im2col, stages;
    int cluster;
    int mt;
    int two_sm;        // CTA-pair (cta_group::2) kernel
    int halo;          // dense 3x3 from one halo patch per chunk (conv3x3_halo_kernel), dilation = p.dil
    int grid;
    size_t smem;
};

namespace {

template <int BLOCK_N, int KSWZ, bool IM2COL, int STAGES, int CLUSTER, int MT>
cudaError_t launch_tc(const GemmPlan *pl, cudaStream_t s, void *out_override, float *out2_override) {
    auto kfn = gemm_conv_kernel<BLOCK_N, KSWZ, IM2COL, STAGES, CLUSTER, MT>;
    constexpr size_t smem = SmemLayout<BLOCK_N, KSWZ, STAGES, MT>::kTotal;
    // per-device attribute; cheap enough to set on every launch (and outside graph replay)
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    GemmParams p = pl->p;
    if (out_override) p.out = out_override;
    if (out2_override) p.out2 = out2_override;
    if (out_override) p.use_tma_store = 0;         // the store map is bound to the plan's own output
    return launch_chain(kfn, (unsigned)pl->grid, kNumThreads, smem, s, CLUSTER, pl->map_a, pl->map_b, pl->map_out, p);
}

template <int BLOCK_N, int KSWZ, bool IM2COL, int STAGES>
cudaError_t launch_t(const GemmPlan *pl, cudaStream_t s, void *out_override, float *out2_override) {
    if (pl->cluster == 2) return launch_tc<BLOCK_N, KSWZ, IM2COL, STAGES, 2, 1>(pl, s, out_override, out2_override);
    return launch_tc<BLOCK_N, KSWZ, IM2COL, STAGES, 1, 1>(pl, s, out_override, out2_override);
}

}  // namespace

namespace {
template <bool IM2COL, int STAGES, int MT>
cudaError_t launch_2sm(const GemmPlan *pl, cudaStream_t s) {
    auto kfn = gemm_conv2sm_kernel<IM2COL, STAGES, MT>;
    constexpr size_t smem = SmemLayout2Sm<STAGES, MT>::kTotal;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_chain(kfn, (unsigned)pl->grid, kNumThreads, smem, s, 2, pl->map_a, pl->map_b, pl->map_out, pl->p);
}
}  // namespace

namespace {
template <int DIL>
cudaError_t launch_halo(const GemmPlan *pl, cudaStream_t s) {
    constexpr int MT = 2, SB = 8;
    auto kfn = conv3x3_halo_kernel<DIL, MT, SB>;
    constexpr size_t smem = HaloSmem<DIL, MT, SB>::kTotal;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_chain(kfn, (unsigned)pl->grid, kNumThreads, smem, s, 2, pl->map_a, pl->map_b, pl->map_out, pl->p);
}
}  // namespace

cudaError_t gemm_plan_launch(const GemmPlan *pl, cudaStream_t s, void *out_override, float *out2_override) {
    if (pl->halo && !out_override && !out2_override) return pl->p.dil == 2 ? launch_halo<2>(pl, s) : launch_halo<1>(pl, s);
    if (pl->two_sm && !out_override && !out2_override) {
        if (pl->mt == 2) return pl->im2col ? launch_2sm<true, 4, 2>(pl, s) : launch_2sm<false, 4, 2>(pl, s);
        return pl->im2col ? launch_2sm<true, 7, 1>(pl, s) : launch_2sm<false, 7, 1>(pl, s);
    }
    if (pl->mt == 2) {      // 256-row CTA tiles
        if (pl->block_n == 128 && pl->kswz == 128)          // 4 stages of 48 KB
            return pl->im2col ? launch_tc<128, 128, true, 4, 1, 2>(pl, s, out_override, out2_override)
                              : launch_tc<128, 128, false, 4, 1, 2>(pl, s, out_override, out2_override);
        if (pl->block_n == 64 && pl->kswz == 128 && !pl->im2col)
            return launch_tc<64, 128, false, 4, 1, 2>(pl, s, out_override, out2_override);
        if (pl->block_n == 64 && pl->kswz == 64 && !pl->im2col)
            return launch_tc<64, 64, false, 8, 1, 2>(pl, s, out_override, out2_override);
        return cudaErrorInvalidValue;
    }
    if (pl->im2col) {
        if (pl->block_n == 128 && pl->kswz == 128) return launch_t<128, 128, true, 6>(pl, s, out_override, out2_override);
        return cudaErrorInvalidValue;
    }
    if (pl->kswz == 64) {
        if (pl->block_n == 64) return launch_t<64, 64, false, 8>(pl, s, out_override, out2_override);
        return cudaErrorInvalidValue;
    }
    switch (pl->block_n) {
        case 16: return launch_t<16, 128, false, 8>(pl, s, out_override, out2_override);
        case 32: return launch_t<32, 128, false, 8>(pl, s, out_override, out2_override);
        case 64: return launch_t<64, 128, false, 8>(pl, s, out_override, out2_override);
        case 128: return launch_t<128, 128, false, 6>(pl, s, out_override, out2_override);
        case 256: return launch_t<256, 128, false, 4>(pl, s, out_override, out2_override);
    }
    return cudaErrorInvalidValue;
}

static inline bool im2col_early(int ksize) { return ksize == 3; }

GemmPlan *gemm_plan_create(const GemmConvDesc &d, char *err, size_t errlen) {
    if (!gemm_driver_ready(err, errlen)) return nullptr;
    if (d.ksize != 1 && d.ksize != 3) {
        snprintf(err, errlen, "gemm conv: kernel size %d unsupported", d.ksize);
        return nullptr;
    }
    const int kswz = (d.Cin_pad % 64 == 0) ? 128 : 64;
    const int block_k = kswz / 2;
    if (d.Cin_pad % block_k != 0 || d.Cin_pad < 32) {
        snprintf(err, errlen, "gemm conv: Cin_pad=%d must be a multiple of 32", d.Cin_pad);
        return nullptr;
    }
    int block_n;
    if (d.Cout_pad % 256 == 0) block_n = 256;
    else if (d.Cout_pad % 128 == 0) block_n = 128;
    else if (d.Cout_pad == 64 || d.Cout_pad == 32 || d.Cout_pad == 16) block_n = d.Cout_pad;
    else {
        snprintf(err, errlen, "gemm conv: Cout_pad=%d unsupported", d.Cout_pad);
        return nullptr;
    }
    const bool im2col = d.ksize == 3;
    if (im2col && d.Cout_pad % 128 == 0) block_n = 128;
    if (im2col && !(block_n == 128 && kswz == 128)) {
        snprintf(err, errlen, "gemm conv 3x3: only Cout_pad %% 128 == 0 and Cin %% 64 == 0 supported");
        return nullptr;
    }
    if (kswz == 64 && block_n != 64) {
        snprintf(err, errlen, "gemm conv: Cin_pad=32 only supported with Cout_pad=64");
        return nullptr;
    }
    GemmPlan *pl = (GemmPlan *)calloc(1, sizeof(GemmPlan));
    const long long M = (long long)d.B * d.H * d.W;
    const int taps = d.ksize * d.ksize;
    pl->block_n = block_n;
    pl->kswz = kswz;
    pl->im2col = im2col;
    GemmParams &p = pl->p;
    p.bias = d.bias;
    p.residual = (const __nv_bfloat16 *)d.residual;
    p.out = d.out;
    p.out2 = d.out2;
    p.M = M;
    p.H = d.H;
    p.W = d.W;
    p.Cout = d.Cout;
    p.Cout_pad = d.Cout_pad;
    p.k_blocks_per_tap = d.Cin_pad / block_k;
    p.taps = taps;
    p.dil = d.dil;
    p.ldo = d.ldo;
    p.out_col0 = d.out_col0;
    p.ldr = d.ldr;
    p.out_is_f32 = d.out_dtype == LWOP_DT_F32;
    p.act = d.act;
    // 256-row tiles for the N = 128 convolutions with a deep K (3x3: K = 1152); LWOP_MT=1 disables
    const char *mtenv = getenv("LWOP_MT");
    const int mt_mode = mtenv ? atoi(mtenv) : 0;     // 0 = heuristic, 1 = never, 2 = wherever the tile shapes allow
    const bool mt_legal = M >= 4096 && ((block_n == 128 && kswz == 128) || (block_n == 64 && !im2col_early(d.ksize)));
    const bool mt_heur = block_n == 128 && kswz == 128 && taps * d.Cin_pad >= 512;
    pl->mt = (mt_legal && mt_mode != 1 && (mt_heur || mt_mode == 2)) ? 2 : 1;
    const int tile_m = kBlockM * pl->mt;
    p.num_m_tiles = (int)((M + tile_m - 1) / tile_m);
    {
        static const char *pe = getenv("LWOP_GEMM_PARK");
        p.park = pe ? (unsigned)atoi(pe) : 1u;
    }
    p.num_n_tiles = d.Cout_pad / block_n;
    // LWOP_CLUSTER=2: clusters of 2 CTAs share each weight tile through TMA multicast.  Measured on B200
    // (profiles/r1_notes.md): no gain -- at cluster size 2 the multicast does not reduce L2->SM traffic,
    // which is what bounds these kernels -- so the default stays 1.
    const char *cenv = getenv("LWOP_CLUSTER");
    pl->cluster = (cenv && atoi(cenv) == 2 && pl->mt == 1) ? 2 : 1;
    if (p.num_m_tiles < 2) pl->cluster = 1;
    // CTA-pair kernel (cta_group::2) for the N = 128 convolutions with wide bf16 output; LWOP_2SM=0 disables
    const char *e2 = getenv("LWOP_2SM");
    const bool want_2sm = e2 ? atoi(e2) != 0 : true;
    pl->two_sm = (want_2sm && block_n == 128 && kswz == 128 && d.Cout_pad == 128 && d.Cout == 128 && p.num_m_tiles >= 4 &&
                  d.out_dtype == LWOP_DT_BF16 && d.out != nullptr && (d.out_col0 % 8) == 0 && (d.ldo % 8) == 0)
                     ? 1 : 0;
    if (pl->two_sm) pl->cluster = 2;
    const int work = ((p.num_m_tiles + pl->cluster - 1) / pl->cluster) * p.num_n_tiles;
    const int max_clusters = g_num_sms / pl->cluster;
    pl->grid = (work < max_clusters ? work : max_clusters) * pl->cluster;
    // dense 3x3 from halo patches (LWOP_HALO=0 keeps the im2col form)
    const char *eh = getenv("LWOP_HALO");
    pl->halo = (pl->two_sm && im2col && (d.dil == 1 || d.dil == 2) && d.Cin_pad % 64 == 0 && !(eh && atoi(eh) == 0)) ? 1 : 0;
    if (pl->halo) {
        p.B = d.B;
        p.tiles_x = (d.W + 7) / 8;
        p.tiles_y = (d.H + 15) / 16;
        p.num_tiles = d.B * p.tiles_x * p.tiles_y;
        const int items = (p.num_tiles + 3) / 4;                    // MT = 2 tiles per CTA, two CTAs per item
        pl->grid = (items < max_clusters ? items : max_clusters) * 2;
    }

    const CUtensorMapSwizzle swz = kswz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r;
    if (pl->halo) {
        cuuint64_t dims[4] = {(cuuint64_t)d.Cin_pad, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.lda * 2, (cuuint64_t)d.W * d.lda * 2, (cuuint64_t)d.H * d.W * d.lda * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(8 + 2 * d.dil), (cuuint32_t)(16 + 2 * d.dil), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        r = g_encode_tiled(&pl->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(d.in), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (!im2col) {
        cuuint64_t dims[2] = {(cuuint64_t)d.Cin_pad, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)d.lda * 2};
        cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)kBlockM};
        cuuint32_t estr[2] = {1, 1};
        r = g_encode_tiled(&pl->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d.in), dims, strides,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)d.Cin_pad, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.lda * 2, (cuuint64_t)d.W * d.lda * 2,
                                 (cuuint64_t)d.H * d.W * d.lda * 2};
        int lower[2] = {-d.dil, -d.dil};                  // {W, H}: filter window may start `pad` left/above
        int upper[2] = {-d.dil, -d.dil};                  // pad - (k-1)*dil = dil - 2*dil
        cuuint32_t estr[4] = {1, 1, 1, 1};
        r = g_encode_im2col(&pl->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(d.in), dims, strides,
                            lower, upper, (cuuint32_t)block_k, (cuuint32_t)kBlockM, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r == CUDA_SUCCESS && im2col && !pl->halo) {
        // Same workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): drivers up to
        // 13.1 mis-encode im2col maps of tensors smaller than 128 KiB; clear bit 21 of word 1.
        int drv = 0;
        cudaDriverGetVersion(&drv);
        const unsigned long long bytes = (unsigned long long)d.B * d.H * d.W * d.lda * 2ull;
        if (drv <= 13010 && bytes < 131072ull) reinterpret_cast<uint64_t *>(&pl->map_a)[1] &= ~(1ull << 21);
    }
    if (r != CUDA_SUCCESS) {
        snprintf(err, errlen, "tensor map (A, %s) encode failed: CUresult %d", im2col ? "im2col" : "tiled", (int)r);
        free(pl);
        return nullptr;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)taps * d.Cin_pad, (cuuint64_t)d.Cout_pad};
        cuuint64_t strides[1] = {(cuuint64_t)taps * d.Cin_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)(block_n / pl->cluster)};
        cuuint32_t estr[2] = {1, 1};
        r = g_encode_tiled(&pl->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d.w_packed), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(err, errlen, "tensor map (B) encode failed: CUresult %d", (int)r);
            free(pl);
            return nullptr;
        }
    }
    // wide bf16 outputs leave through shared-memory staging + TMA store (full 128-byte lines)
    const bool wide = d.out_dtype == LWOP_DT_BF16 && d.Cout == d.Cout_pad && (d.out_col0 % 8) == 0 &&
                      (d.ldo % 8) == 0 && block_n >= 64 && d.out != nullptr;
    p.use_tma_store = 0;
    if (pl->halo) {
        cuuint64_t dims[4] = {(cuuint64_t)d.ldo, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.ldo * 2, (cuuint64_t)d.W * d.ldo * 2, (cuuint64_t)d.H * d.W * d.ldo * 2};
        cuuint32_t box[4] = {64, 8, 4, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        r = g_encode_tiled(&pl->map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.out, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(err, errlen, "tensor map (out, halo) encode failed: CUresult %d", (int)r);
            free(pl);
            return nullptr;
        }
        p.use_tma_store = 1;
    } else if (wide) {
        cuuint64_t dims[2] = {(cuuint64_t)d.ldo, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)d.ldo * 2};
        cuuint32_t box[2] = {64, 32};
        cuuint32_t estr[2] = {1, 1};
        r = g_encode_tiled(&pl->map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d.out, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(err, errlen, "tensor map (out) encode failed: CUresult %d", (int)r);
            free(pl);
            return nullptr;
        }
        p.use_tma_store = 1;
    } else {
        pl->map_out = pl->map_b;       // unused, but must be a valid descriptor for the prefetch-free path
    }
    return pl;
}

void gemm_plan_destroy(GemmPlan *p) { free(p); }

}  // namespace lwop
