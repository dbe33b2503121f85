// Fused head pair: two 2-layer 1x1-conv heads that share one input, as ONE kernel.
//
//   hidden_a = relu(X * W1a^T + b1a)   heat = hidden_a * W2a^T + b2a      (14 channels)
//   hidden_b = relu(X * W1b^T + b1b)   paf  = hidden_b * W2b^T + b2b      (26 channels)
//
// Reference: the initial-stage heads conv(128->512)+conv(512->14) and conv(128->512)+conv(512->26)
// (src/lightweight_openpose.py:32-35, concatenated at :36) and the refinement-stage heads
// conv(128->128)+conv(128->14|26) (:43-46), all tf.layers.conv2d 1x1 + bias (+ReLU on the first of each pair,
// src/net_factory.py:4-12).  The wide hidden tensors (2 x [M,512] for the initial stage) never leave the SM:
// back-to-back GEMMs on the 5th-generation tensor cores.
//
// CTA tile = 128 pixels (rows of the flattened [M, 128] input).  The stacked first layer (N1 = 2*K2 hidden
// channels) is walked in passes of 128 columns:
//   MMA1(p)   acc1[p&1] (TMEM, 128 columns)  = X[128x128] * W1[p-th 128 rows]^T          (K = 128)
//   epilogue-1 (group p&1, 4 warps): tcgen05.ld -> +bias -> ReLU -> bf16 -> tcgen05.st back into tensor memory as
//             the A operand A2[p&1] (64 columns: lane = row, two hidden channels per 32-bit column), released to the
//             MMA thread through an mbarrier after tcgen05.wait::st
//   MMA2(p)   acc2_head (TMEM, 16 or 32 columns) += A2[p&1][128x128, TMEM] * W2_head[:, p-th K slice]^T
//             (tcgen05.mma with the A operand in tensor memory: the hidden activations never touch shared memory,
//             whose bandwidth the N = 128 MMA1 operand reads already use up)
// so MMA1(p+1) overlaps epilogue-1(p), and the two epilogue groups alternate passes.
//   epilogue-2 (group 0): acc2 -> +bias -> fp32 maps and/or the bf16 concat buffer.
// warp 0: TMA producer for W1 chunks; warp 2: TMEM allocator, then TMA producer for X tiles; warp 3: TMA producer for
// W2 slices; warp 1: MMA1 issuer; warp 12: MMA2 issuer; warps 4-11: epilogue groups.
#include <stdlib.h>

#include "lwop_kernels.cuh"

namespace lwop {

namespace {

constexpr int kHpThreads = 416;            // 4 control warps, 8 epilogue warps, 1 MMA2 issuer
constexpr int kHpMma2Warp = 12;
constexpr int kHpEpiWarp0 = 4;
constexpr int kW1Stages = 6;                  // [128 x 64] bf16 chunks, 16 KB each
constexpr int kW2Stages = 4;                  // [<=32 x 128] bf16 slices as two 4 KB K-chunks
constexpr int kXStages = 2;                   // [128 x 128] bf16 input tiles, 32 KB each

struct HpSmem {
    static constexpr int kXBytes = 2 * 16384;
    static constexpr int kW1Bytes = 16384;
    static constexpr int kW2Bytes = 8192;
    static constexpr int kXOff = 0;
    static constexpr int kW1Off = kXOff + kXStages * kXBytes;
    static constexpr int kW2Off = kW1Off + kW1Stages * kW1Bytes;
    static constexpr int kBarOff = kW2Off + kW2Stages * kW2Bytes;
    static constexpr int kNumBars = 2 * kXStages + 2 * kW1Stages + 2 * kW2Stages + 4 + 4 + 4;
    static constexpr int kBias1Off = ((kBarOff + kNumBars * 8 + 8 + 15) / 16) * 16;
    static constexpr int kBias2Off = kBias1Off + 1024 * 4;
    static constexpr int kTotal = kBias2Off + 64 * 4;
};
static_assert(HpSmem::kTotal <= 232448, "shared memory budget exceeded");

struct HpParams {
    const float *b1a, *b1b;      // [K2] each
    const float *b2a, *b2b;      // [16], [32] (padded)
    long long M;
    int K2;                      // hidden width per head (512 or 128)
    int passes;                  // 2 * K2 / 128
    int num_m_tiles;
    __nv_bfloat16 *cat;          // optional bf16 [M][ldc]: heat at columns cat_col_a.., paf at cat_col_b..
    int ldc, cat_col_a, cat_col_b;
    float *out_a, *out_b;        // optional fp32 [M][14], [M][26]
    unsigned long long *prof;    // optional stall counters (LWOP_HEADS_PROF=1), nullptr otherwise
    unsigned park;               // roles whose barrier waits park the warp (bit 0 TMA, 1 MMA, 3 epilogue)
};

enum { HP_TOTAL = 0, HP_MMA_X, HP_MMA_ACC1_EMPTY, HP_MMA_W1, HP_MMA_A2_FULL, HP_MMA_W2, HP_MMA_ACC2_EMPTY,
       HP_E_ACC1_FULL, HP_E_A2_EMPTY, HP_E1_WORK, HP_E_ACC2_FULL, HP_E2_WORK, HP_N };
#define HP_WAIT(slot, stmt)                                           \
    do {                                                              \
        if (p.prof) {                                                 \
            const long long _t0 = clock64();                          \
            stmt;                                                     \
            hp_acc[slot] += (unsigned long long)(clock64() - _t0);    \
        } else {                                                      \
            stmt;                                                     \
        }                                                             \
    } while (0)

// TMEM columns: acc1 stages [0,256); two acc2 stages of 64 columns ([+0,+16) heat, [+32,+64) paf) at 256 / 320;
// two 64-column A2 buffers (bf16 hidden activations, the TMEM A operand of MMA2) at 384 / 448
constexpr uint32_t kAcc2Col = 256, kAcc2Stride = 64, kAcc2OffB = 32;
constexpr uint32_t kA2TmemCol = 384;

// PAIR: the two CTAs of a cluster take two row tiles and issue every MMA as one cta_group::2 instruction (M = 256): each
// CTA stages only half of every weight tile (64 of the 128 W1 rows, 8 / 16 of the W2 rows), which is what the
// N = 128 first GEMM is short of -- its operand reads alone fill the shared-memory pipe in the single-CTA form.
template <bool PAIR>
__global__ void __launch_bounds__(kHpThreads, 1)
head_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1a,
                 const __grid_constant__ CUtensorMap map_w1b, const __grid_constant__ CUtensorMap map_w2a,
                 const __grid_constant__ CUtensorMap map_w2b, const HpParams p) {
    using L = HpSmem;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("lwop: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t *x_full = reinterpret_cast<uint64_t *>(smem + L::kBarOff);
    uint64_t *x_empty = x_full + kXStages;
    uint64_t *w1_full = x_empty + kXStages;
    uint64_t *w1_empty = w1_full + kW1Stages;
    uint64_t *w2_full = w1_empty + kW1Stages;
    uint64_t *w2_empty = w2_full + kW2Stages;
    uint64_t *acc1_full = w2_empty + kW2Stages;      // [2]
    uint64_t *acc1_empty = acc1_full + 2;            // [2]
    uint64_t *a2_full = acc1_empty + 2;              // [2]
    uint64_t *a2_empty = a2_full + 2;                // [2]
    uint64_t *acc2_full = a2_empty + 2;              // [2]
    uint64_t *acc2_empty = acc2_full + 2;            // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc2_empty + 2);
    float *bias1_s = reinterpret_cast<float *>(smem + L::kBias1Off);
    float *bias2_s = reinterpret_cast<float *>(smem + L::kBias2Off);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
    const int P = p.passes, half = P / 2;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;
    // work units: row tiles (single CTA) or pairs of row tiles (cluster); this CTA's tile of unit u is u * 2 + rank
    const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int num_units = PAIR ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;
    constexpr int kW1Rows = PAIR ? 64 : 128;                 // W1 rows this CTA stages per chunk
    constexpr uint32_t kW1Tx = PAIR ? 2 * (L::kW1Bytes / 2) : L::kW1Bytes;
    unsigned long long hp_acc[HP_N];
#pragma unroll
    for (int i = 0; i < HP_N; ++i) hp_acc[i] = 0;
    const long long hp_start = clock64();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w1a);
        tma_prefetch_desc(&map_w1b);
        tma_prefetch_desc(&map_w2a);
        tma_prefetch_desc(&map_w2b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kXStages; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < kW1Stages; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); }
        for (int i = 0; i < kW2Stages; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1);
            mbar_init(&acc1_empty[i], PAIR ? 8 : 4);
            mbar_init(&a2_full[i], PAIR ? 8 : 4);
            mbar_init(&a2_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc2_full[i], 1);
            mbar_init(&acc2_empty[i], PAIR ? 8 : 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) tmem_alloc_2sm<512>(tmem_ptr);
        else tmem_alloc<512>(tmem_ptr);
    }
    for (int i = threadIdx.x; i < 2 * p.K2; i += kHpThreads) bias1_s[i] = i < p.K2 ? p.b1a[i] : p.b1b[i - p.K2];
    for (int i = threadIdx.x; i < 48; i += kHpThreads) bias2_s[i] = i < 16 ? p.b2a[i] : p.b2b[i - 16];
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();                  // the previous kernel of the chain has completed: its output may be read, our output written
    pdl_launch_dependents();     // the next kernel may start its prologue while this one runs

    // PAIR: commits of the leader's MMA threads arrive on the barrier at the same offset in both CTAs
    auto commit = [&](uint64_t *bar) {
        if constexpr (PAIR) umma_commit_2sm(bar, 3);
        else umma_commit(bar);
    };
    if (warp == 0) {
        // ===================== TMA producer: first-layer weight chunks (one stream across all tiles) =====================
        if (lane == 0) {
            int ws = 0;
            uint32_t wph = 0;
            for (int u = unit0; u < num_units; u += unit_step) {
                for (int pp = 0; pp < P; ++pp) {
                    const CUtensorMap *mw = pp < half ? &map_w1a : &map_w1b;
                    const int row0 = (pp < half ? pp : pp - half) * 128 + rank * kW1Rows;
                    for (int kc = 0; kc < 2; ++kc) {
                        mbar_wait_sel(&w1_empty[ws], wph ^ 1, p.park & 1u);
                        if (leader) mbar_arrive_expect_tx(&w1_full[ws], kW1Tx);
                        if constexpr (PAIR) tma_load_2d_2sm(smem + L::kW1Off + ws * L::kW1Bytes, mw, &w1_full[ws], kc * 64, row0);
                        else tma_load_2d(smem + L::kW1Off + ws * L::kW1Bytes, mw, &w1_full[ws], kc * 64, row0);
                        if (++ws == kW1Stages) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== TMA producer: input tiles (its own warp: an X stage wait must not hold back weights) =====================
        if (lane == 0) {
            int xs = 0;
            uint32_t xph = 0;
            for (int u = unit0; u < num_units; u += unit_step) {
                const int tile = PAIR ? u * 2 + rank : u;       // a dummy tile past the end loads zeros (TMA OOB fill)
                mbar_wait_sel(&x_empty[xs], xph ^ 1, p.park & 1u);
                if (leader) mbar_arrive_expect_tx(&x_full[xs], (PAIR ? 2 : 1) * L::kXBytes);
                if constexpr (PAIR) {
                    tma_load_2d_2sm(smem + L::kXOff + xs * L::kXBytes, &map_x, &x_full[xs], 0, tile * 128);
                    tma_load_2d_2sm(smem + L::kXOff + xs * L::kXBytes + 16384, &map_x, &x_full[xs], 64, tile * 128);
                } else {
                    tma_load_2d(smem + L::kXOff + xs * L::kXBytes, &map_x, &x_full[xs], 0, tile * 128);
                    tma_load_2d(smem + L::kXOff + xs * L::kXBytes + 16384, &map_x, &x_full[xs], 64, tile * 128);
                }
                if (++xs == kXStages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ===================== TMA producer: second-layer weight slices =====================
        if (lane == 0) {
            int ws = 0;
            uint32_t wph = 0;
            for (int u = unit0; u < num_units; u += unit_step) {
                for (int pp = 0; pp < P; ++pp) {
                    const bool is_a = pp < half;
                    const int k0 = (is_a ? pp : pp - half) * 128;
                    const uint32_t bytes = is_a ? 2 * 16 * 128 : 2 * 32 * 128;      // both K chunks, all rows (of both CTAs)
                    const int r0 = PAIR ? rank * (is_a ? 8 : 16) : 0;               // PAIR: this CTA stages half of the rows
                    mbar_wait_sel(&w2_empty[ws], wph ^ 1, p.park & 1u);
                    if (leader) mbar_arrive_expect_tx(&w2_full[ws], bytes);
                    unsigned char *dst = smem + L::kW2Off + ws * L::kW2Bytes;
                    if constexpr (PAIR) {
                        tma_load_2d_2sm(dst, is_a ? &map_w2a : &map_w2b, &w2_full[ws], k0, r0);
                        tma_load_2d_2sm(dst + 4096, is_a ? &map_w2a : &map_w2b, &w2_full[ws], k0 + 64, r0);
                    } else {
                        tma_load_2d(dst, is_a ? &map_w2a : &map_w2b, &w2_full[ws], k0, 0);
                        tma_load_2d(dst + 4096, is_a ? &map_w2a : &map_w2b, &w2_full[ws], k0 + 64, 0);
                    }
                    if (++ws == kW2Stages) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA1 issuer: acc1[p & 1] = X * W1[pass p]^T =====================
        // The whole warp walks the (warp-uniform) schedule and waits on the barriers; one elected lane issues.
        // MMA1 and MMA2 have their own issuing warps: a single warp walking both instruction streams (~200 mostly
        // dependent instructions per pass) was itself the bottleneck -- the tensor pipe needs ~750 clocks per pass
        // (tools/ubench/umma_rate.cu, mixed shapes), the one-warp schedule took ~1200.
        constexpr uint32_t kIdesc1 = make_idesc_bf16(PAIR ? 256 : 128, 128);
        if (leader) {                      // PAIR: only the leader CTA issues; its commits arrive in both CTAs
            int xs = 0, w1s = 0;
            uint32_t xph = 0, w1ph = 0;
            uint32_t use0 = 0, use1 = 0;   // uses of acc1 stage 0 / 1, phase bookkeeping
            for (int u = unit0; u < num_units; u += unit_step) {
                HP_WAIT(HP_MMA_X, mbar_wait_sel(&x_full[xs], xph, p.park & 2u));
                const uint32_t xa = smem_u32(smem + L::kXOff + xs * L::kXBytes);
                for (int pp = 0; pp < P; ++pp) {
                    const int s = pp & 1;
                    HP_WAIT(HP_MMA_ACC1_EMPTY, mbar_wait_sel(&acc1_empty[s], ((s ? use1 : use0) & 1) ^ 1, p.park & 2u));
                    tcgen05_fence_after();
                    const uint32_t d1 = tmem_base + (uint32_t)(s * 128);
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc) {
                        HP_WAIT(HP_MMA_W1, mbar_wait_sel(&w1_full[w1s], w1ph, p.park & 2u));
                        tcgen05_fence_after();
                        if (elect_one()) {
                            const uint64_t da = make_kmajor_desc<128>(xa + kc * 16384);
                            const uint64_t db = make_kmajor_desc<128>(smem_u32(smem + L::kW1Off + w1s * L::kW1Bytes));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if constexpr (PAIR)
                                    umma_bf16_ss_2sm(d1, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc1, (kc | k) != 0 ? 1u : 0u);
                                else
                                    umma_bf16_ss(d1, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc1, (kc | k) != 0 ? 1u : 0u);
                            }
                            commit(&w1_empty[w1s]);
                            if (kc == 1) {
                                commit(&acc1_full[s]);
                                if (pp == P - 1) commit(&x_empty[xs]);
                            }
                        }
                        __syncwarp();
                        if (++w1s == kW1Stages) { w1s = 0; w1ph ^= 1; }
                    }
                    if (s) ++use1; else ++use0;
                }
                if (++xs == kXStages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == kHpMma2Warp) {
        // ===================== MMA2 issuer: acc2_head += A2[p & 1] (tensor memory) * W2_head[:, K slice p]^T =====================
        constexpr uint32_t kIdescA = make_idesc_bf16(PAIR ? 256 : 128, 16);
        constexpr uint32_t kIdescB = make_idesc_bf16(PAIR ? 256 : 128, 32);
        if (leader) {
            int w2s = 0;
            uint32_t w2ph = 0;
            uint32_t useb0 = 0, useb1 = 0;     // uses of A2 buffer 0 / 1
            int local_tile = 0;
            for (int u = unit0; u < num_units; u += unit_step, ++local_tile) {
                const int st = local_tile & 1;                 // acc2 stage: the previous tile's outputs drain meanwhile
                for (int pp = 0; pp < P; ++pp) {
                    const int buf = pp & 1;
                    HP_WAIT(HP_MMA_A2_FULL, mbar_wait_sel(&a2_full[buf], (buf ? useb1 : useb0) & 1, p.park & 2u));
                    if (pp == 0)
                        HP_WAIT(HP_MMA_ACC2_EMPTY, mbar_wait_sel(&acc2_empty[st], (uint32_t)(((local_tile >> 1) & 1) ^ 1), p.park & 2u));
                    HP_WAIT(HP_MMA_W2, mbar_wait_sel(&w2_full[w2s], w2ph, p.park & 2u));
                    tcgen05_fence_after();
                    const bool is_a = pp < half;
                    const int kslice = is_a ? pp : pp - half;
                    const uint32_t d = tmem_base + kAcc2Col + (uint32_t)st * kAcc2Stride + (is_a ? 0u : kAcc2OffB);
                    const uint32_t w2 = smem_u32(smem + L::kW2Off + w2s * L::kW2Bytes);
                    if (elect_one()) {
                        // A operand = the hidden activations the epilogue wrote back to tensor memory (no shared-memory round trip)
                        const uint32_t a2t = tmem_base + kA2TmemCol + (uint32_t)(buf * 64);
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
                            const uint64_t db = make_kmajor_desc<128>(w2 + kc * 4096);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if constexpr (PAIR)
                                    umma_bf16_ts_2sm(d, a2t + (uint32_t)((kc * 4 + k) * 8), db + (uint64_t)(k * 2),
                                                     is_a ? kIdescA : kIdescB, (kslice | kc | k) != 0 ? 1u : 0u);
                                else
                                    umma_bf16_ts(d, a2t + (uint32_t)((kc * 4 + k) * 8), db + (uint64_t)(k * 2),
                                                 is_a ? kIdescA : kIdescB, (kslice | kc | k) != 0 ? 1u : 0u);
                            }
                        }
                        commit(&a2_empty[buf]);
                        commit(&w2_empty[w2s]);
                        if (pp == P - 1) commit(&acc2_full[st]);
                    }
                    __syncwarp();
                    if (buf) ++useb1; else ++useb0;
                    if (++w2s == kW2Stages) { w2s = 0; w2ph ^= 1; }
                }
            }
        }
    } else if (warp >= kHpEpiWarp0 && warp < kHpEpiWarp0 + 8) {
        // ===================== epilogue groups =====================
        const int q = (warp - kHpEpiWarp0) & 3;               // TMEM lane quarter == warp % 4
        const int grp = (warp - kHpEpiWarp0) >> 2;            // owns acc1 stage / A2 buffer `grp`
        const int row = q * 32 + lane;                        // row of the tile == TMEM lane
        uint32_t use = 0;
        // releases go to the MMA thread: the leader CTA's copy of the barrier in the PAIR form
        auto arrive = [&](uint64_t *bar) {
            if constexpr (PAIR) mbar_arrive_leader(bar);
            else mbar_arrive(bar);
        };
        // ---- epilogue-2: the two narrow linear outputs of a tile (group 0; acc2 stage = local tile & 1) ----
        auto epilogue2 = [&](int lt, int tile) {
            const int st = lt & 1;
            HP_WAIT(HP_E_ACC2_FULL, mbar_wait_sel(&acc2_full[st], (uint32_t)((lt >> 1) & 1), p.park & 8u));
            const long long hp_e2 = p.prof ? clock64() : 0;
            tcgen05_fence_after();
            uint32_t ra[16], rb[32];
            const uint32_t t2 = tmem_base + ((uint32_t)(q * 32) << 16) + kAcc2Col + (uint32_t)st * kAcc2Stride;
            tmem_ld_32x32b_x16(t2, ra);
            tmem_ld_32x32b_x32(t2 + kAcc2OffB, rb);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) arrive(&acc2_empty[st]);
            const long long m = (long long)tile * 128 + row;
            if (m < p.M) {
                float va[14], vb[26];
#pragma unroll
                for (int j = 0; j < 14; ++j) va[j] = __uint_as_float(ra[j]) + bias2_s[j];
#pragma unroll
                for (int j = 0; j < 26; ++j) vb[j] = __uint_as_float(rb[j]) + bias2_s[16 + j];
                if (p.out_a) {
                    float2 *o = reinterpret_cast<float2 *>(p.out_a + m * 14);     // 56-byte rows: 8-byte aligned
#pragma unroll
                    for (int j = 0; j < 7; ++j) o[j] = make_float2(va[2 * j], va[2 * j + 1]);
                }
                if (p.out_b) {
                    float2 *o = reinterpret_cast<float2 *>(p.out_b + m * 26);     // 104-byte rows
#pragma unroll
                    for (int j = 0; j < 13; ++j) o[j] = make_float2(vb[2 * j], vb[2 * j + 1]);
                }
                if (p.cat) {
                    __nv_bfloat16 *o = p.cat + m * p.ldc;
                    if (p.cat_col_a == 0 && p.cat_col_b == 14 && (p.ldc & 7) == 0) {
                        // the concat layout of lightweight_openpose.py:36 (heat 0..13, paf 14..39): 80 contiguous bytes per
                        // row = five 128-bit stores instead of forty 16-bit ones
                        float v[40];
#pragma unroll
                        for (int j = 0; j < 14; ++j) v[j] = va[j];
#pragma unroll
                        for (int j = 0; j < 26; ++j) v[14 + j] = vb[j];
                        uint4 *o4 = reinterpret_cast<uint4 *>(o);
#pragma unroll
                        for (int q4 = 0; q4 < 5; ++q4) {
                            uint4 w;
                            w.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); w.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                            w.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); w.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                            o4[q4] = w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 14; ++j) o[p.cat_col_a + j] = __float2bfloat16_rn(va[j]);
#pragma unroll
                        for (int j = 0; j < 26; ++j) o[p.cat_col_b + j] = __float2bfloat16_rn(vb[j]);
                    }
                }
            }
            if (p.prof) hp_acc[HP_E2_WORK] += (unsigned long long)(clock64() - hp_e2);
        };
        int local_tile = 0, prev_tile = -1;
        for (int u = unit0; u < num_units; u += unit_step, ++local_tile) {
            const int tile = PAIR ? u * 2 + rank : u;
            for (int pp = grp; pp < P; pp += 2, ++use) {
                HP_WAIT(HP_E_ACC1_FULL, mbar_wait_sel(&acc1_full[grp], use & 1, p.park & 8u));
                tcgen05_fence_after();
                const long long hp_e0 = p.prof ? clock64() : 0;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * 128);
                const float *b1 = bias1_s + pp * 128;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {                 // 32 hidden columns at a time
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), r);
                    tmem_ld_wait();
                    if (c == 3) {                             // accumulator stage drained
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive(&acc1_empty[grp]);
                    }
                    const float4 *bs4 = reinterpret_cast<const float4 *>(b1 + c * 32);
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b0 = bs4[2 * j], bq = bs4[2 * j + 1];
                        const float2 s0 = fadd2(make_float2(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1])),
                                                make_float2(b0.x, b0.y));
                        const float2 s1 = fadd2(make_float2(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3])),
                                                make_float2(b0.z, b0.w));
                        const float2 s2 = fadd2(make_float2(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5])),
                                                make_float2(bq.x, bq.y));
                        const float2 s3 = fadd2(make_float2(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7])),
                                                make_float2(bq.z, bq.w));
                        pk[j * 4 + 0] = pack_relu_bf16x2(s0.x, s0.y); pk[j * 4 + 1] = pack_relu_bf16x2(s1.x, s1.y);
                        pk[j * 4 + 2] = pack_relu_bf16x2(s2.x, s2.y); pk[j * 4 + 3] = pack_relu_bf16x2(s3.x, s3.y);
                    }
                    if (c == 0) {
                        // the MMA2 that read this A2 buffer two passes ago has retired (it was issued right behind this
                        // pass's MMA1, so the wait sits after the first load + math rather than in front of them)
                        HP_WAIT(HP_E_A2_EMPTY, mbar_wait_sel(&a2_empty[grp], (use & 1) ^ 1, p.park & 8u));
                        tcgen05_fence_after();
                    }
                    // TMEM A operand of MMA2: lane = row, hidden column k of this pass in 32-bit column k / 2
                    tmem_st_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + kA2TmemCol + (uint32_t)(grp * 64 + c * 16), pk);
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();                                  // orders the lanes' stores before lane 0's release
                if (lane == 0) arrive(&a2_full[grp]);
                if (p.prof) hp_acc[HP_E1_WORK] += (unsigned long long)(clock64() - hp_e0);
                // the previous tile's outputs: its last MMA2 is issued two passes into this tile, so group 0 drains
                // acc2 after its first pass here instead of stalling the new tile's accumulator stage
                if (grp == 0 && pp == 0 && prev_tile >= 0) epilogue2(local_tile - 1, prev_tile);
            }
            prev_tile = tile;
        }
        if (grp == 0 && prev_tile >= 0) epilogue2(local_tile - 1, prev_tile);
    }
    if (p.prof && lane == 0 && (warp == 1 || warp == kHpMma2Warp || warp == kHpEpiWarp0)) {
        if (warp == 1) hp_acc[HP_TOTAL] = (unsigned long long)(clock64() - hp_start);
#pragma unroll
        for (int i = 0; i < HP_N; ++i)
            if (hp_acc[i]) atomicAdd(p.prof + i, hp_acc[i]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();      // the peer's tensor cores may still read this CTA's operands / write its TMEM
    if (warp == 2) {
        tcgen05_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm<512>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

struct HeadPairPlan {
    CUtensorMap map_x, map_w1a, map_w1b, map_w2a, map_w2b;
    HpParams p;
    int grid;
    int pair;        // cta_group::2 form: clusters of two CTAs, weight boxes of half the rows
};

bool head_pair_supported(int cin, int hidden) { return cin == 128 && (hidden == 512 || hidden == 128); }

HeadPairPlan *head_pair_plan_create(const HeadPairDesc &d, char *err, size_t errlen) {
    if (!gemm_driver_ready(err, errlen)) return nullptr;
    if (!head_pair_supported(d.Cin, d.hidden)) {
        snprintf(err, errlen, "fused head pair: unsupported Cin=%d hidden=%d", d.Cin, d.hidden);
        return nullptr;
    }
    HeadPairPlan *pl = (HeadPairPlan *)calloc(1, sizeof(HeadPairPlan));
    HpParams &p = pl->p;
    p.b1a = d.b1a; p.b1b = d.b1b; p.b2a = d.b2a; p.b2b = d.b2b;
    p.M = d.M;
    p.K2 = d.hidden;
    p.passes = 2 * d.hidden / 128;
    p.num_m_tiles = (int)((d.M + 127) / 128);
    {
        static const char *pk = getenv("LWOP_HEADS_PARK");
        p.park = pk ? (unsigned)atoi(pk) : 1u;
    }
    p.cat = (__nv_bfloat16 *)d.cat; p.ldc = d.ldc; p.cat_col_a = d.cat_col_a; p.cat_col_b = d.cat_col_b;
    p.out_a = d.out_a; p.out_b = d.out_b;
    int sms = gemm_num_sms();
    if (sms <= 0) sms = 148;
    pl->grid = p.num_m_tiles < sms ? p.num_m_tiles : sms;
    {
        static const char *pe = getenv("LWOP_HEADS_PAIR");
        pl->pair = ((pe ? atoi(pe) != 0 : false) && p.num_m_tiles >= 4) ? 1 : 0;   // opt-in: measured 2-5 % slower than the single-CTA form
        if (pl->pair) {
            const int units = (p.num_m_tiles + 1) / 2;
            pl->grid = 2 * (units < sms / 2 ? units : sms / 2);
        }
    }
    bool ok = true;
    {
        cuuint64_t dims[2] = {(cuuint64_t)d.Cin, (cuuint64_t)d.M};
        cuuint64_t strides[1] = {(cuuint64_t)d.lda * 2};
        cuuint32_t box[2] = {64, 128};
        ok = ok && encode_tiled_bf16(&pl->map_x, 2, d.in, dims, strides, box, 128, err, errlen);
    }
    for (int hsel = 0; hsel < 2 && ok; ++hsel) {
        cuuint64_t dims[2] = {(cuuint64_t)d.Cin, (cuuint64_t)d.hidden};
        cuuint64_t strides[1] = {(cuuint64_t)d.Cin * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(pl->pair ? 64 : 128)};
        ok = encode_tiled_bf16(hsel ? &pl->map_w1b : &pl->map_w1a, 2, hsel ? d.w1b : d.w1a, dims, strides, box, 128,
                               err, errlen);
    }
    for (int hsel = 0; hsel < 2 && ok; ++hsel) {
        cuuint64_t dims[2] = {(cuuint64_t)d.hidden, (cuuint64_t)(hsel ? 32 : 16)};
        cuuint64_t strides[1] = {(cuuint64_t)d.hidden * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)((hsel ? 32 : 16) / (pl->pair ? 2 : 1))};
        ok = encode_tiled_bf16(hsel ? &pl->map_w2b : &pl->map_w2a, 2, hsel ? d.w2b : d.w2a, dims, strides, box, 128,
                               err, errlen);
    }
    if (!ok) {
        free(pl);
        return nullptr;
    }
    return pl;
}

void head_pair_plan_destroy(HeadPairPlan *p) { free(p); }

cudaError_t head_pair_plan_launch(const HeadPairPlan *pl, cudaStream_t s, float *out_a_override, float *out_b_override) {
    auto kfn = pl->pair ? head_pair_kernel<true> : head_pair_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, HpSmem::kTotal);
    if (e != cudaSuccess) return e;
    HpParams p = pl->p;
    if (out_a_override) p.out_a = out_a_override;
    if (out_b_override) p.out_b = out_b_override;
    static const bool prof = getenv("LWOP_HEADS_PROF") != nullptr;
    if (prof) {   // debugging aid: per-role stall cycles averaged per CTA, printed to stderr (synchronises)
        static const char *names[HP_N] = {"total", "mma:x", "mma:acc1_empty", "mma:w1", "mma:a2_full", "mma:w2",
                                          "mma:acc2_empty", "epi0:acc1_full", "epi0:a2_empty", "epi0:e1_work",
                                          "epi0:acc2_full", "epi0:e2_work"};
        unsigned long long *d = nullptr, hbuf[HP_N];
        cudaMalloc(&d, sizeof(hbuf));
        cudaMemsetAsync(d, 0, sizeof(hbuf), s);
        p.prof = d;
        launch_chain(kfn, (unsigned)pl->grid, kHpThreads, HpSmem::kTotal, s, pl->pair ? 2 : 1, pl->map_x, pl->map_w1a, pl->map_w1b,
                     pl->map_w2a, pl->map_w2b, p);
        cudaStreamSynchronize(s);
        cudaMemcpy(hbuf, d, sizeof(hbuf), cudaMemcpyDeviceToHost);
        cudaFree(d);
        fprintf(stderr, "lwop heads prof: hidden=%d M=%lld grid=%d |", p.K2, p.M, pl->grid);
        for (int i = 0; i < HP_N; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)hbuf[i] / pl->grid);
        fprintf(stderr, "\n");
        return cudaGetLastError();
    }
    return launch_chain(kfn, (unsigned)pl->grid, kHpThreads, HpSmem::kTotal, s, pl->pair ? 2 : 1, pl->map_x, pl->map_w1a, pl->map_w1b,
                        pl->map_w2a, pl->map_w2b, p);
}

}  // namespace lwop
