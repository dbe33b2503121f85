// Shared device/host helpers for the lwop kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lwop.h"

namespace lwop {

// ---------------------------------------------------------------------------
// TF 'SAME' padding (reference net_factory.py:6,16,26 pass padding='same'):
// out = ceil(in/stride); total = max((out-1)*stride + (k-1)*dil + 1 - in, 0);
// the smaller half goes before.
// ---------------------------------------------------------------------------
struct SamePad {
    int out, before;
};
__host__ __device__ inline SamePad same_pad(int in, int k, int stride, int dil) {
    SamePad p;
    p.out = (in + stride - 1) / stride;
    int total = (p.out - 1) * stride + (k - 1) * dil + 1 - in;
    if (total < 0) total = 0;
    p.before = total / 2;
    return p;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == LWOP_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == LWOP_ACT_ELU) return v > 0.0f ? v : expm1f(v);
    return v;
}

// Epilogue activation of the tensor-core kernels: ELU through ex2.approx with a cubic Taylor branch
// near zero (exp(v)-1 by subtraction loses everything below ~1e-7 there).  Error <= ~3e-7 absolute,
// far below the bf16 rounding of the stored value; the fp32 check path keeps expm1f.
__device__ __forceinline__ float apply_act_fast(float v, int act) {
    if (act == LWOP_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == LWOP_ACT_ELU) {
        const float small = v * (1.0f + v * (0.5f + v * (1.0f / 6.0f)));
        const float big = __expf(v) - 1.0f;
        const float neg = v > -0.0625f ? small : big;
        return v > 0.0f ? v : neg;
    }
    return v;
}

// element load/store as float for the two storage types
__device__ __forceinline__ float ld_as_float(const float *p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_from_float(float *p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// explicit shared-space 64-bit load (keeps the compiler from emitting generic LD for smem tiles)
__device__ __forceinline__ uint2 lds_u64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}

// packed fp32 FMA (Blackwell FFMA2): two independent fused multiply-adds per instruction
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
// packed fp32 add (Blackwell FADD2)
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
// ELU of a pair for the bf16 epilogues, branch-free: max(v, 0) + (2^(min(v, 0) * log2 e) - 1).  exp(0) = 1 exactly, so
// v >= 0 passes through unchanged; for v < 0 the subtraction leaves an absolute error <= 6e-8 (one ulp of 1.0),
// far below the bf16 rounding of the layer's O(1) outputs.  5 instructions per element (packed add / fma),
// against ~11 for the select-based form; the fp32 check path keeps expm1f.
__device__ __forceinline__ float2 elu2_fast(float2 v) {
    const float2 m = make_float2(fminf(v.x, 0.0f), fminf(v.y, 0.0f));
    const float2 t = ffma2(m, make_float2(1.4426950408889634f, 1.4426950408889634f), make_float2(0.0f, 0.0f));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
    const float2 pos = make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f));
    return fadd2(fadd2(pos, make_float2(-1.0f, -1.0f)), e);
}
// ReLU + round-to-nearest-even pack of two fp32 values into bf16x2 (lo in the low half): one instruction
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

// ---------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 (see /opt/skills/guides/blackwell_cuda_programming.md)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the waiting warp until the phase completes (or this many ns
// pass) instead of returning at once.  A polling loop issues ~6 instructions per probe and takes issue slots from the
// warps that share the scheduler with the waiter (measured: ~1/3 of all instructions issued by the fused separable
// kernel are probes); a parked warp issues nothing but wakes up later.  Which one wins depends on the role: the dense
// GEMM kernels are faster parked, the depthwise producers / epilogues of the fused kernels are faster polling.
constexpr uint32_t kMbarSuspendNs = 20000u;
__device__ __forceinline__ bool mbar_try_wait_park(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
    return ok != 0;
}
// Bounded waits: a pipeline bug must trap (-> CUDA error the host reports) instead of hanging the GPU box.
// 2^26 immediate probes / 2^18 parked probes are several seconds, far beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("lwop: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait_park(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_park(bar, parity)) {
        if (++spins > (1u << 18)) {
            printf("lwop: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait_sel(uint64_t *bar, uint32_t parity, bool park) {
    if (park) mbar_wait_park(bar, parity);
    else mbar_wait(bar, parity);
}

__device__ __forceinline__ void tma_prefetch_desc(const void *desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2D tiled load: coordinates (c0 = innermost, c1)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *desc, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 2D tiled load multicast to every CTA of the cluster named in cta_mask: the box lands at the same
// shared-memory offset in each of them and signals the barrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_multicast(void *smem_dst, const void *desc, uint64_t *bar, int c0,
                                                      int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 4D tiled load: coordinates (c, w, h, n)
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const void *desc, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
// 4D im2col load: base pixel (c, w, h, n) inside the bounding box + filter offsets (w, h)
__device__ __forceinline__ void tma_load_im2col_4d(void *smem_dst, const void *desc, uint64_t *bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c), "r"(w),
        "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

// 2D tiled store smem -> global (bulk async group), coordinates (c0 = innermost, c1)
__device__ __forceinline__ void tma_store_2d(const void *desc, const void *smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk groups have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all committed bulk groups have completed
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One lane of a converged warp (warp-uniform control flow around it lets the compiler keep MMA descriptors and
// barrier addresses in uniform registers instead of converting them per instruction).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster drive ONE tcgen05.mma with M = 256 ----
// Shared-window address of the same object in CTA 0 of the pair (the pair's ranks differ in bit 24).
__device__ __forceinline__ uint32_t leader_smem_u32(const void *p) { return smem_u32(p) & 0xFEFFFFFFu; }
// TMA loads issued by either CTA of the pair into ITS OWN shared memory, signalling the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const void *desc, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(void *smem_dst, const void *desc, uint64_t *bar, int c, int w,
                                                       int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_smem_u32(bar)), "r"(c), "r"(w),
        "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}
// arrive on the LEADER CTA's copy of an mbarrier (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_smem_u32(bar)) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *smem_result) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of each CTA: its 128 rows] * B[smem: each CTA holds N/2 rows]; leader thread issues
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the A operand in tensor memory (each CTA's own 128 lanes)
__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in the CTAs named in cta_mask when all prior pair-MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}

// ---- programmatic dependent launch ----
// Kernels of the forward chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next
// kernel's CTAs may become resident and run their prologue (barrier init, TMEM allocation, descriptor prefetch, bias /
// tap loads -- none of which depend on the previous kernel) while the previous kernel drains.  pdl_wait() blocks until
// the previous kernel has completed and its writes are visible; nothing that kernel produced may be read and nothing
// it reads may be written before it.  pdl_launch_dependents() lets the NEXT kernel start its own prologue early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- tcgen05 ----
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows = lanes, K bf16 elements packed two per 32-bit column, low
// half = even k) is read from tensor memory, e.g. activations a previous epilogue wrote back with tcgen05.st
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> tensor memory: 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier when all previously issued MMAs of this thread complete
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// same, arriving on the barrier at this offset in every CTA of the cluster named in cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t *bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i = lane i)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor for a tile whose rows are `kSwizzleBytes`
// wide (one swizzle atom: 8 rows x kSwizzleBytes, atoms stacked along M/N every
// 8*kSwizzleBytes bytes).  Bit layout per the tcgen05 matrix-descriptor format:
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4,
// [46,48) version = 1, [61,64) layout (2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B).
template <int kSwizzleBytes>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    constexpr uint64_t layout = kSwizzleBytes == 128 ? 2 : (kSwizzleBytes == 64 ? 4 : 6);
    constexpr uint64_t sbo = (8 * kSwizzleBytes) >> 4;
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;   // LBO (ignored for swizzled K-major; canonical value 1)
    d |= sbo << 32;
    d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B (K-major), fp32 D.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
    return (1u << 4)               // D format: F32
           | (1u << 7)             // A format: BF16
           | (1u << 10)            // B format: BF16
           | (0u << 15)            // A K-major
           | (0u << 16)            // B K-major
           | ((n >> 3) << 17)      // N >> 3
           | ((m >> 4) << 24);     // M >> 4
}

}  // namespace lwop
