// Micro-benchmark: issue-to-completion time of back-to-back tcgen05.mma (bf16, M = 128, K = 16) for several N,
// with the A operand in shared memory (SS) or tensor memory (TS).  One CTA per SM, all SMs busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lightweight_openpose_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "lwop_common.cuh"
using namespace lwop;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tptr);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t da = make_kmajor_desc<128>(smem_u32(smem));
        const uint64_t db = make_kmajor_desc<128>(smem_u32(smem + 16384));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (TS) umma_bf16_ts(tb, tb + 320 + k * 8, db + (uint64_t)(k * 2), idesc, 1u);
                else umma_bf16_ss(tb, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

// alternating groups of 8 MMAs: N = 128 (A from shared memory, accumulator columns 0..127) and N = 32 (A from tensor
// memory, accumulator columns 256..287) -- the head-pair kernel's instruction mix; does switching shapes cost anything?
__global__ void __launch_bounds__(128, 1) kmix(int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tptr);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x == 0) {
        constexpr uint32_t i1 = make_idesc_bf16(128, 128), i2 = make_idesc_bf16(128, 32);
        const uint64_t da = make_kmajor_desc<128>(smem_u32(smem));
        const uint64_t db = make_kmajor_desc<128>(smem_u32(smem + 16384));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_bf16_ss(tb, da + (uint64_t)((k & 3) * 2), db + (uint64_t)((k & 3) * 2), i1, 1u);
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_bf16_ts(tb + 256, tb + 384 + k * 8, db + (uint64_t)((k & 3) * 2), i2, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

template <int N, bool TS>
void run(long long *d) {
    const int iters = 2000;
    cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<N, TS><<<148, 128, 64 * 1024>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%3d %s: %.1f clks per MMA (%s)\n", N, TS ? "TS" : "SS", (double)h[0] / (iters * 4.0), cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 148 * 8);
    run<16, false>(d); run<32, false>(d); run<64, false>(d); run<128, false>(d); run<256, false>(d);
    run<16, true>(d); run<32, true>(d); run<64, true>(d); run<128, true>(d);
    {
        const int iters = 500;
        cudaFuncSetAttribute(kmix, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        for (int rep = 0; rep < 2; ++rep) kmix<<<148, 128, 64 * 1024>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("mix 8 x (N=128 SS) + 8 x (N=32 TS): %.1f clks per group of 16 (back-to-back rates predict %.1f) (%s)\n",
               (double)h[0] / iters, 8 * 64.1 + 8 * 44.6, cudaGetErrorString(e));
    }
    return 0;
}
