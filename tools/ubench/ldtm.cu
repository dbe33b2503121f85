// Micro-benchmark: tcgen05.ld drain rate of TMEM per SM (bytes/clk) for 4 / 8 warps, one or two loads in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int INFLIGHT>
__global__ void k(int iters, long long *out, uint32_t *sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[INFLIGHT][32];
#pragma unroll
        for (int f = 0; f < INFLIGHT; ++f) ld32(base + (uint32_t)(((i * INFLIGHT + f) & 7) * 32), r[f]);
        ld_wait();
#pragma unroll
        for (int f = 0; f < INFLIGHT; ++f)
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[f][j];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tptr) : "memory");
}

int main() {
    long long *d;
    uint32_t *sink;
    cudaMalloc(&d, 148 * 8);
    cudaMalloc(&sink, 4096);
    const int iters = 2000;
    for (int threads : {128, 256, 512}) {
        for (int inflight : {1, 2}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (inflight == 1) k<1><<<148, threads>>>(iters, d, sink);
                else k<2><<<148, threads>>>(iters, d, sink);
            }
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * inflight * (threads / 32) * 4096.0;
            printf("warps=%d inflight=%d: %lld clks, %.1f B/clk/SM, %.0f clks per x32 round (%s)\n", threads / 32, inflight, h[0],
                   bytes / h[0], (double)h[0] / iters, cudaGetErrorString(e));
        }
    }
    return 0;
}
