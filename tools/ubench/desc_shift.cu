// Experiment: which shared-memory rows does a K-major SWIZZLE_128B tcgen05 A descriptor read when its start address is
// NOT aligned to the 1024-byte swizzle pattern and its stride between 8-row groups (SBO) is not a multiple of 1024?
// (The dense 3x3 convolutions want to address all nine taps as shifted views of ONE halo tile in shared memory.)
// The A region is filled the way TMA fills it for a 1024-aligned buffer (16-byte chunk index XOR (row & 7), row = the
// 128-byte row index counted from the aligned base); B is the identity, so D[r][n] = A_view[r][n] exposes exactly
// which (row, channel) the tensor core fetched for MMA row r, column n.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lightweight_openpose_b200/csrc -o desc_shift desc_shift.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "lwop_common.cuh"
using namespace lwop;

constexpr int kRows = 320;          // 128-byte rows in the A region

__global__ void __launch_bounds__(128, 1) k(int start_row, int sbo_bytes, int base_off, int mode, float *out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<64>(&tptr);
    __nv_bfloat16 *A = reinterpret_cast<__nv_bfloat16 *>(smem);
    unsigned char *Bp = smem + kRows * 128;                     // 40960: 1024-aligned
    for (int i = threadIdx.x; i < kRows * 64; i += 128) {
        const int p = i / 64, ch = i % 64;
        const float v = mode == 0 ? (float)(p & 255) : (mode == 1 ? (float)ch : (float)(p >> 8));
        const int off = p * 128 + (((ch >> 3) ^ (p & 7)) << 4) + (ch & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(smem + off) = __float2bfloat16_rn(v);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += 128) {
        const int n = i / 64, kk = i % 64;
        const int off = n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(Bp + off) = __float2bfloat16_rn(n == kk ? 1.0f : 0.0f);
    }
    (void)A;
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64);
        uint64_t da = make_kmajor_desc<128>(smem_u32(smem) + (uint32_t)start_row * 128u);
        da &= ~(uint64_t(0x3FFF) << 32);
        da |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
        da |= (uint64_t)(base_off & 7) << 49;
        const uint64_t db = make_kmajor_desc<128>(smem_u32(Bp));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tb, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, kk > 0 ? 1u : 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tb);
}

int main() {
    float *d;
    cudaMalloc(&d, 128 * 64 * 4);
    const int smem = kRows * 128 + 8192 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> rowv(128 * 64), chv(128 * 64);
    struct T { int s, sbo, bo; };
    std::vector<T> tests;
    for (int s : {0, 1, 3, 8, 11, 21})
        for (int sbo : {1024, 1280, 1536, 2048, 2304})
            for (int bo : {0, -1}) tests.push_back({s, sbo, bo});       // -1: base offset = start row & 7
    for (auto t : tests) {
        const int bo = t.bo < 0 ? (t.s & 7) : t.bo;
        if (t.bo < 0 && bo == 0) continue;
        k<<<1, 128, smem>>>(t.s, t.sbo, bo, 0, d);
        cudaMemcpy(rowv.data(), d, rowv.size() * 4, cudaMemcpyDeviceToHost);
        k<<<1, 128, smem>>>(t.s, t.sbo, bo, 1, d);
        cudaError_t e = cudaMemcpy(chv.data(), d, chv.size() * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        int bad_row = 0, bad_ch = 0, mixed = 0;
        for (int r = 0; r < 128; ++r) {
            const int want = t.s + (r / 8) * (t.sbo / 128) + r % 8;
            bool same = true;
            for (int n = 0; n < 64; ++n) {
                if ((int)rowv[r * 64 + n] != (want & 255)) ++bad_row;
                if ((int)chv[r * 64 + n] != n) ++bad_ch;
                if (rowv[r * 64 + n] != rowv[r * 64]) same = false;
            }
            if (!same) ++mixed;
        }
        printf("start_row %2d sbo %4d base_off %d : wrong-row elems %5d  wrong-channel elems %5d  rows mixing sources %3d", t.s, t.sbo,
               bo, bad_row, bad_ch, mixed);
        if (bad_row || bad_ch) {
            printf("   e.g. r=0: row %d ch[0..7]=", (int)rowv[0]);
            for (int n = 0; n < 8; ++n) printf("%d,", (int)chv[n]);
            printf(" r=9: row %d (want %d) ch0 %d ch8 %d", (int)rowv[9 * 64], t.s + (t.sbo / 128) + 1, (int)chv[9 * 64], (int)chv[9 * 64 + 8]);
        }
        printf("\n");
    }
    return 0;
}
