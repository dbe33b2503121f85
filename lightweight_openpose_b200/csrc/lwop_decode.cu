// On-device pose decode, batched over images.  Restates the observable
// behaviour of the reference src/pose_decode.py:
//   K1 peaks_kernel   : find_peaks_v2 (:52-76) + patch refinement of NMS (:107-185)
//   K2 limbs_kernel   : find_connected_joints (:188-301) with the PAF bicubic
//                       up-sampling of decode_pose (:487) evaluated on the fly
//   K3 group_kernel   : joint list (:481-482), group_limbs_of_same_person (:304-396),
//                       keypoint table of plot_pose (:403-429)
//   K4 pack_kernel    : compaction of the per-image tables for the host copy
// All float arithmetic that feeds an integer decision uses explicit _rn
// intrinsics so the compiler cannot contract it into FMAs: the sequence of
// roundings is the one of the numpy/cv2 code (float32 two-pass bicubic,
// float64 limb scores).
#include <stdlib.h>

#include "lwop_kernels.cuh"

namespace lwop {

namespace {

__constant__ int c_limb_src[13] = {0, 1, 3, 4, 6, 7, 9, 10, 12, 13, 13, 13, 13};   // pose_decode.py:17-18
__constant__ int c_limb_dst[13] = {1, 2, 4, 5, 7, 8, 10, 11, 13, 0, 3, 6, 9};

constexpr int kRadius = 5;          // dis_thres, pose_decode.py:53
constexpr int kPatchHalf = 2;       // win_size,  pose_decode.py:141
constexpr int kMaxSamples = 64;     // largest num_intermed_pts (pose_decode.py:188; the reference default is 10)
constexpr int kMaxWorkPersons = LWOP_NUM_LIMBS * LWOP_MAX_PEAKS;   // one new person per connection at most

// cv2 interpolateCubic, A = -0.75, float32, no contraction
__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
    const float A = -0.75f;
    const float x1 = __fadd_rn(x, 1.0f);
    // ((A*(x+1) - 5A)*(x+1) + 8A)*(x+1) - 4A
    float t = __fsub_rn(__fmul_rn(A, x1), __fmul_rn(5.0f, A));
    t = __fadd_rn(__fmul_rn(t, x1), __fmul_rn(8.0f, A));
    c[0] = __fsub_rn(__fmul_rn(t, x1), __fmul_rn(4.0f, A));
    // ((A+2)*x - (A+3))*x*x + 1
    t = __fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), x), __fadd_rn(A, 3.0f));
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(t, x), x), 1.0f);
    const float xm = __fsub_rn(1.0f, x);
    t = __fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), xm), __fadd_rn(A, 3.0f));
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(t, xm), xm), 1.0f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c[0]), c[1]), c[2]);
}

// source position of destination index d: fx = (float)((d + 0.5) * scale - 0.5)
__device__ __forceinline__ void src_pos(int d, double scale, int &s, float &frac) {
    const float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    const float fl = floorf(f);
    s = (int)fl;
    frac = __fsub_rn(f, fl);
}

// horizontal pass of OpenCV's resize engine (HResizeCubic): products summed left to right
__device__ __forceinline__ float dot4(const float (&v)[4], const float (&c)[4]) {
    float r = __fmul_rn(v[0], c[0]);
    r = __fadd_rn(r, __fmul_rn(v[1], c[1]));
    r = __fadd_rn(r, __fmul_rn(v[2], c[2]));
    r = __fadd_rn(r, __fmul_rn(v[3], c[3]));
    return r;
}
// vertical pass (VResizeCubic): products summed right to left -- with this pair of orders the result equals
// cv2.resize(INTER_CUBIC, float32) bit for bit when OpenCV runs its own code (oracle/decode_oracle.py)
__device__ __forceinline__ float dot4_rev(const float (&v)[4], const float (&c)[4]) {
    float r = __fmul_rn(v[3], c[3]);
    r = __fadd_rn(r, __fmul_rn(v[2], c[2]));
    r = __fadd_rn(r, __fmul_rn(v[1], c[1]));
    r = __fadd_rn(r, __fmul_rn(v[0], c[0]));
    return r;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cv2 saturate_cast<int>(double) = round half to even
__device__ __forceinline__ int cv_round(double v) { return (int)rint(v); }

// ---------------------------------------------------------------------------
// K1: per (image, joint channel): threshold, greedy radius-5 NMS, refinement.
// ---------------------------------------------------------------------------
constexpr int kRefineMax = 48;      // separable refinement handles up-sampled patches up to 48 x 48 (5 x 8 = 40 here)

// NMS(bool_gaussian_filt=True) (:163-164): scipy.ndimage.gaussian_filter(patch, sigma=3).  truncate 4.0 -> radius 12.
// Weights at offsets -12..0 as scipy computes them in float64 (exp(-0.5 / 9 * x^2) / sum; symmetric); generated with
// numpy, checked by tests/test_oracle_decode.py against oracle.decode_oracle.gaussian_weights().
constexpr int kGaussRadius = 12;
__constant__ double c_gauss3[kGaussRadius + 1] = {
    0x1.763a210dfb306p-15, 0x1.4fbe39149e277p-13, 0x1.0d8a5ad43c165p-11, 0x1.8345966f69518p-10, 0x1.f1e9915139406p-9,
    0x1.1e6bccad344bap-7,  0x1.26defcaeb0202p-6,  0x1.0fa58939b528fp-5,  0x1.bfde9c12bec92p-5,  0x1.4a614d1afd337p-4,
    0x1.b42a57d56c0bep-4,  0x1.01a25f86eb137p-3,  0x1.105a329f98197p-3};
// scipy 'reflect' border (d c b a | a b c d | d c b a), any distance
__device__ __forceinline__ int reflect_index(int i, int n) {
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}
// one output of scipy's NI_Correlate1D, symmetric-kernel form: float64 accumulation, outermost pair first
template <typename F>
__device__ __forceinline__ float gauss_line(F &&at, int l, int n) {
    double tmp = __dmul_rn((double)at(l), c_gauss3[kGaussRadius]);
#pragma unroll 1
    for (int j = -kGaussRadius; j < 0; ++j) {
        const double pair = __dadd_rn((double)at(reflect_index(l + j, n)), (double)at(reflect_index(l - j, n)));
        tmp = __dadd_rn(tmp, __dmul_rn(pair, c_gauss3[j + kGaussRadius]));
    }
    return (float)tmp;
}

// Greedy radius-5 suppression of find_peaks_v2 (:62-70) as a parallel fixed point.  The reference takes the best
// remaining candidate, drops every remaining candidate within Euclidean distance 5 of it and repeats; with the
// total order "larger score first, exact ties towards the larger row-major index"  that is the same as:
//     a candidate is KEPT  iff no KEPT candidate of higher order lies within the radius.
// Every candidate therefore only needs the final state of the higher-order candidates in its 81-pixel disk:
//   any of them kept                -> suppressed
//   all of them suppressed (or none) -> kept
//   otherwise                        -> undecided, look again in the next sweep.
// Decisions are final when taken (states only move undecided -> kept / suppressed), so sweeps may read states that
// other threads update concurrently; the highest undecided candidate is always decidable, so the loop terminates.
// Local maxima of the disk are kept in the first sweep and drop their neighbourhoods at once; a candidate that is
// still open remembers the best open higher-order neighbour it saw and only scans its disk again once that one has
// been decided -- so a sweep costs one state read per open candidate plus the scans of the few whose turn has come,
// instead of one block-wide arg-max reduction per kept peak.
constexpr int kStateNone = 0, kStateUndecided = 1, kStateKept = 2, kStateDropped = 3;
constexpr int kKeptListMax = 1024;
constexpr int kDiskOffsets = 80;        // pixels with 0 < dx^2 + dy^2 <= 25 (keeps dis > dis_thres, :68)
// the disk's offsets (dy << 8 | dx & 0xff) in order of increasing distance (ties row-major), built at compile time
struct DiskTable {
    short v[kDiskOffsets];
};
constexpr DiskTable make_disk_table() {
    DiskTable t{};
    int k = 0;
    for (int d2 = 1; d2 <= kRadius * kRadius; ++d2)
        for (int dy = -kRadius; dy <= kRadius; ++dy)
            for (int dx = -kRadius; dx <= kRadius; ++dx)
                if (dx * dx + dy * dy == d2) t.v[k++] = (short)((dy * 256) | (dx & 0xff));
    return t;
}
__constant__ DiskTable c_disk = make_disk_table();

__global__ void __launch_bounds__(256)
peaks_kernel(const float *__restrict__ heat, int h, int w, int C, float thre1, double factor, int mode,
             float *__restrict__ peaks, int *__restrict__ peak_counts) {
    extern __shared__ unsigned char smem_raw[];
    const int npix = h * w;
    float *val = reinterpret_cast<float *>(smem_raw);                 // the channel's map
    short *blk = reinterpret_cast<short *>(val + npix);               // per candidate: the neighbour it waits for
    const int state_off = npix * 4 + ((npix + 1) & ~1) * 2;
    volatile unsigned char *state = smem_raw + state_off;             // per pixel: kState*
    __shared__ int s_n;
    __shared__ short s_disk[kDiskOffsets];                            // disk offsets (dy << 8 | dx & 0xff), nearest first
    __shared__ int s_list[kKeptListMax];                              // kept pixels, unordered
    __shared__ int s_kept[LWOP_MAX_PEAKS];                            // kept pixels, best first
    // separable refinement: tap tables (shared by all peaks: they depend on the destination index only) and
    // the per-warp horizontal pass
    __shared__ float s_h[8][5][kRefineMax];
    __shared__ float s_ct[kRefineMax][4];
    __shared__ signed char s_st[kRefineMax];

    const int ch = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const float *src = heat + (long long)b * npix * C + ch;
    const double scale = 1.0 / factor;              // cv2: scale_x = 1. / inv_scale_x
    if (tid == 0) s_n = 0;
    if (tid >= 128 && tid < 128 + kDiskOffsets) s_disk[tid - 128] = c_disk.v[tid - 128];
    if (tid < kRefineMax) {
        int sx;
        float fx, c[4];
        src_pos(tid, scale, sx, fx);
        cubic_coeffs(fx, c);
        s_st[tid] = (signed char)max(-128, min(127, sx));
#pragma unroll
        for (int k = 0; k < 4; ++k) s_ct[tid][k] = c[k];
    }
    pdl_wait();                 // the maps come from the previous kernel of the chain (forward heads)
    pdl_launch_dependents();
    // channel plane -> shared memory: the strided loads are issued in independent batches of 8 per thread so
    // their latencies overlap; candidates (np.where(img > thre1), :54) are marked on the way
    for (int p0 = 0; p0 < npix; p0 += 8 * (int)blockDim.x) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int p = p0 + k * (int)blockDim.x + tid;
            v[k] = p < npix ? __ldg(src + (long long)p * C) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int p = p0 + k * (int)blockDim.x + tid;
            if (p < npix) {
                val[p] = v[k];
                blk[p] = -1;
                state[p] = v[k] > thre1 ? kStateUndecided : kStateNone;
            }
        }
    }
    __syncthreads();

    int n = 0;
    if (mode & LWOP_PEAKS_CROSS) {
        // find_peaks (:37-49): value > thre1 and equal to the maximum over the 3x3 cross footprint (scipy
        // maximum_filter, 'reflect' border = the border pixel itself), in np.nonzero (row-major) order.
        __shared__ int s_cnt[256];
        const int chunk = (npix + (int)blockDim.x - 1) / (int)blockDim.x;
        const int p0 = tid * chunk, p1 = min(npix, p0 + chunk);
        auto is_peak = [&](int p) {
            const float v = val[p];
            if (!(v > thre1)) return false;
            const int y = p / w, x = p - y * w;
            float m = v;
            if (y > 0) m = fmaxf(m, val[p - w]);
            if (y < h - 1) m = fmaxf(m, val[p + w]);
            if (x > 0) m = fmaxf(m, val[p - 1]);
            if (x < w - 1) m = fmaxf(m, val[p + 1]);
            return m == v;
        };
        int c = 0;
        for (int p = p0; p < p1; ++p) c += is_peak(p) ? 1 : 0;
        s_cnt[tid] = c;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int k = 0; k < (int)blockDim.x; ++k) { const int v = s_cnt[k]; s_cnt[k] = run; run += v; }
            s_n = run;
        }
        __syncthreads();
        int o = s_cnt[tid];
        for (int p = p0; p < p1; ++p)
            if (is_peak(p)) {
                if (o < LWOP_MAX_PEAKS) s_kept[o] = p;
                ++o;
            }
        n = s_n;
        __syncthreads();
    } else {
        // One warp per scan: the lanes look at the disk offsets (nearest first, 32 per round), so a scan is a few
        // shared-memory round trips instead of 81 dependent ones.  A candidate with an open higher-order neighbour in
        // the first round(s) stops there (that neighbour becomes the one it waits for); only a candidate that is about
        // to be kept looks at all 81.  A kept higher-order neighbour that a shortened scan did not reach drops the
        // candidate itself (the scatter below), so stopping early never turns a candidate into a peak wrongly.
        while (true) {
            int pending = 0;
            for (int base = warp * 32; base < npix; base += (int)blockDim.x) {
                const int p = base + lane;
                bool need = false;
                if (p < npix && state[p] == kStateUndecided) {
                    const int bq = blk[p];
                    if (bq >= 0 && state[bq] == kStateUndecided) pending = 1;    // still behind the neighbour it waits for
                    else need = true;
                }
                unsigned todo = __ballot_sync(0xffffffffu, need);
                while (todo) {
                    const int pp = base + __ffs(todo) - 1;
                    todo &= todo - 1;
                    if (state[pp] != kStateUndecided) continue;                  // dropped by a scatter in the meantime
                    const float v = val[pp];
                    const int y = pp / w, x = pp - y * w;
                    bool any_kept = false, found = false;
                    int waitfor = -1;
#pragma unroll 1
                    for (int o0 = 0; o0 < kDiskOffsets && !any_kept && !found; o0 += 32) {
                        const int o = o0 + lane;
                        bool kept_hi = false, open_hi = false;
                        int q = -1;
                        if (o < kDiskOffsets) {
                            const int off = s_disk[o];
                            const int yy = y + (off >> 8), xx = x + (int)(signed char)(off & 0xff);
                            if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
                                q = yy * w + xx;
                                const int sq = state[q];
                                if (sq == kStateUndecided || sq == kStateKept) {
                                    const float vq = val[q];
                                    if (vq > v || (vq == v && q > pp)) {         // higher order than pp
                                        kept_hi = sq == kStateKept;
                                        open_hi = !kept_hi;
                                    }
                                }
                            }
                        }
                        any_kept = __any_sync(0xffffffffu, kept_hi);
                        const unsigned om = __ballot_sync(0xffffffffu, open_hi);
                        if (om) {
                            found = true;
                            waitfor = __shfl_sync(0xffffffffu, q, __ffs(om) - 1);
                        }
                    }
                    if (any_kept) {
                        if (lane == 0) state[pp] = kStateDropped;
                    } else if (found) {
                        if (lane == 0) blk[pp] = (short)waitfor;
                        pending = 1;
                    } else {
                        // no higher-order candidate left in the disk: kept; every open candidate in the disk is of
                        // lower order and dropped for good
                        if (lane == 0) state[pp] = kStateKept;
                        for (int o = lane; o < kDiskOffsets; o += 32) {
                            const int off = s_disk[o];
                            const int yy = y + (off >> 8), xx = x + (int)(signed char)(off & 0xff);
                            if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                            const int q = yy * w + xx;
                            if (state[q] != kStateUndecided) continue;
                            const float vq = val[q];
                            if (vq < v || (vq == v && q < pp)) state[q] = kStateDropped;
                        }
                    }
                    __syncwarp();
                }
            }
            if (!__syncthreads_or(pending)) break;
        }
        for (int p = tid; p < npix; p += blockDim.x)
            if (state[p] == kStateKept) {
                const int i = atomicAdd(&s_n, 1);
                if (i < kKeptListMax) s_list[i] = p;
            }
        __syncthreads();
        n = s_n;
        // kept order of the greedy loop = descending (score, index): rank by counting
        const int nl = min(n, kKeptListMax);
        for (int i = tid; i < nl; i += blockDim.x) {
            const int p = s_list[i];
            const float v = val[p];
            int r = 0;
            for (int j = 0; j < nl; ++j) {
                const int q = s_list[j];
                const float vq = val[q];
                r += (vq > v || (vq == v && q > p)) ? 1 : 0;
            }
            if (r < LWOP_MAX_PEAKS) s_kept[r] = p;
        }
        __syncthreads();
    }
    // more peaks than the table holds: the count is stored as found, consumers clamp it and report the overflow
    if (tid == 0) peak_counts[b * 16 + ch] = n;
    if (n > LWOP_MAX_PEAKS) n = LWOP_MAX_PEAKS;

    // refinement (:148-183): bicubic up-sampling of the clipped 5x5 patch by `factor` (float32 two-pass,
    // horizontal then vertical, cv2 arithmetic), argmax = first maximum in row-major order; warp per peak.
    if (mode & LWOP_PEAKS_NO_REFINE) {
        // bool_refine_center=False (:176-182): low-res peak snapped to the up-scaled grid, score = map value
        for (int i = tid; i < n; i += blockDim.x) {
            const int p = s_kept[i];
            const int py = p / w, px = p - py * w;
            float *o = peaks + (((long long)b * LWOP_NUM_JOINTS + ch) * LWOP_MAX_PEAKS + i) * 4;
            o[0] = (float)floor(__dsub_rn(__dmul_rn(__dadd_rn((double)px, 0.5), factor), 0.5));
            o[1] = (float)floor(__dsub_rn(__dmul_rn(__dadd_rn((double)py, 0.5), factor), 0.5));
            o[2] = val[p];
            o[3] = 0.0f;
        }
        return;
    }
    if (mode & LWOP_PEAKS_GAUSS) {
        // bool_gaussian_filt=True: the up-sampled patch is smoothed (axis 0, then axis 1, each pass rounded to float32)
        // before the argmax.  The whole block works on one peak at a time; the two patch buffers sit behind the
        // channel map in dynamic shared memory (the host sizes it for this mode and checks the patch fits).
        float *gu = reinterpret_cast<float *>(smem_raw + ((state_off + npix + 15) & ~15));   // behind map, waits and states
        float *gt = gu + kRefineMax * kRefineMax;
        __shared__ float s_rv[8];
        __shared__ int s_ri[8];
        for (int i = 0; i < n; ++i) {
            const int p = s_kept[i];
            const int py = p / w, px = p - py * w;
            const int x0 = max(0, px - kPatchHalf), y0 = max(0, py - kPatchHalf);
            const int x1 = min(w - 1, px + kPatchHalf), y1 = min(h - 1, py + kPatchHalf);
            const int pw = x1 - x0 + 1, ph = y1 - y0 + 1;
            const int dw = cv_round((double)pw * factor), dh = cv_round((double)ph * factor);
            for (int e = tid; e < ph * dw; e += blockDim.x) {            // horizontal bicubic pass
                const int r = e / dw, ox = e - r * dw;
                const int sx = s_st[ox];
                float t[4], c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    t[k] = val[(y0 + r) * w + x0 + clampi(sx - 1 + k, 0, pw - 1)];
                    c[k] = s_ct[ox][k];
                }
                s_h[0][r][ox] = dot4(t, c);
            }
            __syncthreads();
            for (int d = tid; d < dw * dh; d += blockDim.x) {            // vertical bicubic pass -> gu
                const int oy = d / dw, ox = d - oy * dw;
                const int sy = s_st[oy];
                float t[4], c[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    t[r] = s_h[0][clampi(sy - 1 + r, 0, ph - 1)][ox];
                    c[r] = s_ct[oy][r];
                }
                gu[d] = dot4_rev(t, c);
            }
            __syncthreads();
            for (int d = tid; d < dw * dh; d += blockDim.x) {            // gaussian along axis 0 (rows) -> gt
                const int oy = d / dw, ox = d - oy * dw;
                gt[d] = gauss_line([&](int y) { return gu[y * dw + ox]; }, oy, dh);
            }
            __syncthreads();
            float best_v = -INFINITY;
            int best_i = 0x7fffffff;
            for (int d = tid; d < dw * dh; d += blockDim.x) {            // gaussian along axis 1 (columns) + argmax
                const int oy = d / dw, ox = d - oy * dw;
                const float v = gauss_line([&](int x) { return gt[oy * dw + x]; }, ox, dw);
                if (v > best_v) { best_v = v; best_i = d; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
            }
            if (lane == 0) { s_rv[warp] = best_v; s_ri[warp] = best_i; }
            __syncthreads();
            if (tid == 0) {
                for (int k = 1; k < nwarps; ++k)
                    if (s_rv[k] > best_v || (s_rv[k] == best_v && s_ri[k] < best_i)) { best_v = s_rv[k]; best_i = s_ri[k]; }
                const int my = best_i / dw, mx = best_i - my * dw;
                const double cxp = __dsub_rn(__dmul_rn(__dadd_rn((double)(px - x0), 0.5), factor), 0.5);
                const double cyp = __dsub_rn(__dmul_rn(__dadd_rn((double)(py - y0), 0.5), factor), 0.5);
                const double ax = __dsub_rn(__dmul_rn(__dadd_rn((double)px, 0.5), factor), 0.5);
                const double ay = __dsub_rn(__dmul_rn(__dadd_rn((double)py, 0.5), factor), 0.5);
                float *o = peaks + (((long long)b * LWOP_NUM_JOINTS + ch) * LWOP_MAX_PEAKS + i) * 4;
                o[0] = (float)floor(__dadd_rn(ax, __dsub_rn((double)mx, cxp)));
                o[1] = (float)floor(__dadd_rn(ay, __dsub_rn((double)my, cyp)));
                o[2] = best_v;
                o[3] = 0.0f;
            }
            __syncthreads();
        }
        return;
    }
    for (int i = warp; i < n; i += nwarps) {
        const int p = s_kept[i];
        const int py = p / w, px = p - py * w;
        const int x0 = max(0, px - kPatchHalf), y0 = max(0, py - kPatchHalf);
        const int x1 = min(w - 1, px + kPatchHalf), y1 = min(h - 1, py + kPatchHalf);
        const int pw = x1 - x0 + 1, ph = y1 - y0 + 1;
        const int dw = cv_round((double)pw * factor), dh = cv_round((double)ph * factor);
        float best_v = -INFINITY;
        int best_i = 0x7fffffff;
        if (dw <= kRefineMax && dh <= kRefineMax) {
            // horizontal pass: every patch row at every destination column
            for (int e = lane; e < ph * dw; e += 32) {
                const int r = e / dw, ox = e - r * dw;
                const int sx = s_st[ox];
                float t[4], c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    t[k] = val[(y0 + r) * w + x0 + clampi(sx - 1 + k, 0, pw - 1)];
                    c[k] = s_ct[ox][k];
                }
                s_h[warp][r][ox] = dot4(t, c);
            }
            __syncwarp();
            // vertical pass + running argmax
            for (int d = lane; d < dw * dh; d += 32) {
                const int oy = d / dw, ox = d - oy * dw;
                const int sy = s_st[oy];
                float t[4], c[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    t[r] = s_h[warp][clampi(sy - 1 + r, 0, ph - 1)][ox];
                    c[r] = s_ct[oy][r];
                }
                const float v = dot4_rev(t, c);
                if (v > best_v) { best_v = v; best_i = d; }
            }
            __syncwarp();
        } else {
            for (int d = lane; d < dw * dh; d += 32) {      // large factors: evaluate each destination pixel directly
                const int oy = d / dw, ox = d - oy * dw;
                int sx, sy;
                float fx, fy, cx[4], cy[4];
                src_pos(ox, scale, sx, fx);
                src_pos(oy, scale, sy, fy);
                cubic_coeffs(fx, cx);
                cubic_coeffs(fy, cy);
                float rows[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int yy = y0 + clampi(sy - 1 + r, 0, ph - 1);
                    float t[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) t[k] = val[yy * w + x0 + clampi(sx - 1 + k, 0, pw - 1)];
                    rows[r] = dot4(t, cx);
                }
                const float v = dot4_rev(rows, cy);
                if (v > best_v) { best_v = v; best_i = d; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
        if (lane == 0) {
            const int my = best_i / dw, mx = best_i - my * dw;
            // floor( resized(peak) + (argmax - resized(peak - patch_origin)) ), float64 (:171-182, :104)
            const double cxp = __dsub_rn(__dmul_rn(__dadd_rn((double)(px - x0), 0.5), factor), 0.5);
            const double cyp = __dsub_rn(__dmul_rn(__dadd_rn((double)(py - y0), 0.5), factor), 0.5);
            const double ax = __dsub_rn(__dmul_rn(__dadd_rn((double)px, 0.5), factor), 0.5);
            const double ay = __dsub_rn(__dmul_rn(__dadd_rn((double)py, 0.5), factor), 0.5);
            const double rx = floor(__dadd_rn(ax, __dsub_rn((double)mx, cxp)));
            const double ry = floor(__dadd_rn(ay, __dsub_rn((double)my, cyp)));
            float *o = peaks + (((long long)b * LWOP_NUM_JOINTS + ch) * LWOP_MAX_PEAKS + i) * 4;
            o[0] = (float)rx;
            o[1] = (float)ry;
            o[2] = best_v;
            o[3] = 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------
// bicubic sample of one PAF channel at full-resolution pixel (X, Y): what
// cv2.resize(pafs, (W, H), INTER_CUBIC)[Y, X, c] holds (pose_decode.py:487).
// ---------------------------------------------------------------------------
struct AxisTaps {
    int idx[4];
    float c[4];
};
__device__ __forceinline__ AxisTaps axis_taps(int d, double scale, int n) {
    AxisTaps t;
    int s;
    float f;
    src_pos(d, scale, s, f);
    cubic_coeffs(f, t.c);
#pragma unroll
    for (int k = 0; k < 4; ++k) t.idx[k] = clampi(s - 1 + k, 0, n - 1);
    return t;
}

// ---------------------------------------------------------------------------
// K2: per (image, limb type): score every (src, dst) joint pair, keep per src the best dst.
// Work is spread over (pair, sample) items -- one thread evaluates ONE of the 10 line-integral samples of one pair
// (the 16-tap bicubic of both PAF channels) -- so the long dependent chains of global loads run in parallel; the
// per-pair reduction and the per-source arg-max then replay the reference's sequential order exactly.
// ---------------------------------------------------------------------------
constexpr int kPairChunk = 192;     // (src, dst) pairs scored per pass
constexpr int kLimbThreads = 256;

// numpy's pairwise sum of n <= 128 contiguous float64 values (np.add.reduce, what ndarray.mean() runs, :271):
// a plain loop below 8 elements, otherwise 8 running sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus the tail
__device__ __forceinline__ double numpy_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

__global__ void __launch_bounds__(kLimbThreads)
limbs_kernel(const float *__restrict__ paf, int h, int w, int H, int W, double thre2, int direct, int nsamp,
             double max_dist_thresh, const float *__restrict__ peaks, const int *__restrict__ peak_counts,
             double *__restrict__ limbs, int *__restrict__ counts) {
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int ja = c_limb_src[t], jb = c_limb_dst[t];
    extern __shared__ double s_val[];                             // [kPairChunk][nsamp]
    __shared__ double s_mean[kPairChunk];
    __shared__ unsigned char s_ok[kPairChunk];
    __shared__ double s_best[LWOP_MAX_PEAKS];
    __shared__ int s_bestj[LWOP_MAX_PEAKS];
    pdl_wait();                                                   // peak table of the kernel before
    pdl_launch_dependents();
    const int *pc = peak_counts + b * 16;
    const int na = min(pc[ja], LWOP_MAX_PEAKS), nb = min(pc[jb], LWOP_MAX_PEAKS);
    double *out = limbs + ((long long)b * LWOP_NUM_LIMBS + t) * LWOP_MAX_PEAKS * 5;
    if (na == 0 || nb == 0) {                                     // (:224-228)
        if (tid == 0) counts[b * LWOP_COUNTS + 18 + t] = 0;
        return;
    }
    for (int i = tid; i < na; i += kLimbThreads) {
        s_best[i] = 0.0;                                          // best_score = 0.0 (:243)
        s_bestj[i] = 0x7fffffff;
    }
    int ida = 0, idb = 0;                                         // unique ids count channels then peaks (:135,183)
    for (int c = 0; c < ja; ++c) ida += min(pc[c], LWOP_MAX_PEAKS);
    for (int c = 0; c < jb; ++c) idb += min(pc[c], LWOP_MAX_PEAKS);
    const float *pa = peaks + ((long long)b * LWOP_NUM_JOINTS + ja) * LWOP_MAX_PEAKS * 4;
    const float *pb = peaks + ((long long)b * LWOP_NUM_JOINTS + jb) * LWOP_MAX_PEAKS * 4;
    const float *pafb = paf + (long long)b * (direct ? (long long)H * W : (long long)h * w) * LWOP_NUM_PAFS;
    const int cx_ch = 2 * t;                                      // paf_xy_coords_per_limb (:23-24): (2t, 2t+1)
    const double scale_x = 1.0 / ((double)W / (double)w);         // cv2: 1. / (dsize / ssize)
    const double scale_y = 1.0 / ((double)H / (double)h);
    const int total = na * nb;
    __syncthreads();

    for (int chunk0 = 0; chunk0 < total; chunk0 += kPairChunk) {
        const int npairs = min(kPairChunk, total - chunk0);
        // ---- one thread per (pair, sample): score_k = paf(x_k, y_k) . unit(limb) ----
        for (int item = tid; item < npairs * nsamp; item += kLimbThreads) {
            const int pi = item / nsamp, k = item - pi * nsamp;
            const int gp = chunk0 + pi;
            const int i = gp / nb, j = gp - i * nb;
            const double ax = pa[i * 4 + 0], ay = pa[i * 4 + 1];
            const double bx = pb[j * 4 + 0], by = pb[j * 4 + 1];
            const double dx = bx - ax, dy = by - ay;              // limb_dir (:249)
            const double norm = __dadd_rn(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), 1e-8);   // (:255)
            const double ux = dx / norm, uy = dy / norm;
            // np.linspace(start, stop, num): arange(num) * ((stop - start) / (num - 1)) + start, last point = stop
            // (num == 1: the single point is `start`)
            const double div = (double)(nsamp - 1);
            const double stepx = dx / div, stepy = dy / div;
            const bool last = nsamp > 1 && k == nsamp - 1;
            const double fxk = last ? bx : (nsamp > 1 ? __dadd_rn(__dmul_rn((double)k, stepx), ax) : ax);
            const double fyk = last ? by : (nsamp > 1 ? __dadd_rn(__dmul_rn((double)k, stepy), ay) : ay);
            int X = (int)rint(fxk), Y = (int)rint(fyk);           // np.round: half to even (:260-264)
            X = clampi(X, 0, W - 1);                              // joints are inside the frame; guard only
            Y = clampi(Y, 0, H - 1);
            double sk;
            if (direct) {       // caller passed the up-sampled PAF (find_connected_joints stage API, :267-268)
                const float2 v = *reinterpret_cast<const float2 *>(pafb + ((long long)Y * W + X) * LWOP_NUM_PAFS + cx_ch);
                sk = __dadd_rn(__dmul_rn((double)v.x, ux), __dmul_rn((double)v.y, uy));
            } else {
                const AxisTaps tx = axis_taps(X, scale_x, w), ty = axis_taps(Y, scale_y, h);
                float rx[4], ry[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float *row = pafb + (long long)ty.idx[r] * w * LWOP_NUM_PAFS;
                    float vx[4], vy[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 v = *reinterpret_cast<const float2 *>(row + tx.idx[q] * LWOP_NUM_PAFS + cx_ch);
                        vx[q] = v.x;
                        vy[q] = v.y;
                    }
                    rx[r] = dot4(vx, tx.c);
                    ry[r] = dot4(vy, tx.c);
                }
                const double px = (double)dot4_rev(rx, ty.c), py = (double)dot4_rev(ry, ty.c);
                sk = __dadd_rn(__dmul_rn(px, ux), __dmul_rn(py, uy));   // intermed_paf.dot(limb_dir) (:270)
            }
            s_val[pi * nsamp + k] = sk;
        }
        __syncthreads();
        // ---- one thread per pair: the acceptance criteria (:271-283) ----
        for (int pi = tid; pi < npairs; pi += kLimbThreads) {
            const int gp = chunk0 + pi;
            const int i = gp / nb, j = gp - i * nb;
            const double dx = (double)pb[j * 4 + 0] - (double)pa[i * 4 + 0], dy = (double)pb[j * 4 + 1] - (double)pa[i * 4 + 1];
            const double *sv = s_val + pi * nsamp;
            int above = 0;
            for (int k = 0; k < nsamp; ++k) above += sv[k] > thre2 ? 1 : 0;
            const double mean = numpy_sum(sv, nsamp) / (double)nsamp;                 // score_intermed_pts.mean() (:271)
            // long_dist_thresh = [W, H] * max_dist_thresh (:209,250)
            const bool skip = fabs(dx) > __dmul_rn((double)W, max_dist_thresh) || fabs(dy) > __dmul_rn((double)H, max_dist_thresh);
            s_mean[pi] = mean;
            // criterion 1: count > 0.8 * num_intermed_pts (:276-277), criterion 2: mean > thre2 (:283)
            s_ok[pi] = (!skip && (double)above > __dmul_rn(0.8, (double)nsamp) && mean > thre2) ? 1 : 0;
        }
        __syncthreads();
        // ---- one thread per source joint: best destination so far, j ascending, strict > (first j wins ties, :285) ----
        const int i0 = chunk0 / nb, i1 = (chunk0 + npairs - 1) / nb;
        for (int i = i0 + tid; i <= i1; i += kLimbThreads) {
            const int jlo = i == i0 ? chunk0 - i0 * nb : 0;
            const int jhi = i == i1 ? (chunk0 + npairs - 1) - i1 * nb : nb - 1;
            double best = s_best[i];
            int best_j = s_bestj[i];
            for (int j = jlo; j <= jhi; ++j) {
                const int pi = i * nb + j - chunk0;
                if (s_ok[pi] && s_mean[pi] > best) {
                    best = s_mean[pi];
                    best_j = j;
                }
            }
            s_best[i] = best;
            s_bestj[i] = best_j;
        }
        __syncthreads();
    }
    if (tid == 0) {                                               // rows in src order (:297)
        int r = 0;
        for (int i = 0; i < na; ++i)
            if (s_bestj[i] != 0x7fffffff) {
                double *o = out + r * 5;
                o[0] = (double)(ida + i);
                o[1] = (double)(idb + s_bestj[i]);
                o[2] = s_best[i];
                o[3] = (double)i;
                o[4] = (double)s_bestj[i];
                ++r;
            }
        counts[b * LWOP_COUNTS + 18 + t] = r;
    }
}

// ---------------------------------------------------------------------------
// K3: per image (one warp): joint list, greedy grouping, keypoint table.
//
// Two kernels.  group_fast_kernel (documented at its definition) covers images whose working list never exceeds 128
// persons; an image that needs more sets counts[b][31] and group_kernel (list in shared memory with a global spill,
// any length) redoes it.
// ---------------------------------------------------------------------------
constexpr int kLimbSrcC[13] = {0, 1, 3, 4, 6, 7, 9, 10, 12, 13, 13, 13, 13};   // pose_decode.py:17-18, compile-time copy
constexpr int kLimbDstC[13] = {1, 2, 4, 5, 7, 8, 10, 11, 13, 0, 3, 6, 9};
constexpr int kGroupWarps = 2;      // images per CTA of the fast kernel

// joint id -> (joint type, index within the type) through the exclusive prefix of the per-type counts; ids outside
// the list (only a caller of the stage API can pass them) are clamped so that no lookup leaves the tables
__device__ __forceinline__ void joint_ref(const int *off, int id, int &type, int &idx) {
    id = max(0, min(id, off[LWOP_NUM_JOINTS] - 1));
    int t = 0;
#pragma unroll
    for (int c = 1; c < LWOP_NUM_JOINTS; ++c) t += id >= off[c] ? 1 : 0;
    type = t;
    idx = max(0, min(id - off[t], LWOP_MAX_PEAKS - 1));
}

constexpr int kSlots = 8;           // the fast grouping kernel holds 32 * kSlots working-list rows
constexpr int kFastPersons = 32 * kSlots;

// keypoint table of one kept person (:403-429): for every limb type whose both ends are set, (x, y, 1) of both joints.
// All coordinate loads are issued first (26 independent loads), the stores follow in the reference's order.
__device__ __forceinline__ void keypoint_row(const short *ids /* column c of this person at ids[c * kFastPersons] */,
                                             const int *off, const float *__restrict__ pk, float *__restrict__ ko) {
    float2 xy[2 * LWOP_NUM_LIMBS];
    int jt[2 * LWOP_NUM_LIMBS];
#pragma unroll
    for (int t = 0; t < LWOP_NUM_LIMBS; ++t) {
        const int ia = ids[c_limb_src[t] * kFastPersons], ib = ids[c_limb_dst[t] * kFastPersons];
        const bool on = ia >= 0 && ib >= 0;                       // `-1 in joint_indices` skips the limb (:411)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int ty, ix;
            joint_ref(off, e == 0 ? ia : ib, ty, ix);
            jt[2 * t + e] = on ? ty : -1;
            xy[2 * t + e] = __ldg(reinterpret_cast<const float2 *>(pk + (ty * LWOP_MAX_PEAKS + ix) * 4));
        }
    }
#pragma unroll
    for (int k = 0; k < 2 * LWOP_NUM_LIMBS; ++k)
        if (jt[k] >= 0) {
            ko[jt[k] * 3 + 0] = xy[k].x;
            ko[jt[k] * 3 + 1] = xy[k].y;
            ko[jt[k] * 3 + 2] = 1.0f;
        }
}

// per-type offsets, counts[b][0], [4..17]; one warp
__device__ __forceinline__ int group_offsets(int lane, const int *pc, int *cnt, int *off, bool &peak_overflow) {
    const int raw = lane < LWOP_NUM_JOINTS ? pc[lane] : 0;
    const int my_pc = min(raw, LWOP_MAX_PEAKS);
    peak_overflow = __any_sync(0xffffffffu, raw > LWOP_MAX_PEAKS);
    int incl = my_pc;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane < LWOP_NUM_JOINTS) {
        off[lane] = incl - my_pc;
        cnt[4 + lane] = my_pc;
    }
    if (lane == LWOP_NUM_JOINTS - 1) {
        off[LWOP_NUM_JOINTS] = incl;
        cnt[0] = incl;
    }
    __syncwarp();
    return my_pc;
}

// the joint list (x, y, score, id, 0, type) (:481-482); one warp, one joint per lane and step (independent loads)
__device__ __forceinline__ void write_joint_list(int lane, const int *off, const float *__restrict__ pk, double *__restrict__ jl) {
    const int total = off[LWOP_NUM_JOINTS];
    for (int id = lane; id < total; id += 32) {
        int c, i;
        joint_ref(off, id, c, i);
        const float4 pv = __ldg(reinterpret_cast<const float4 *>(pk + (c * LWOP_MAX_PEAKS + i) * 4));
        double *o = jl + (long long)id * 6;
        o[0] = pv.x;
        o[1] = pv.y;
        o[2] = pv.z;
        o[3] = (double)id;
        o[4] = 0.0;
        o[5] = (double)c;
    }
}

// One warp per image.  Working list (at most 256 rows, in list order), all of it in shared memory:
//   joint ids    transposed: s_id[column][row] (-1 = unset; -2 in unused and buried rows: they never match) -- lane = row
//                for the membership test of the serial loop (:329-331: two conflict-free loads, two compares and one
//                ballot per 32 rows), lane = column for merges and new rows;
//   score, count one entry per row, updated by a single lane.
// A lone warp issues one instruction every ~6 cycles, so the loop is written for instruction count, not for latency
// (ncu: the register-resident forms of this kernel spent their time in the select chains that keep per-lane arrays in
// registers): ~30 instructions for a connection with one match, ~25 more for a merge.
// The first 32 connections of all 13 limb types (ids, score, heat-map scores of the two joints) are staged in shared
// memory before the serial loop starts: three dependent round trips to memory per image instead of three per type.
struct ConnStage {
    int a[LWOP_NUM_LIMBS][32], b[LWOP_NUM_LIMBS][32];
    float sa[LWOP_NUM_LIMBS][32], sb[LWOP_NUM_LIMBS][32];
    double s[LWOP_NUM_LIMBS][32];
};

__global__ void __launch_bounds__(32 * kGroupWarps)
group_fast_kernel(const float *__restrict__ peaks, const int *__restrict__ peak_counts, const double *__restrict__ limbs,
                  int *__restrict__ counts, double *__restrict__ joints, double *__restrict__ persons,
                  float *__restrict__ keypoints, int B) {
    __shared__ int s_off[kGroupWarps][LWOP_NUM_JOINTS + 2];
    __shared__ int s_nconn[kGroupWarps][16];
    __shared__ short s_ids[kGroupWarps][LWOP_NUM_JOINTS][kFastPersons];
    __shared__ double s_score[kGroupWarps][kFastPersons];
    __shared__ int s_count[kGroupWarps][kFastPersons];
    __shared__ ConnStage s_conn[kGroupWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kGroupWarps + warp;
    pdl_wait();
    pdl_launch_dependents();
    if (b >= B) return;
    int *cnt = counts + b * LWOP_COUNTS;
    int *off = s_off[warp];
    short(*id)[kFastPersons] = s_ids[warp];
    ConnStage &cs = s_conn[warp];
    const float *pk = peaks + (long long)b * LWOP_NUM_JOINTS * LWOP_MAX_PEAKS * 4;
    const double *limbs_b = limbs + (long long)b * LWOP_NUM_LIMBS * LWOP_MAX_PEAKS * 5;
    if (lane < LWOP_NUM_LIMBS) s_nconn[warp][lane] = min(cnt[18 + lane], LWOP_MAX_PEAKS);
    bool peak_overflow;
    const int my_pc = group_offsets(lane, peak_counts + b * 16, cnt, off, peak_overflow);
    const int *nconn = s_nconn[warp];
    // stage connection `lane` of every type: all row loads first, then all the joint-score lookups
    {
        int ca[LWOP_NUM_LIMBS], cb[LWOP_NUM_LIMBS];
#pragma unroll
        for (int t = 0; t < LWOP_NUM_LIMBS; ++t) {
            ca[t] = cb[t] = 0;
            if (lane < nconn[t]) {
                const double *row = limbs_b + ((long long)t * LWOP_MAX_PEAKS + lane) * 5;
                ca[t] = (int)row[0];
                cb[t] = (int)row[1];
                cs.s[t][lane] = row[2];
            }
        }
#pragma unroll
        for (int t = 0; t < LWOP_NUM_LIMBS; ++t)
            if (lane < nconn[t]) {
                int ty, ix;
                joint_ref(off, ca[t], ty, ix);
                cs.sa[t][lane] = pk[(ty * LWOP_MAX_PEAKS + ix) * 4 + 2];
                joint_ref(off, cb[t], ty, ix);
                cs.sb[t][lane] = pk[(ty * LWOP_MAX_PEAKS + ix) * 4 + 2];
                cs.a[t][lane] = ca[t];
                cs.b[t][lane] = cb[t];
            }
    }
    for (int c = 0; c < LWOP_NUM_JOINTS; ++c)
#pragma unroll
        for (int sl = 0; sl < kSlots; ++sl) id[c][sl * 32 + lane] = -2;       // unused rows never match an id
    write_joint_list(lane, off, pk, joints + (long long)b * LWOP_MAX_JOINTS * 6);
    __syncwarp();

    double *score = s_score[warp];
    int *count = s_count[warp];
    // The reference's list.pop() is a tombstone here: a merged-away person keeps its row (ids -2, count -1, so it never
    // matches and never survives the final filter).  Rows are only ever appended, so row order = list order, which is
    // all the "first two matches" rule and the output order depend on.  ns = rows used so far.
    int ns = 0, total_conn = 0;
    bool handover = false;
#pragma unroll 1
    for (int t = 0; t < LWOP_NUM_LIMBS && !handover; ++t) {
        const int ja = c_limb_src[t], jb = c_limb_dst[t];
        const int n = nconn[t];
        total_conn += n;
        const short *ida = id[ja];
        short *idb = id[jb];
#pragma unroll 1
        for (int r = 0; r < n; ++r) {
            if (r >= 32 && (r & 31) == 0) {          // beyond the staged 32: fetch the next chunk of this type
                if (r + lane < n) {
                    const double *row = limbs_b + ((long long)t * LWOP_MAX_PEAKS + r + lane) * 5;
                    const int ca = (int)row[0], cb = (int)row[1];
                    cs.s[t][lane] = row[2];
                    int ty, ix;
                    joint_ref(off, ca, ty, ix);
                    cs.sa[t][lane] = pk[(ty * LWOP_MAX_PEAKS + ix) * 4 + 2];
                    joint_ref(off, cb, ty, ix);
                    cs.sb[t][lane] = pk[(ty * LWOP_MAX_PEAKS + ix) * 4 + 2];
                    cs.a[t][lane] = ca;
                    cs.b[t][lane] = cb;
                }
                __syncwarp();
            }
            const int e = r & 31;
            const int a_id = cs.a[t][e], b_id = cs.b[t][e];
            // persons that already hold one of the two joints (:329-331), in list order; rows >= ns hold -2
            int nm = 0, m0 = -1, m1 = -1;
#pragma unroll
            for (int sl = 0; sl < kSlots; ++sl) {
                if (sl > 0 && ns <= sl * 32) break;
                const int ra = ida[sl * 32 + lane], rb = idb[sl * 32 + lane];       // both loads in flight together
                const unsigned mask = __ballot_sync(0xffffffffu, (ra == a_id) | (rb == b_id));
                if (mask) {
                    if (nm == 0) {
                        m0 = sl * 32 + __ffs(mask) - 1;
                        const unsigned rest = mask & (mask - 1);
                        if (rest) m1 = sl * 32 + __ffs(rest) - 1;
                    } else if (nm == 1) {
                        m1 = sl * 32 + __ffs(mask) - 1;
                    }
                    nm += __popc(mask);
                }
            }
            if (nm == 1 || nm == 2) {
                bool merge = false;
                if (nm == 2) {                       // do the two persons share a set joint?  (:351-352)
                    const bool both = lane < LWOP_NUM_JOINTS && id[lane][m0] >= 0 && id[lane][m1] >= 0;
                    merge = !__any_sync(0xffffffffu, both);
                }
                if (!merge) {
                    // one match whose dst column differs (:336-347), or two overlapping persons: the first one,
                    // unconditionally (:362-366)
                    if ((nm == 2 || idb[m0] != b_id) && lane == 0) {
                        idb[m0] = (short)b_id;
                        count[m0] += 1;
                        score[m0] = __dadd_rn(score[m0], __dadd_rn((double)cs.sb[t][e], cs.s[t][e]));
                    }
                } else {
                    // disjoint persons: p1[:14] += p2[:14] + 1, scores and counts add, + s, list.pop(m1) (:353-361)
                    if (lane < LWOP_NUM_JOINTS) {
                        id[lane][m0] = (short)(id[lane][m0] + id[lane][m1] + 1);
                        id[lane][m1] = -2;
                    } else if (lane == 16) {
                        score[m0] = __dadd_rn(__dadd_rn(score[m0], score[m1]), cs.s[t][e]);
                        count[m0] += count[m1];
                        count[m1] = -1;
                    }
                }
            } else {
                // 0 or >= 3 matches: new person (:367-379)
                if (ns >= kFastPersons) {
                    handover = true;
                    break;
                }
                if (lane < LWOP_NUM_JOINTS) {
                    id[lane][ns] = (short)(lane == ja ? a_id : (lane == jb ? b_id : -1));
                } else if (lane == 16) {
                    count[ns] = 2;
                    score[ns] = __dadd_rn(__dadd_rn(__dadd_rn(0.0, (double)cs.sa[t][e]), (double)cs.sb[t][e]), cs.s[t][e]);   // Python sum() order
                }
                ++ns;
            }
            __syncwarp();                            // this connection's row updates are visible to every lane
        }
    }
    if (handover) {                  // more than 256 rows: group_kernel takes this image
        if (lane == 0) cnt[31] = 1;
        return;
    }
    // drop persons with < 3 parts or mean score < 0.2 (:382-391; tombstones have count -1); keypoint table (:403-429)
    int kept = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < ns; k0 += 32) {
        const int k = k0 + lane;
        const int c15 = k < ns ? count[k] : -1;
        const double c14 = k < ns ? score[k] : 0.0;
        const bool keep = !(c15 < 3 || (c14 / (double)c15) < 0.2);
        const unsigned kmask = __ballot_sync(0xffffffffu, keep);
        const int rank = kept + __popc(kmask & ((1u << lane) - 1u));
        kept += __popc(kmask);
        if (keep) {
            double *po = persons + ((long long)b * LWOP_MAX_PERSONS + rank) * 16;
#pragma unroll
            for (int c = 0; c < LWOP_NUM_JOINTS; ++c) po[c] = (double)id[c][k];
            po[14] = c14;
            po[15] = (double)c15;
            float *ko = keypoints + ((long long)b * LWOP_MAX_PERSONS + rank) * 42;
#pragma unroll
            for (int e = 0; e < 42; ++e) ko[e] = 0.0f;
            keypoint_row(&id[0][k], off, pk, ko);
        }
    }
    if (lane == 0) {
        cnt[1] = kept;                   // <= 256 = LWOP_MAX_PERSONS
        cnt[2] = peak_overflow ? 1 : 0;
        cnt[3] = total_conn;
        cnt[31] = 0;
    }
}

constexpr int kSmemPersons = 128;   // working persons kept in shared memory; the rest spill to the global scratch

// The general form (working list of any length): one warp per image that group_fast_kernel handed over
// (counts[b][31] != 0), or every image when `force` is set.
__global__ void __launch_bounds__(32)
group_kernel(const float *__restrict__ peaks, const int *__restrict__ peak_counts, const double *__restrict__ limbs,
             int *__restrict__ counts, double *__restrict__ joints, double *__restrict__ persons,
             float *__restrict__ keypoints, double *__restrict__ work, int force) {
    const int b = blockIdx.x, lane = threadIdx.x;
    // working list of persons before the final filter: at most one new person per connection
    __shared__ double Ps[kSmemPersons][16];
    __shared__ double s_cs[LWOP_MAX_PEAKS];              // connection scores of the current limb type
    __shared__ short s_ca[LWOP_MAX_PEAKS], s_cb[LWOP_MAX_PEAKS];   // their (src_id, dst_id)
    __shared__ float s_sa[LWOP_MAX_PEAKS], s_sb[LWOP_MAX_PEAKS];   // heat-map scores of the two joints (float32 values)
    // columns ja / jb of every working person for the current limb type, as integers: the membership test of the
    // serial loop (:329-331) reads these instead of two doubles per person; kept in step with the rows below
    __shared__ short s_pa[kMaxWorkPersons], s_pb[kMaxWorkPersons];
    __shared__ int s_off[LWOP_NUM_JOINTS + 2];
    pdl_wait();
    pdl_launch_dependents();
    int *cnt = counts + b * LWOP_COUNTS;
    if (!force && cnt[31] == 0) return;
    double *Pg = work + (long long)b * kMaxWorkPersons * 16;
    // explicit address spaces (a generic pointer would turn the shared-memory accesses into generic loads)
    auto row_ld = [&](int k, int c) -> double { return k < kSmemPersons ? Ps[k][c] : Pg[(long long)(k - kSmemPersons) * 16 + c]; };
    auto row_st = [&](int k, int c, double v) {
        if (k < kSmemPersons) Ps[k][c] = v;
        else Pg[(long long)(k - kSmemPersons) * 16 + c] = v;
    };

    const float *pk = peaks + (long long)b * LWOP_NUM_JOINTS * LWOP_MAX_PEAKS * 4;
    const int my_nconn = lane < LWOP_NUM_LIMBS ? min(cnt[18 + lane], LWOP_MAX_PEAKS) : 0;
    bool peak_overflow;
    const int my_pc = group_offsets(lane, peak_counts + b * 16, cnt, s_off, peak_overflow);
    write_joint_list(lane, s_off, pk, joints + (long long)b * LWOP_MAX_JOINTS * 6);
    __syncwarp();

    int np = 0, total_conn = 0;
    bool overflow = false;
    for (int t = 0; t < LWOP_NUM_LIMBS; ++t) {
        const int ja = c_limb_src[t], jb = c_limb_dst[t];
        const int nconn = __shfl_sync(0xffffffffu, my_nconn, t);
        total_conn += nconn;
        const double *rows = limbs + ((long long)b * LWOP_NUM_LIMBS + t) * LWOP_MAX_PEAKS * 5;
        for (int e = lane; e < nconn; e += 32) {
            const int ca = (int)rows[e * 5 + 0], cb = (int)rows[e * 5 + 1];
            s_ca[e] = (short)ca;
            s_cb[e] = (short)cb;
            s_cs[e] = rows[e * 5 + 2];
            int ty, ix;
            joint_ref(s_off, ca, ty, ix);
            s_sa[e] = pk[(ty * LWOP_MAX_PEAKS + ix) * 4 + 2];
            joint_ref(s_off, cb, ty, ix);
            s_sb[e] = pk[(ty * LWOP_MAX_PEAKS + ix) * 4 + 2];
        }
        for (int k = lane; k < np; k += 32) {                     // this type's two columns of every working person
            s_pa[k] = (short)row_ld(k, ja);
            s_pb[k] = (short)row_ld(k, jb);
        }
        __syncwarp();
        for (int r = 0; r < nconn; ++r) {
            const int a_id = s_ca[r], b_id = s_cb[r];
            const double s = s_cs[r];
            // persons that already hold one of the two joints (:329-331), in list order
            int nm = 0, m0 = -1, m1 = -1;
            for (int base = 0; base < np; base += 32) {
                const int k = base + lane;
                const bool hit = k < np && (s_pa[k] == a_id || s_pb[k] == b_id);
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (mask) {
                    if (nm == 0) {
                        m0 = base + __ffs(mask) - 1;
                        const unsigned rest = mask & (mask - 1);
                        if (rest) m1 = base + __ffs(rest) - 1;
                    } else if (nm == 1) {
                        m1 = base + __ffs(mask) - 1;
                    }
                    nm += __popc(mask);
                }
            }
            const double js_b = (double)s_sb[r];
            const double bd = (double)b_id;
            if (nm == 1) {                                        // (:336-347)
                if (lane == 0 && s_pb[m0] != b_id) {
                    row_st(m0, jb, bd);
                    row_st(m0, 15, row_ld(m0, 15) + 1.0);
                    row_st(m0, 14, row_ld(m0, 14) + __dadd_rn(js_b, s));
                    s_pb[m0] = (short)b_id;
                }
            } else if (nm == 2) {                                 // (:348-366)
                double v0 = 0.0, v1 = 0.0;
                if (lane < 16) {
                    v0 = row_ld(m0, lane);
                    v1 = row_ld(m1, lane);
                }
                const bool share = __any_sync(0xffffffffu, lane < 14 && v0 >= 0.0 && v1 >= 0.0);
                if (!share) {
                    if (lane < 14) v0 += __dadd_rn(v1, 1.0);
                    if (lane == 14) v0 = __dadd_rn(__dadd_rn(v0, v1), s);
                    if (lane == 15) v0 += v1;
                    if (lane < 16) row_st(m0, lane, v0);
                    if (lane == ja) s_pa[m0] = (short)v0;
                    if (lane == jb) s_pb[m0] = (short)v0;
                    __syncwarp();
                    // list.pop(m1): rows and column caches move up by one, 32 persons per step (read all, then write all)
                    for (int base = m1; base + 1 < np; base += 32) {
                        const int k = base + lane;
                        const bool live = k + 1 < np;
                        double rv[16];
                        short ca = 0, cb = 0;
                        if (live) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) rv[c] = row_ld(k + 1, c);
                            ca = s_pa[k + 1];
                            cb = s_pb[k + 1];
                        }
                        __syncwarp();
                        if (live) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) row_st(k, c, rv[c]);
                            s_pa[k] = ca;
                            s_pb[k] = cb;
                        }
                        __syncwarp();
                    }
                    --np;
                } else if (lane == 0) {
                    row_st(m0, jb, bd);
                    row_st(m0, 15, row_ld(m0, 15) + 1.0);
                    row_st(m0, 14, row_ld(m0, 14) + __dadd_rn(js_b, s));
                    s_pb[m0] = (short)b_id;
                }
            } else {                                              // 0 or >= 3 matches: new person (:367-379)
                if (np < kMaxWorkPersons) {
                    if (lane < 14) row_st(np, lane, lane == ja ? (double)a_id : (lane == jb ? bd : -1.0));
                    if (lane == 15) row_st(np, 15, 2.0);
                    if (lane == 14) {
                        const double js_a = (double)s_sa[r];
                        row_st(np, 14, __dadd_rn(__dadd_rn(__dadd_rn(0.0, js_a), js_b), s));
                    }
                    if (lane == 0) {
                        s_pa[np] = (short)a_id;
                        s_pb[np] = (short)b_id;
                    }
                    ++np;
                } else {
                    overflow = true;
                }
            }
            __syncwarp();
        }
    }
    // drop persons with < 3 parts or mean score < 0.2 (:382-391); keypoint table (:403-429)
    double *po = persons + (long long)b * LWOP_MAX_PERSONS * 16;
    float *ko = keypoints + (long long)b * LWOP_MAX_PERSONS * 42;
    int kept = 0;
    for (int k = 0; k < np; ++k) {
        const double r15 = row_ld(k, 15);
        const bool drop = r15 < 3.0 || (row_ld(k, 14) / r15) < 0.2;
        if (drop) continue;
        if (kept >= LWOP_MAX_PERSONS) {
            overflow = true;
            break;
        }
        if (lane < 16) po[kept * 16 + lane] = row_ld(k, lane);
        for (int e = lane; e < 42; e += 32) ko[kept * 42 + e] = 0.0f;
        __syncwarp();
        if (lane == 0) {
            for (int t = 0; t < LWOP_NUM_LIMBS; ++t) {
                const double ia = row_ld(k, c_limb_src[t]), ib = row_ld(k, c_limb_dst[t]);
                if (ia < 0.0 || ib < 0.0) continue;               // `-1 in joint_indices` (:411)
                const double ids[2] = {ia, ib};
                for (int e = 0; e < 2; ++e) {
                    int jt, ix;
                    joint_ref(s_off, (int)ids[e], jt, ix);
                    const float2 xy = *reinterpret_cast<const float2 *>(pk + (jt * LWOP_MAX_PEAKS + ix) * 4);
                    ko[kept * 42 + jt * 3 + 0] = xy.x;
                    ko[kept * 42 + jt * 3 + 1] = xy.y;
                    ko[kept * 42 + jt * 3 + 2] = 1.0f;
                }
            }
        }
        __syncwarp();
        ++kept;
    }
    if (lane == 0) {
        cnt[1] = kept;
        cnt[2] = (peak_overflow ? 1 : 0) | (overflow ? 2 : 0);
        cnt[3] = total_conn;
        cnt[31] = 0;
    }
}

// K4: exclusive prefix over images of (joints, persons, connections), one block
__global__ void __launch_bounds__(1024)
pack_offsets_kernel(const int *__restrict__ counts, int B, int *__restrict__ offsets) {
    __shared__ int s[3][1024];
    const int tid = threadIdx.x;
    int carry[3] = {0, 0, 0};
    pdl_wait();
    pdl_launch_dependents();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + tid;
        const int idx[3] = {0, 1, 3};
        int v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = b < B ? counts[b * LWOP_COUNTS + idx[k]] : 0;
            s[k][tid] = v[k];
        }
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] = tid >= o ? s[k][tid - o] : 0;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 3; ++k) s[k][tid] += t[k];
            __syncthreads();
        }
        if (b < B) {
#pragma unroll
            for (int k = 0; k < 3; ++k) offsets[b * 4 + k] = carry[k] + s[k][tid] - v[k];
            offsets[b * 4 + 3] = 0;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) carry[k] += s[k][1023];
        __syncthreads();
    }
    if (tid == 0) {
        offsets[B * 4 + 0] = carry[0];
        offsets[B * 4 + 1] = carry[1];
        offsets[B * 4 + 2] = carry[2];
        offsets[B * 4 + 3] = 0;
    }
}

__global__ void __launch_bounds__(128)
pack_rows_kernel(lwop_decode_tables t, const int *__restrict__ offsets, double *__restrict__ joints_p,
                 double *__restrict__ limbs_p, double *__restrict__ persons_p, float *__restrict__ keypoints_p,
                 long long cap_j, long long cap_l, long long cap_p) {
    // rows that would land beyond a packed buffer's capacity are dropped; the host detects that from the counts
    const int b = blockIdx.x, tid = threadIdx.x;
    pdl_wait();
    const int *cnt = t.counts + b * LWOP_COUNTS;
    const int *off = offsets + b * 4;
    const int nj = cnt[0], np = cnt[1];
    const double *js = t.joints + (long long)b * LWOP_MAX_JOINTS * 6;
    if ((long long)off[0] + nj <= cap_j)
        for (int i = tid; i < nj * 6; i += blockDim.x) joints_p[(long long)off[0] * 6 + i] = js[i];
    if ((long long)off[1] + np <= cap_p) {
        const double *ps = t.persons + (long long)b * LWOP_MAX_PERSONS * 16;
        for (int i = tid; i < np * 16; i += blockDim.x) persons_p[(long long)off[1] * 16 + i] = ps[i];
        const float *ks = t.keypoints + (long long)b * LWOP_MAX_PERSONS * 42;
        for (int i = tid; i < np * 42; i += blockDim.x) keypoints_p[(long long)off[1] * 42 + i] = ks[i];
    }
    long long row = off[2];
    for (int l = 0; l < LWOP_NUM_LIMBS; ++l) {
        const int nc = cnt[18 + l];
        const double *ls = t.limbs + ((long long)b * LWOP_NUM_LIMBS + l) * LWOP_MAX_PEAKS * 5;
        if (row + nc <= cap_l)
            for (int i = tid; i < nc * 5; i += blockDim.x) limbs_p[row * 5 + i] = ls[i];
        row += nc;
    }
}

__global__ void zero_counts_kernel(int *counts, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = 0;
}

// stage API (lwop_nms): clamps the per-channel peak counts to the table capacity and reports the overflow in
// counts[b][2] (the fused chain leaves both to the grouping kernel)
__global__ void peaks_finalize_kernel(int *peak_counts, int *counts, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int over = 0;
    for (int c = 0; c < LWOP_NUM_JOINTS; ++c) {
        const int n = peak_counts[b * 16 + c];
        if (n > LWOP_MAX_PEAKS) {
            over = 1;
            peak_counts[b * 16 + c] = LWOP_MAX_PEAKS;
        }
    }
    if (over) counts[b * LWOP_COUNTS + 2] |= 1;
}

size_t peaks_smem_bytes(int npix, int mode) {
    size_t smem = (size_t)npix * 4 + (size_t)((npix + 1) & ~1) * 2 + (size_t)npix;      // map + waits + per-pixel state
    smem = ((smem + 15) & ~(size_t)15) + 16;
    if (mode & LWOP_PEAKS_GAUSS) smem += 2 * (size_t)kRefineMax * kRefineMax * sizeof(float) + 16;
    return smem;
}

cudaError_t prepare_peaks(size_t smem) {
    if (smem > 160 * 1024) return cudaErrorInvalidValue;
    if (smem > 16 * 1024) return cudaFuncSetAttribute(peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    return cudaSuccess;
}

cudaError_t prepare_limbs(size_t smem) {
    if (smem > 32 * 1024) return cudaFuncSetAttribute(limbs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    return cudaSuccess;
}

}  // namespace

size_t decode_scratch_work_bytes(int B) { return (size_t)B * kMaxWorkPersons * 16 * sizeof(double); }

size_t decode_scratch_peaks_bytes(int B) {
    return (size_t)B * LWOP_NUM_JOINTS * LWOP_MAX_PEAKS * 4 * sizeof(float);
}

cudaError_t launch_decode_pack(const lwop_decode_tables &t, int B, int *offsets, double *joints_packed,
                               double *limbs_packed, double *persons_packed, float *keypoints_packed,
                               cudaStream_t s, long long *launches, long long cap_j, long long cap_l, long long cap_p) {
    cudaError_t e = launch_chain(pack_offsets_kernel, 1, 1024, 0, s, 1, (const int *)t.counts, B, offsets);
    if (e != cudaSuccess) return e;
    e = launch_chain(pack_rows_kernel, (unsigned)B, 128, 0, s, 1, t, (const int *)offsets, joints_packed, limbs_packed,
                     persons_packed, keypoints_packed, cap_j, cap_l, cap_p);
    if (launches) *launches += 2;
    return e;
}

// The whole chain: peaks -> limbs -> grouping (fast, then the general kernel for the images the fast one handed
// over).  Every kernel is launched with the programmatic-dependent-launch attribute and waits (griddepcontrol.wait)
// before it touches what the kernel before it produced, so the chain can follow the forward's last kernel inside one
// CUDA graph without launch gaps.
cudaError_t launch_decode(const DecodeArgs &a, cudaStream_t s, long long *launches) {
    const int npix = a.h * a.w;
    if (npix > 16384) return cudaErrorInvalidValue;
    const size_t smem = peaks_smem_bytes(npix, 0);
    cudaError_t e = prepare_peaks(smem);
    if (e != cudaSuccess) return e;
    const double factor = (double)a.H / (double)a.h;              // scale = img_H / heat_H (:462)
    const size_t lsmem = (size_t)kPairChunk * 10 * sizeof(double);
    e = launch_chain(peaks_kernel, dim3(LWOP_NUM_JOINTS, a.B), 256, smem, s, 1, a.heat, a.h, a.w, (int)LWOP_NUM_JOINTS, a.thre1,
                     factor, 0, a.peaks, a.peak_counts);
    if (e != cudaSuccess) return e;
    e = launch_chain(limbs_kernel, dim3(LWOP_NUM_LIMBS, a.B), kLimbThreads, lsmem, s, 1, a.paf, a.h, a.w, a.H, a.W, a.thre2, 0,
                     10, 1.0, (const float *)a.peaks, (const int *)a.peak_counts, a.t.limbs, a.t.counts);
    if (e != cudaSuccess) return e;
    e = launch_chain(group_fast_kernel, (unsigned)((a.B + kGroupWarps - 1) / kGroupWarps), 32 * kGroupWarps, 0, s, 1,
                     (const float *)a.peaks, (const int *)a.peak_counts, (const double *)a.t.limbs, a.t.counts, a.t.joints,
                     a.t.persons, a.t.keypoints, a.B);
    if (e != cudaSuccess) return e;
    e = launch_chain(group_kernel, (unsigned)a.B, 32, 0, s, 1, (const float *)a.peaks, (const int *)a.peak_counts,
                     (const double *)a.t.limbs, a.t.counts, a.t.joints, a.t.persons, a.t.keypoints, a.work, 0);
    if (launches) *launches += 4;
    return e;
}

// ---- separately callable stages (reference NMS :107, find_connected_joints :188, group_limbs_of_same_person :304) ----
cudaError_t launch_decode_peaks(const float *heat, int B, int h, int w, double factor, float thre1, int mode,
                                float *peaks, int *peak_counts, int *counts, cudaStream_t s) {
    const int npix = h * w;
    if (npix > 16384) return cudaErrorInvalidValue;
    // the smoothed patch lives in shared memory: the clipped 5x5 patch times `factor` must fit 48 x 48
    if ((mode & LWOP_PEAKS_GAUSS) && (int)rint(5.0 * factor) > kRefineMax) return cudaErrorInvalidValue;
    const size_t smem = peaks_smem_bytes(npix, mode);
    cudaError_t e = prepare_peaks(smem);
    if (e != cudaSuccess) return e;
    const int ncounts = B * LWOP_COUNTS;
    zero_counts_kernel<<<(ncounts + 255) / 256, 256, 0, s>>>(counts, ncounts);
    peaks_kernel<<<dim3(LWOP_NUM_JOINTS, B), 256, smem, s>>>(heat, h, w, LWOP_NUM_JOINTS, thre1, factor, mode, peaks,
                                                             peak_counts);
    peaks_finalize_kernel<<<(B + 127) / 128, 128, 0, s>>>(peak_counts, counts, B);
    return cudaGetLastError();
}

cudaError_t launch_decode_limbs(const float *paf, int upsampled, int B, int h, int w, int H, int W, double thre2,
                                int num_intermed_pts, double max_dist_thresh, const float *peaks, const int *peak_counts,
                                double *limbs, int *counts, cudaStream_t s) {
    if (num_intermed_pts < 1 || num_intermed_pts > kMaxSamples) return cudaErrorInvalidValue;
    const size_t lsmem = (size_t)kPairChunk * num_intermed_pts * sizeof(double);
    cudaError_t e = prepare_limbs(lsmem);
    if (e != cudaSuccess) return e;
    limbs_kernel<<<dim3(LWOP_NUM_LIMBS, B), kLimbThreads, lsmem, s>>>(paf, h, w, H, W, thre2, upsampled ? 1 : 0,
                                                                     num_intermed_pts, max_dist_thresh, peaks, peak_counts,
                                                                     limbs, counts);
    return cudaGetLastError();
}

cudaError_t launch_decode_group(const float *peaks, const int *peak_counts, const double *limbs, int *counts,
                                double *joints, double *persons, float *keypoints, double *work, int B, cudaStream_t s) {
    group_fast_kernel<<<(B + kGroupWarps - 1) / kGroupWarps, 32 * kGroupWarps, 0, s>>>(peaks, peak_counts, limbs, counts, joints,
                                                                                     persons, keypoints, B);
    group_kernel<<<B, 32, 0, s>>>(peaks, peak_counts, limbs, counts, joints, persons, keypoints, work, 0);
    return cudaGetLastError();
}

}  // namespace lwop
