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
constexpr int kSamples = 10;        // num_intermed_pts, pose_decode.py:188
constexpr int kMaxPixPerThread = 64;
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

__global__ void __launch_bounds__(256)
peaks_kernel(const float *__restrict__ heat, int h, int w, int C, float thre1, double factor, int mode,
             float *__restrict__ peaks, int *__restrict__ peak_counts, int *__restrict__ counts) {
    extern __shared__ unsigned char smem_raw[];
    const int npix = h * w;
    float *val = reinterpret_cast<float *>(smem_raw);                 // the channel's map
    int *cand = reinterpret_cast<int *>(val + npix);                  // candidate pixels; index < 0 = suppressed
    __shared__ int s_nc, s_n;
    __shared__ int s_kept[LWOP_MAX_PEAKS];
    // per-warp scratch of the separable refinement: horizontal pass + tap tables
    __shared__ float s_h[8][5][kRefineMax];
    __shared__ float s_cx[8][kRefineMax][4], s_cy[8][kRefineMax][4];
    __shared__ signed char s_sx[8][kRefineMax], s_sy[8][kRefineMax];

    const int ch = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const float *src = heat + (long long)b * npix * C + ch;
    if (tid == 0) { s_nc = 0; s_n = 0; }
    // channel plane -> shared memory: the strided loads are issued in independent batches of 8 per thread so
    // their latencies overlap (the candidate scan runs afterwards on the shared-memory copy)
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
            if (p < npix) val[p] = v[k];
        }
    }
    __syncthreads();
    for (int p = tid; p < npix; p += blockDim.x)
        if (val[p] > thre1) cand[atomicAdd(&s_nc, 1)] = p;          // np.where(img > thre1), pose_decode.py:54
    __syncthreads();
    const int nc = s_nc;

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
        if (n > LWOP_MAX_PEAKS) {
            if (tid == 0) atomicOr(&counts[b * LWOP_COUNTS + 2], 1);
            n = LWOP_MAX_PEAKS;
        }
        __syncthreads();
    } else {
        // The greedy loop of find_peaks_v2 (:62-70) verbatim: take the best remaining candidate, drop every
        // remaining candidate within Euclidean distance 5 of it, repeat.  "Best" = larger score, exact ties
        // towards the larger row-major index.
        // One block-wide arg-max per kept peak (a single-warp variant was measured slower: the frames' heavy
        // channels have ~1000 candidates and 30-40 peaks, and they set the kernel's tail).
        __shared__ float s_wv[8];
        __shared__ int s_wp[8];
        __shared__ int s_best_p;
        // Each thread owns candidates tid, tid + 256, ...: it drops those within the radius of the peak kept in the
        // previous round and takes the arg-max of its survivors in the same pass (no barrier between the two).
        int prev = -1, prev_y = 0, prev_x = 0;
        while (true) {
            float bv = -INFINITY;
            int bp = -1;
            for (int i = tid; i < nc; i += blockDim.x) {
                const int p = cand[i];
                if (p < 0) continue;
                if (prev >= 0) {
                    const int py = p / w, px = p - py * w;
                    const int dy = py - prev_y, dx = px - prev_x;
                    if (dx * dx + dy * dy <= kRadius * kRadius) {          // keeps dis > dis_thres (:68); drops `prev` too
                        cand[i] = -1;
                        continue;
                    }
                }
                const float v = val[p];
                if (v > bv || (v == bv && p > bp)) { bv = v; bp = p; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (ov > bv || (ov == bv && op > bp)) { bv = ov; bp = op; }
            }
            if (lane == 0) { s_wv[warp] = bv; s_wp[warp] = bp; }
            __syncthreads();
            if (tid == 0) {
                float fv = s_wv[0];
                int fp = s_wp[0];
                for (int k = 1; k < nwarps; ++k)
                    if (s_wv[k] > fv || (s_wv[k] == fv && s_wp[k] > fp)) { fv = s_wv[k]; fp = s_wp[k]; }
                s_best_p = fp;
                if (fp >= 0 && n < LWOP_MAX_PEAKS) s_kept[n] = fp;
            }
            __syncthreads();
            const int best = s_best_p;
            if (best < 0) break;
            if (n >= LWOP_MAX_PEAKS) {                                    // more peaks than the table holds
                if (tid == 0) atomicOr(&counts[b * LWOP_COUNTS + 2], 1);
                break;
            }
            ++n;
            prev = best;
            prev_y = best / w;
            prev_x = best - prev_y * w;
        }
    }
    if (tid == 0) peak_counts[b * 16 + ch] = n;

    // refinement (:148-183): bicubic up-sampling of the clipped 5x5 patch by `factor` (float32 two-pass,
    // horizontal then vertical, cv2 arithmetic), argmax = first maximum in row-major order; warp per peak.
    const double scale = 1.0 / factor;              // cv2: scale_x = 1. / inv_scale_x
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
        float *gu = reinterpret_cast<float *>(cand + npix) + 4;
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
            for (int d = tid; d < dw; d += blockDim.x) {
                int sx;
                float fx, c[4];
                src_pos(d, scale, sx, fx);
                cubic_coeffs(fx, c);
                s_sx[0][d] = (signed char)sx;
#pragma unroll
                for (int k = 0; k < 4; ++k) s_cx[0][d][k] = c[k];
            }
            for (int d = tid; d < dh; d += blockDim.x) {
                int sy;
                float fy, c[4];
                src_pos(d, scale, sy, fy);
                cubic_coeffs(fy, c);
                s_sy[0][d] = (signed char)sy;
#pragma unroll
                for (int k = 0; k < 4; ++k) s_cy[0][d][k] = c[k];
            }
            __syncthreads();
            for (int e = tid; e < ph * dw; e += blockDim.x) {            // horizontal bicubic pass
                const int r = e / dw, ox = e - r * dw;
                const int sx = s_sx[0][ox];
                float t[4], c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    t[k] = val[(y0 + r) * w + x0 + clampi(sx - 1 + k, 0, pw - 1)];
                    c[k] = s_cx[0][ox][k];
                }
                s_h[0][r][ox] = dot4(t, c);
            }
            __syncthreads();
            for (int d = tid; d < dw * dh; d += blockDim.x) {            // vertical bicubic pass -> gu
                const int oy = d / dw, ox = d - oy * dw;
                const int sy = s_sy[0][oy];
                float t[4], c[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    t[r] = s_h[0][clampi(sy - 1 + r, 0, ph - 1)][ox];
                    c[r] = s_cy[0][oy][r];
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
            // tap tables per destination column / row
            for (int d = lane; d < dw; d += 32) {
                int sx;
                float fx, c[4];
                src_pos(d, scale, sx, fx);
                cubic_coeffs(fx, c);
                s_sx[warp][d] = (signed char)sx;
#pragma unroll
                for (int k = 0; k < 4; ++k) s_cx[warp][d][k] = c[k];
            }
            for (int d = lane; d < dh; d += 32) {
                int sy;
                float fy, c[4];
                src_pos(d, scale, sy, fy);
                cubic_coeffs(fy, c);
                s_sy[warp][d] = (signed char)sy;
#pragma unroll
                for (int k = 0; k < 4; ++k) s_cy[warp][d][k] = c[k];
            }
            __syncwarp();
            // horizontal pass: every patch row at every destination column
            for (int e = lane; e < ph * dw; e += 32) {
                const int r = e / dw, ox = e - r * dw;
                const int sx = s_sx[warp][ox];
                float t[4], c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    t[k] = val[(y0 + r) * w + x0 + clampi(sx - 1 + k, 0, pw - 1)];
                    c[k] = s_cx[warp][ox][k];
                }
                s_h[warp][r][ox] = dot4(t, c);
            }
            __syncwarp();
            // vertical pass + running argmax
            for (int d = lane; d < dw * dh; d += 32) {
                const int oy = d / dw, ox = d - oy * dw;
                const int sy = s_sy[warp][oy];
                float t[4], c[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    t[r] = s_h[warp][clampi(sy - 1 + r, 0, ph - 1)][ox];
                    c[r] = s_cy[warp][oy][r];
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

__global__ void __launch_bounds__(kLimbThreads)
limbs_kernel(const float *__restrict__ paf, int h, int w, int H, int W, double thre2, int direct,
             const float *__restrict__ peaks, const int *__restrict__ peak_counts, double *__restrict__ limbs,
             int *__restrict__ counts) {
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int ja = c_limb_src[t], jb = c_limb_dst[t];
    const int *pc = peak_counts + b * 16;
    const int na = pc[ja], nb = pc[jb];
    __shared__ double s_val[kPairChunk][kSamples];
    __shared__ double s_mean[kPairChunk];
    __shared__ unsigned char s_ok[kPairChunk];
    __shared__ double s_best[LWOP_MAX_PEAKS];
    __shared__ int s_bestj[LWOP_MAX_PEAKS];
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
    for (int c = 0; c < ja; ++c) ida += pc[c];
    for (int c = 0; c < jb; ++c) idb += pc[c];
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
        for (int item = tid; item < npairs * kSamples; item += kLimbThreads) {
            const int pi = item / kSamples, k = item - pi * kSamples;
            const int gp = chunk0 + pi;
            const int i = gp / nb, j = gp - i * nb;
            const double ax = pa[i * 4 + 0], ay = pa[i * 4 + 1];
            const double bx = pb[j * 4 + 0], by = pb[j * 4 + 1];
            const double dx = bx - ax, dy = by - ay;              // limb_dir (:249)
            const double norm = __dadd_rn(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), 1e-8);   // (:255)
            const double ux = dx / norm, uy = dy / norm;
            const double stepx = dx / 9.0, stepy = dy / 9.0;      // np.linspace(start, stop, 10)
            const double fxk = k == kSamples - 1 ? bx : __dadd_rn(__dmul_rn((double)k, stepx), ax);
            const double fyk = k == kSamples - 1 ? by : __dadd_rn(__dmul_rn((double)k, stepy), ay);
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
            s_val[pi][k] = sk;
        }
        __syncthreads();
        // ---- one thread per pair: the acceptance criteria (:271-283) ----
        for (int pi = tid; pi < npairs; pi += kLimbThreads) {
            const int gp = chunk0 + pi;
            const int i = gp / nb, j = gp - i * nb;
            const double dx = (double)pb[j * 4 + 0] - (double)pa[i * 4 + 0], dy = (double)pb[j * 4 + 1] - (double)pa[i * 4 + 1];
            const double *sv = s_val[pi];
            int above = 0;
#pragma unroll
            for (int k = 0; k < kSamples; ++k) above += sv[k] > thre2 ? 1 : 0;
            // numpy pairwise mean of 10 float64 values (:271)
            double sum = __dadd_rn(__dadd_rn(__dadd_rn(sv[0], sv[1]), __dadd_rn(sv[2], sv[3])),
                                   __dadd_rn(__dadd_rn(sv[4], sv[5]), __dadd_rn(sv[6], sv[7])));
            sum = __dadd_rn(__dadd_rn(sum, sv[8]), sv[9]);
            const double mean = sum / 10.0;
            const bool skip = fabs(dx) > (double)W || fabs(dy) > (double)H;       // (:209,250)
            s_mean[pi] = mean;
            s_ok[pi] = (!skip && above > 8 && mean > thre2) ? 1 : 0;
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
// ---------------------------------------------------------------------------
constexpr int kSmemPersons = 128;   // working persons kept in shared memory; the rest spill to the global scratch

__global__ void __launch_bounds__(32)
group_kernel(const float *__restrict__ peaks, const int *__restrict__ peak_counts, const double *__restrict__ limbs,
             int *__restrict__ counts, double *__restrict__ joints, double *__restrict__ persons,
             float *__restrict__ keypoints, double *__restrict__ work) {
    const int b = blockIdx.x, lane = threadIdx.x;
    // working list of persons before the final filter: at most one new person per connection
    __shared__ double Ps[kSmemPersons][16];
    __shared__ float s_score[LWOP_MAX_JOINTS];           // joint scores (float32 values, exact in float)
    __shared__ double s_cs[LWOP_MAX_PEAKS];              // connection scores of the current limb type
    __shared__ short s_ca[LWOP_MAX_PEAKS], s_cb[LWOP_MAX_PEAKS];   // their (src_id, dst_id)
    // columns ja / jb of every working person for the current limb type, as integers: the membership test of the
    // serial loop (:329-331) reads these instead of two doubles per person; kept in step with the rows below
    __shared__ short s_pa[kMaxWorkPersons], s_pb[kMaxWorkPersons];
    __shared__ int s_off[LWOP_NUM_JOINTS + 1];
    __shared__ float s_x[LWOP_MAX_JOINTS], s_y[LWOP_MAX_JOINTS];   // joint coordinates for the keypoint table
    __shared__ unsigned char s_type[LWOP_MAX_JOINTS];
    double *Pg = work + (long long)b * kMaxWorkPersons * 16;
    // explicit address spaces (a generic pointer would turn the shared-memory accesses into generic loads)
    auto row_ld = [&](int k, int c) -> double { return k < kSmemPersons ? Ps[k][c] : Pg[(long long)(k - kSmemPersons) * 16 + c]; };
    auto row_st = [&](int k, int c, double v) {
        if (k < kSmemPersons) Ps[k][c] = v;
        else Pg[(long long)(k - kSmemPersons) * 16 + c] = v;
    };

    const int *pc = peak_counts + b * 16;
    int *cnt = counts + b * LWOP_COUNTS;
    const int my_pc = lane < LWOP_NUM_JOINTS ? pc[lane] : 0;
    const int my_nconn = lane < LWOP_NUM_LIMBS ? cnt[18 + lane] : 0;
    {   // exclusive prefix sum of the per-type peak counts
        int incl = my_pc;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane < LWOP_NUM_JOINTS) {
            s_off[lane] = incl - my_pc;
            cnt[4 + lane] = my_pc;
        }
        if (lane == LWOP_NUM_JOINTS - 1) {
            s_off[LWOP_NUM_JOINTS] = incl;
            cnt[0] = incl;
        }
    }
    __syncwarp();
    double *jl = joints + (long long)b * LWOP_MAX_JOINTS * 6;
    for (int c = 0; c < LWOP_NUM_JOINTS; ++c) {                   // (x, y, score, id, 0, type) (:481-482)
        const float *pk = peaks + ((long long)b * LWOP_NUM_JOINTS + c) * LWOP_MAX_PEAKS * 4;
        const int n = __shfl_sync(0xffffffffu, my_pc, c);
        for (int i = lane; i < n; i += 32) {
            const int id = s_off[c] + i;
            const float4 pv = *reinterpret_cast<const float4 *>(pk + i * 4);
            double *o = jl + (long long)id * 6;
            o[0] = pv.x;
            o[1] = pv.y;
            o[2] = pv.z;
            o[3] = (double)id;
            o[4] = 0.0;
            o[5] = (double)c;
            s_score[id] = pv.z;
            s_x[id] = pv.x;
            s_y[id] = pv.y;
            s_type[id] = (unsigned char)c;
        }
    }
    __syncwarp();

    int np = 0, total_conn = 0;
    bool overflow = false;
    for (int t = 0; t < LWOP_NUM_LIMBS; ++t) {
        const int ja = c_limb_src[t], jb = c_limb_dst[t];
        const int nconn = __shfl_sync(0xffffffffu, my_nconn, t);
        total_conn += nconn;
        const double *rows = limbs + ((long long)b * LWOP_NUM_LIMBS + t) * LWOP_MAX_PEAKS * 5;
        for (int e = lane; e < nconn; e += 32) {
            s_ca[e] = (short)rows[e * 5 + 0];
            s_cb[e] = (short)rows[e * 5 + 1];
            s_cs[e] = rows[e * 5 + 2];
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
            const double js_b = (double)s_score[b_id];
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
                        const double js_a = (double)s_score[a_id];
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
                    const int jid = (int)ids[e];
                    const int jt = s_type[jid];
                    ko[kept * 42 + jt * 3 + 0] = s_x[jid];
                    ko[kept * 42 + jt * 3 + 1] = s_y[jid];
                    ko[kept * 42 + jt * 3 + 2] = 1.0f;
                }
            }
        }
        __syncwarp();
        ++kept;
    }
    if (lane == 0) {
        cnt[1] = kept;
        cnt[3] = total_conn;
        if (overflow) atomicOr(&cnt[2], 2);
    }
}

// K4: exclusive prefix over images of (joints, persons, connections), one block
__global__ void __launch_bounds__(1024)
pack_offsets_kernel(const int *__restrict__ counts, int B, int *__restrict__ offsets) {
    __shared__ int s[3][1024];
    const int tid = threadIdx.x;
    int carry[3] = {0, 0, 0};
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

}  // namespace

size_t decode_scratch_work_bytes(int B) { return (size_t)B * kMaxWorkPersons * 16 * sizeof(double); }

size_t decode_scratch_peaks_bytes(int B) {
    return (size_t)B * LWOP_NUM_JOINTS * LWOP_MAX_PEAKS * 4 * sizeof(float);
}

cudaError_t launch_decode_pack(const lwop_decode_tables &t, int B, int *offsets, double *joints_packed,
                               double *limbs_packed, double *persons_packed, float *keypoints_packed,
                               cudaStream_t s, long long *launches, long long cap_j, long long cap_l, long long cap_p) {
    pack_offsets_kernel<<<1, 1024, 0, s>>>(t.counts, B, offsets);
    pack_rows_kernel<<<B, 128, 0, s>>>(t, offsets, joints_packed, limbs_packed, persons_packed, keypoints_packed, cap_j,
                                       cap_l, cap_p);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeArgs &a, cudaStream_t s, long long *launches) {
    const int npix = a.h * a.w;
    if (npix > kMaxPixPerThread * 256) return cudaErrorInvalidValue;
    const size_t smem = (size_t)npix * 8 + 16;
    if (smem > 160 * 1024) return cudaErrorInvalidValue;
    if (smem > 16 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
    }
    const int ncounts = a.B * LWOP_COUNTS;
    zero_counts_kernel<<<(ncounts + 255) / 256, 256, 0, s>>>(a.t.counts, ncounts);
    const double factor = (double)a.H / (double)a.h;              // scale = img_H / heat_H (:462)
    peaks_kernel<<<dim3(LWOP_NUM_JOINTS, a.B), 256, smem, s>>>(a.heat, a.h, a.w, LWOP_NUM_JOINTS, a.thre1, factor, 0,
                                                               a.peaks, a.peak_counts, a.t.counts);
    limbs_kernel<<<dim3(LWOP_NUM_LIMBS, a.B), kLimbThreads, 0, s>>>(a.paf, a.h, a.w, a.H, a.W, a.thre2, 0, a.peaks,
                                                           a.peak_counts, a.t.limbs, a.t.counts);
    group_kernel<<<a.B, 32, 0, s>>>(a.peaks, a.peak_counts, a.t.limbs, a.t.counts, a.t.joints, a.t.persons,
                                    a.t.keypoints, a.work);
    if (launches) *launches += 4;
    return cudaGetLastError();
}

// ---- separately callable stages (reference NMS :107, find_connected_joints :188, group_limbs_of_same_person :304) ----
cudaError_t launch_decode_peaks(const float *heat, int B, int h, int w, double factor, float thre1, int mode,
                                float *peaks, int *peak_counts, int *counts, cudaStream_t s) {
    const int npix = h * w;
    if (npix > kMaxPixPerThread * 256) return cudaErrorInvalidValue;
    size_t smem = (size_t)npix * 8 + 16;
    if (mode & LWOP_PEAKS_GAUSS) {
        // the smoothed patch lives in shared memory: the clipped 5x5 patch times `factor` must fit 48 x 48
        if ((int)rint(5.0 * factor) > kRefineMax) return cudaErrorInvalidValue;
        smem += 2 * (size_t)kRefineMax * kRefineMax * sizeof(float) + 16;
    }
    if (smem > 160 * 1024) return cudaErrorInvalidValue;
    if (smem > 16 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
    }
    const int ncounts = B * LWOP_COUNTS;
    zero_counts_kernel<<<(ncounts + 255) / 256, 256, 0, s>>>(counts, ncounts);
    peaks_kernel<<<dim3(LWOP_NUM_JOINTS, B), 256, smem, s>>>(heat, h, w, LWOP_NUM_JOINTS, thre1, factor, mode, peaks,
                                                             peak_counts, counts);
    return cudaGetLastError();
}

cudaError_t launch_decode_limbs(const float *paf, int upsampled, int B, int h, int w, int H, int W, double thre2,
                                const float *peaks, const int *peak_counts, double *limbs, int *counts, cudaStream_t s) {
    limbs_kernel<<<dim3(LWOP_NUM_LIMBS, B), kLimbThreads, 0, s>>>(paf, h, w, H, W, thre2, upsampled ? 1 : 0, peaks, peak_counts,
                                                         limbs, counts);
    return cudaGetLastError();
}

cudaError_t launch_decode_group(const float *peaks, const int *peak_counts, const double *limbs, int *counts,
                                double *joints, double *persons, float *keypoints, double *work, int B, cudaStream_t s) {
    group_kernel<<<B, 32, 0, s>>>(peaks, peak_counts, limbs, counts, joints, persons, keypoints, work);
    return cudaGetLastError();
}

}  // namespace lwop
