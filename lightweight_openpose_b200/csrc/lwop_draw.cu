// Canvas rasteriser: the drawing half of plot_pose (reference src/pose_decode.py:431-436) on the device.
//
// Per person (in table order) and limb type (in order) whose two joints are both set the reference issues
//     cv2.circle(canvas, joint0, 3, (255,255,255), -1); cv2.circle(canvas, joint1, 3, ...);
//     cv2.line(canvas, joint0, joint1, colors[person % 16], thickness=2)
// on a copy of the frame -- a painter's algorithm: the last primitive covering a pixel wins.  Here every (person,
// limb) unit is rasterised by one warp into a per-pixel PRIORITY map with atomicMax (priority = 2 * (person * 13 +
// limb) + {0: the two circles, 1: the line}; the priority alone determines the colour), and a second pass resolves
// frame + priority map into the canvas and resets the map.  Which pixels a primitive covers is OpenCV's arithmetic
// (modules/imgproc/src/drawing.cpp, restated in oracle/draw_oracle.py and pinned there against the installed cv2):
//   * filled circle of radius 3 / 1: fixed stencils, clipped to the image;
//   * line of thickness 2, LINE_8: the segment is first clipped to the image grown by 2 pixels per side (clipLine);
//     corners p +- (round(dy r), round(dx r)) in 16.16 fixed point, r = 65536 / |p1 - p0| in double; FillConvexPoly of
//     the four corners = outline by the fixed-point DDA Line2 (after clipLine in 16.16) + scan-line fill with the
//     edge stepping dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y)); then a radius-1 disc at both end points.
// All integer / IEEE-double arithmetic, no FMA contraction where the reference has separate operations.
#include "lwop_kernels.cuh"

namespace lwop {
namespace {

constexpr int kXYShift = 16;
constexpr long long kXYOne = 1ll << kXYShift;
constexpr long long kHalf = kXYOne >> 1;

__constant__ unsigned char c_palette[16][3] = {   // reference pose_decode.py:26-31
    {255, 0, 0},     {0, 255, 0},     {0, 0, 255},    {255, 0, 255},   {0, 255, 255},  {255, 0, 255},
    {100, 60, 39},   {35, 200, 111},  {200, 40, 145}, {238, 248, 220}, {255, 165, 100}, {20, 205, 50},
    {0, 191, 255},   {0, 0, 205},     {220, 20, 60},  {100, 20, 0}};
__constant__ signed char c_limb[13][2] = {{0, 1}, {1, 2}, {3, 4}, {4, 5}, {6, 7}, {7, 8}, {9, 10}, {10, 11},
                                          {12, 13}, {13, 0}, {13, 3}, {13, 6}, {13, 9}};   // reference :17-18
// half widths of the rows dy = -3..3 of cv2's filled radius-3 circle (29 pixels)
__constant__ signed char c_disc3_half[7] = {0, 2, 2, 3, 2, 2, 0};

struct Plot {
    int *prio;
    int W, H, value;
    __device__ __forceinline__ void operator()(long long x, long long y) const {
        if (x >= 0 && x < W && y >= 0 && y < H) atomicMax(prio + (size_t)y * W + x, value);
    }
};

// cv::clipLine(Size2l(w, h), pt1, pt2) on 64-bit points; false when the segment misses the rectangle
__device__ bool clip_line(long long w, long long h, long long &x1, long long &y1, long long &x2, long long &y2) {
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)(__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1)));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)(__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1)));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)(__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1)));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)(__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1)));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Line2(): fixed-point DDA between two 16.16 points; the lanes of the warp take the DDA steps k = lane, lane + 32, ...
// (point k is x1 + k * x_step, y1 + k * y_step: closed form, so the steps are independent)
__device__ void line2(const Plot &plot, long long x1, long long y1, long long x2, long long y2, int lane) {
    if (!clip_line((long long)plot.W << kXYShift, (long long)plot.H << kXYShift, x1, y1, x2, y2)) return;
    long long dx = x2 - x1, dy = y2 - y1;
    const long long ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    long long x_step, y_step;
    int ecount;
    if (ax > ay) {
        if (dx < 0) {
            dy = -dy;
            long long t = x1; x1 = x2; x2 = t;
            t = y1; y1 = y2; y2 = t;
        }
        x_step = kXYOne;
        y_step = (dy << kXYShift) / (ax | 1);        // C++ division: truncation towards zero, as in the reference
        ecount = (int)((x2 - x1) >> kXYShift);
    } else {
        if (dy < 0) {
            dx = -dx;
            long long t = x1; x1 = x2; x2 = t;
            t = y1; y1 = y2; y2 = t;
        }
        x_step = (dx << kXYShift) / (ay | 1);
        y_step = kXYOne;
        ecount = (int)((y2 - y1) >> kXYShift);
    }
    x1 += kHalf;
    y1 += kHalf;
    if (lane == 0) plot((x2 + kHalf) >> kXYShift, (y2 + kHalf) >> kXYShift);   // the end point is put first, on its own
    if (ax > ay) {
        const long long xi = x1 >> kXYShift;         // the major coordinate is truncated once, then incremented
        for (int k = lane; k <= ecount; k += 32) plot(xi + k, (y1 + k * y_step) >> kXYShift);
    } else {
        const long long yi = y1 >> kXYShift;
        for (int k = lane; k <= ecount; k += 32) plot((x1 + k * x_step) >> kXYShift, yi + k);
    }
}

struct Edge {
    int idx, di, ye;
    long long x, dx;
};

// FillConvexPoly(img, v, 4, colour, LINE_8, shift = XY_SHIFT).  The outline edges are independent (lanes split the
// DDA steps); the scan-line fill keeps the reference's sequential edge state, computed redundantly by every lane,
// and the lanes share the pixels of each span.
__device__ void fill_convex_quad(const Plot &plot, const long long (&vx)[4], const long long (&vy)[4], int lane) {
    constexpr int npts = 4;
    const int W = plot.W, H = plot.H;
    long long xmin = vx[0], xmax = vx[0], ymin = vy[0], ymax = vy[0];
    int imin = 0;
    long long px = vx[npts - 1], py = vy[npts - 1];
#pragma unroll
    for (int i = 0; i < npts; ++i) {
        if (vy[i] < ymin) { ymin = vy[i]; imin = i; }
        ymax = vy[i] > ymax ? vy[i] : ymax;
        xmax = vx[i] > xmax ? vx[i] : xmax;
        xmin = vx[i] < xmin ? vx[i] : xmin;
        line2(plot, px, py, vx[i], vy[i], lane);
        px = vx[i]; py = vy[i];
    }
    xmin = (xmin + kHalf) >> kXYShift;
    xmax = (xmax + kHalf) >> kXYShift;
    ymin = (ymin + kHalf) >> kXYShift;
    ymax = (ymax + kHalf) >> kXYShift;
    if (xmax < 0 || ymax < 0 || xmin >= W || ymin >= H) return;
    if (ymax > H - 1) ymax = H - 1;
    Edge edge[2];
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = (int)ymin;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -kXYOne;
    edge[0].dx = edge[1].dx = 0;
    int edges = npts;
    int y = (int)ymin;
    do {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx;
                const int di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    const int ty = (int)((vy[idx & 3] + kHalf) >> kXYShift);
                    if (ty > y) {
                        const long long xs = vx[idx0 & 3], xe = vx[idx & 3];
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (long long)(ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            const int left = edge[0].x > edge[1].x ? 1 : 0;
            long long xx1 = (edge[left].x + kHalf) >> kXYShift;
            long long xx2 = (edge[left ^ 1].x + kHalf) >> kXYShift;
            if (xx2 >= 0 && xx1 < W) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= W) xx2 = W - 1;
                for (long long x = xx1 + lane; x <= xx2; x += 32) plot(x, y);
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

// cv2.line(img, p0, p1, colour, thickness = 2) for integer end points
__device__ void thick_line2(const Plot &plot, long long px0, long long py0, long long px1, long long py1, int lane) {
    // cv::line clips the segment to the image grown by the thickness on every side first
    long long a0 = px0 + 2, b0 = py0 + 2, a1 = px1 + 2, b1 = py1 + 2;
    if (!clip_line((long long)plot.W + 4, (long long)plot.H + 4, a0, b0, a1, b1)) return;
    const long long x0 = (a0 - 2) << kXYShift, y0 = (b0 - 2) << kXYShift;
    const long long x1 = (a1 - 2) << kXYShift, y1 = (b1 - 2) << kXYShift;
    const double inv = 1.0 / (double)kXYOne;
    const double dx = __dmul_rn((double)(x0 - x1), inv), dy = __dmul_rn((double)(y1 - y0), inv);
    double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (fabs(r) > 2.220446049250313e-16) {            // DBL_EPSILON
        r = __ddiv_rn((double)(2ll << (kXYShift - 1)), __dsqrt_rn(r));
        const long long dpx = __double2ll_rn(__dmul_rn(dy, r)), dpy = __double2ll_rn(__dmul_rn(dx, r));   // cvRound
        const long long vx[4] = {x0 + dpx, x0 - dpx, x1 - dpx, x1 + dpx};
        const long long vy[4] = {y0 + dpy, y0 - dpy, y1 - dpy, y1 + dpy};
        fill_convex_quad(plot, vx, vy, lane);
    }
    if (lane < 10) {                                   // radius-1 end caps: centre + 4-neighbours
        const long long cx = ((lane < 5 ? x0 : x1) + kHalf) >> kXYShift, cy = ((lane < 5 ? y0 : y1) + kHalf) >> kXYShift;
        const int k = lane % 5;
        plot(cx + (k == 1 ? -1 : (k == 3 ? 1 : 0)), cy + (k == 0 ? -1 : (k == 4 ? 1 : 0)));
    }
}

__device__ __forceinline__ void disc3(const Plot &plot, long long cx, long long cy, int lane) {
    // 29 pixels: lane -> (row, column) of the stencil
    if (lane < 29) {
        int k = lane, row = 0;
        while (k >= 2 * c_disc3_half[row] + 1) { k -= 2 * c_disc3_half[row] + 1; ++row; }
        plot(cx + k - c_disc3_half[row], cy + row - 3);
    }
}

constexpr int kDrawWarps = 13;        // one warp per limb type
constexpr int kDrawPersonSlots = 8;   // CTAs per image; CTA s takes persons s, s + 8, ... (the person count is only known on the device)

__global__ void __launch_bounds__(32 * kDrawWarps)
draw_mark_kernel(const int *__restrict__ counts, const double *__restrict__ joints, const double *__restrict__ persons,
                 int *__restrict__ prio, int H, int W) {
    const int b = blockIdx.y;
    const int n_persons = min(counts[(size_t)b * LWOP_COUNTS + 1], LWOP_MAX_PERSONS);
    const int limb = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int person = blockIdx.x; person < n_persons; person += gridDim.x) {
        const double *row = persons + ((size_t)b * LWOP_MAX_PERSONS + person) * 16;
        // person_joint_info[joint_to_limb_heatmap_relationship[limb_type]].astype(int); -1 = joint not set (:409-415)
        const int ia = (int)row[c_limb[limb][0]], ib = (int)row[c_limb[limb][1]];
        if (ia < 0 || ib < 0 || ia >= LWOP_MAX_JOINTS || ib >= LWOP_MAX_JOINTS) continue;   // not set / not a row of this table
        const double *ja = joints + ((size_t)b * LWOP_MAX_JOINTS + ia) * 6, *jb = joints + ((size_t)b * LWOP_MAX_JOINTS + ib) * 6;
        const long long xa = (long long)ja[0], ya = (long long)ja[1];      // joint[0:2].astype(int): truncation
        const long long xb = (long long)jb[0], yb = (long long)jb[1];
        Plot plot{prio + (size_t)b * H * W, W, H, 2 * (person * LWOP_NUM_LIMBS + limb)};
        disc3(plot, xa, ya, lane);
        disc3(plot, xb, yb, lane);
        plot.value += 1;
        thick_line2(plot, xa, ya, xb, yb, lane);
    }
}

__device__ __forceinline__ void resolve_pixel(int p, unsigned char &c0, unsigned char &c1, unsigned char &c2) {
    if (p < 0) return;
    if (p & 1) {
        const int person = (p >> 1) / LWOP_NUM_LIMBS;
        c0 = c_palette[person & 15][0]; c1 = c_palette[person & 15][1]; c2 = c_palette[person & 15][2];
    } else {
        c0 = c1 = c2 = 255;
    }
}

// canvas = frame with the marked pixels painted (canvas may be the frame itself: every thread reads its pixels
// before it writes them); the priority map is reset to -1 for the next call.
// Thread = 4 consecutive pixels of the flat [B*H*W] range: one 128-bit priority load, three 32-bit frame words.
__global__ void __launch_bounds__(256)
draw_resolve4_kernel(const uint32_t *img, int4 *__restrict__ prio, uint32_t *canvas, size_t quads) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    const int4 p = prio[q];
    uint32_t w0 = img[3 * q], w1 = img[3 * q + 1], w2 = img[3 * q + 2];
    if ((p.x & p.y & p.z & p.w) >= 0) {                 // at least one non-negative priority
        unsigned char c[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[i] = (w0 >> (8 * i)) & 255; c[4 + i] = (w1 >> (8 * i)) & 255; c[8 + i] = (w2 >> (8 * i)) & 255; }
        resolve_pixel(p.x, c[0], c[1], c[2]);
        resolve_pixel(p.y, c[3], c[4], c[5]);
        resolve_pixel(p.z, c[6], c[7], c[8]);
        resolve_pixel(p.w, c[9], c[10], c[11]);
        w0 = c[0] | (c[1] << 8) | (c[2] << 16) | ((uint32_t)c[3] << 24);
        w1 = c[4] | (c[5] << 8) | (c[6] << 16) | ((uint32_t)c[7] << 24);
        w2 = c[8] | (c[9] << 8) | (c[10] << 16) | ((uint32_t)c[11] << 24);
        prio[q] = make_int4(-1, -1, -1, -1);
    }
    canvas[3 * q] = w0; canvas[3 * q + 1] = w1; canvas[3 * q + 2] = w2;
}

// scalar form: the tail of the flat range, or everything when a pointer is not 4-byte aligned
__global__ void __launch_bounds__(256)
draw_resolve1_kernel(const unsigned char *img, int *__restrict__ prio, unsigned char *canvas,
                     size_t first, size_t n) {
    const size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = prio[i];
    unsigned char c0 = img[3 * i], c1 = img[3 * i + 1], c2 = img[3 * i + 2];
    if (p >= 0) {
        resolve_pixel(p, c0, c1, c2);
        prio[i] = -1;
    }
    canvas[3 * i] = c0; canvas[3 * i + 1] = c1; canvas[3 * i + 2] = c2;
}

}  // namespace

cudaError_t launch_draw_poses(const uint8_t *img, int B, int H, int W, const int *counts, const double *joints,
                              const double *persons, int *prio, uint8_t *canvas, cudaStream_t s, long long *launches) {
    draw_mark_kernel<<<dim3(kDrawPersonSlots, B), 32 * kDrawWarps, 0, s>>>(counts, joints, persons, prio, H, W);
    ++*launches;
    const size_t n = (size_t)B * H * W;
    const bool aligned = (((uintptr_t)img | (uintptr_t)canvas) & 3) == 0;
    const size_t quads = aligned ? n / 4 : 0;
    if (quads) {
        draw_resolve4_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, s>>>((const uint32_t *)img, (int4 *)prio,
                                                                              (uint32_t *)canvas, quads);
        ++*launches;
    }
    if (quads * 4 < n) {
        const size_t rest = n - quads * 4;
        draw_resolve1_kernel<<<(unsigned)((rest + 255) / 256), 256, 0, s>>>(img, prio, canvas, quads * 4, n);
        ++*launches;
    }
    return cudaGetLastError();
}

}  // namespace lwop
