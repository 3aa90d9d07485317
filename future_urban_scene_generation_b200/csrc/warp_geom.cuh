// warp_geom.cuh -- device-side geometry of the fused planar-warp path (sm_100a).
//
// Everything here is integer or strictly-ordered fp64 arithmetic: this translation unit is
// compiled with -fmad=false so that no multiply-add is contracted and results are a function
// of the written operation order only (the CPU oracle is built with -ffp-contract=off and is
// compared bit for bit).
//
// Reference behaviour being reproduced (file:line relative to the reference repo):
//   cv2.fillPoly          warp_learn/planes_utils.py:29, warp_learn/online_visibility.py:84
//   cv2.findHomography    warp_learn/planes_utils.py:71-72
//   project_points        warp_learn/online_visibility.py:28-56
//   camera_planes_dist    warp_learn/online_visibility.py:59-75
#pragma once
#include <cfloat>
#include <cstdint>

namespace fusg {

constexpr int N_KP = 12;
constexpr int N_TEX = 5;
constexpr int N_VIS = 7;
constexpr int MAX_RANGES = 9;   // <=3 interior spans + <=6 outline runs per scanline

// plane -> keypoint index (utils/keypoint_utils.py:9-13 order), online_visibility.py:9-25,110-114
__device__ __constant__ int c_plane_n[N_VIS] = {6, 6, 4, 4, 4, 4, 4};
__device__ __constant__ int c_plane_kp[N_VIS][6] = {
    {0, 1, 3, 2, 9, 8},   {4, 5, 7, 6, 11, 10}, {8, 9, 11, 10, 0, 0}, {2, 6, 11, 9, 0, 0},
    {0, 4, 10, 8, 0, 0},  {2, 6, 7, 3, 0, 0},   {0, 4, 5, 1, 0, 0},
};

// ---------------------------------------------------------------------------------------------
// OpenCV's clipLine (imgproc/drawing.cpp): integer Cohen-Sutherland variant whose intersections are
// computed in fp64 and truncated.  Returns true when a visible segment remains; like the original
// it may leave the endpoints PARTIALLY clipped when it returns false, and CollectPolyEdges then
// uses them as they are.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool clip_line(long long width, long long height, long long &x1, long long &y1, long long &x2, long long &y2) {
    const long long right = width - 1, bottom = height - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Polygon vertices beyond this magnitude are refused (plane_j = -2): the 16.16 edge arithmetic below stays
// inside int64 for |coordinate| <= 2^20, and cv2 itself only takes int32 vertices.
constexpr int POLY_COORD_MAX = 1 << 20;

// ---------------------------------------------------------------------------------------------
// Scanline coverage of cv2.fillPoly (single contour, 8-connected, shift 0) on an H x W canvas: the
// pixels of row y (0 <= y < H) that are set are the union of the returned closed ranges
// [lo[k], hi[k]] intersected with [0, W).  Vertices may lie outside the canvas.
//   outline  = Line() of every edge: clipLine first, then an 8-connected Bresenham traced
//              left-to-right FROM THE CLIPPED END POINTS (closed form per row);
//   interior = even-odd spans [ceil(xa), floor(xb)] between 16.16 fixed-point edge crossings, where
//              an edge with an out-of-canvas end point takes its x (always) and y (unless the
//              clipped segment is horizontal) from the clipped end points and is extrapolated
//              back to its original top row (CollectPolyEdges / FillEdgeCollection).
// Pinned against cv2 4.13.0 on 26,000 random in- and out-of-frame polygons through the oracle
// restatement (tests/test_cpu_oracle.py), which this function follows operation for operation.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int poly_row_ranges(const int *px, const int *py, int n, int y, int H, int W, int *lo, int *hi) {
    int cnt = 0;
    long long xs[6];
    int na = 0;
    int ax = px[n - 1], ay = py[n - 1];
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const int bx = px[i], by = py[i];
        long long t0x = ax, t0y = ay, t1x = bx, t1y = by;
        const bool outside = (unsigned)ax >= (unsigned)W || (unsigned)bx >= (unsigned)W || (unsigned)ay >= (unsigned)H || (unsigned)by >= (unsigned)H;
        bool drawn = true;
        if (outside) drawn = clip_line(W, H, t0x, t0y, t1x, t1y);
        // ---- outline run of the (clipped) segment t0->t1 in row y
        if (drawn) {
            int sx = (int)t0x, sy = (int)t0y, ex = (int)t1x, ey = (int)t1y;
            if (ex < sx) { const int tx = sx, ty = sy; sx = ex; sy = ey; ex = tx; ey = ty; }
            const int ymin = sy < ey ? sy : ey, ymax = sy < ey ? ey : sy;
            if (y >= ymin && y <= ymax) {
                const int ddx = ex - sx;
                const int dys = ey - sy;
                const int ddy = dys < 0 ? -dys : dys;
                const int cy = y > sy ? y - sy : sy - y;
                int l, h;
                if (ddy > ddx) {                    // y-major: one pixel per row
                    l = h = sx + (2 * ddx * cy + ddy - 1) / (2 * ddy);
                } else if (ddy == 0) {              // horizontal (or a single point)
                    l = sx; h = ex;
                } else {                            // x-major: a run of pixels per row
                    const int il = cy == 0 ? 0 : (2 * ddx * cy - ddx + 1 + 2 * ddy - 1) / (2 * ddy);
                    const int ih = cy == ddy ? ddx : (2 * ddx * (cy + 1) - ddx + 1 + 2 * ddy - 1) / (2 * ddy) - 1;
                    l = sx + il; h = sx + ih;
                }
                lo[cnt] = l; hi[cnt] = h; ++cnt;
            }
        }
        // ---- interior crossing (half-open in y, 16.16 fixed point, C truncating division)
        if (ay != by) {
            const int yt = ay < by ? ay : by, yb = ay < by ? by : ay;
            if (y >= yt && y < yb) {
                long long c0x = (long long)ax * 65536LL, c0y = ay, c1x = (long long)bx * 65536LL, c1y = by;
                if (outside) {
                    if (t0y != t1y) { c0y = t0y; c1y = t1y; }
                    c0x = t0x * 65536LL; c1x = t1x * 65536LL;
                }
                const long long dxf = (c1x - c0x) / (c1y - c0y);
                const long long x0 = ay < by ? c0x + ((long long)ay - c0y) * dxf : c1x + ((long long)by - c1y) * dxf;
                xs[na++] = x0 + (long long)(y - yt) * dxf;
            }
        }
        ax = bx; ay = by;
    }
    // sort the (<=6) crossings, pair them even-odd
    for (int i = 1; i < na; ++i) {
        long long v = xs[i];
        int k = i - 1;
        while (k >= 0 && xs[k] > v) { xs[k + 1] = xs[k]; --k; }
        xs[k + 1] = v;
    }
    for (int k = 0; k + 1 < na; k += 2) {
        long long l = (xs[k] + 65535LL) >> 16, h = xs[k + 1] >> 16;
        if (l < W && h >= 0) {
            if (l < 0) l = 0;
            if (h >= W) h = W - 1;
            if (l <= h) { lo[cnt] = (int)l; hi[cnt] = (int)h; ++cnt; }
        }
    }
    return cnt;
}

// ---------------------------------------------------------------------------------------------
// The same coverage in two steps for the hot kernels, which evaluate MANY rows of one polygon: everything about an
// edge that does not depend on the row -- clipLine, the sorted end points of the outline run, the 16.16 slope of the
// interior crossing (a 64-bit division) -- is computed once per edge (poly_edge_setup, one lane per edge), and a row
// only pays for range tests, the Bresenham closed form (one or two 32-bit unsigned divisions: every numerator is
// non-negative) and one 64-bit multiply-add per crossing edge.  Integer operation for operation the same results as
// poly_row_ranges (tests/test_warp_gpu.py compares k_visibility / the gather, which use this form, against the oracle).
// ---------------------------------------------------------------------------------------------
struct PolyEdge {
    int sx, sy;          // left end point of the clipped segment (outline run)
    int ddx, dys;        // ex - sx >= 0, ey - sy
    int ymin, ymax;      // rows of the outline run (ymin > ymax: the segment is not drawn)
    int yt, yb;          // rows [yt, yb) where the edge is an interior crossing (yt == yb: horizontal edge)
    long long x0, dxf;   // 16.16 crossing at row yt, increment per row
};

// edge from vertex a to vertex b of a polygon on an H x W canvas
__device__ __forceinline__ void poly_edge_setup(int ax, int ay, int bx, int by, int H, int W, PolyEdge &e) {
    long long t0x = ax, t0y = ay, t1x = bx, t1y = by;
    const bool outside = (unsigned)ax >= (unsigned)W || (unsigned)bx >= (unsigned)W || (unsigned)ay >= (unsigned)H || (unsigned)by >= (unsigned)H;
    bool drawn = true;
    if (outside) drawn = clip_line(W, H, t0x, t0y, t1x, t1y);
    int sx = (int)t0x, sy = (int)t0y, ex = (int)t1x, ey = (int)t1y;
    if (ex < sx) { const int tx = sx, ty = sy; sx = ex; sy = ey; ex = tx; ey = ty; }
    e.sx = sx; e.sy = sy; e.ddx = ex - sx; e.dys = ey - sy;
    e.ymin = drawn ? (sy < ey ? sy : ey) : 1;
    e.ymax = drawn ? (sy < ey ? ey : sy) : 0;
    e.yt = e.yb = 0; e.x0 = e.dxf = 0;
    if (ay != by) {
        e.yt = ay < by ? ay : by; e.yb = ay < by ? by : ay;
        long long c0x = (long long)ax * 65536LL, c0y = ay, c1x = (long long)bx * 65536LL, c1y = by;
        if (outside) {
            if (t0y != t1y) { c0y = t0y; c1y = t1y; }
            c0x = t0x * 65536LL; c1x = t1x * 65536LL;
        }
        const long long dxf = (c1x - c0x) / (c1y - c0y);
        e.dxf = dxf;
        e.x0 = ay < by ? c0x + ((long long)ay - c0y) * dxf : c1x + ((long long)by - c1y) * dxf;
    }
}

// row y of a polygon whose n edges were prepared by poly_edge_setup (in polygon order): same result as poly_row_ranges
__device__ __forceinline__ int poly_row_ranges_edges(const PolyEdge *ed, int n, int y, int W, int *lo, int *hi) {
    int cnt = 0;
    long long xs[6];
    int na = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const PolyEdge &e = ed[i];
        if (y >= e.ymin && y <= e.ymax) {
            const unsigned ddx = (unsigned)e.ddx, ddy = (unsigned)(e.dys < 0 ? -e.dys : e.dys);
            const unsigned cy = (unsigned)(y > e.sy ? y - e.sy : e.sy - y);
            int l, h;
            if (ddy > ddx) {
                l = h = e.sx + (int)((2u * ddx * cy + ddy - 1u) / (2u * ddy));
            } else if (ddy == 0) {
                l = e.sx; h = e.sx + (int)ddx;
            } else {
                const unsigned il = cy == 0 ? 0u : (2u * ddx * cy - ddx + 2u * ddy) / (2u * ddy);
                const unsigned ih = cy == ddy ? ddx : (2u * ddx * (cy + 1u) - ddx + 2u * ddy) / (2u * ddy) - 1u;
                l = e.sx + (int)il; h = e.sx + (int)ih;
            }
            lo[cnt] = l; hi[cnt] = h; ++cnt;
        }
        if (y >= e.yt && y < e.yb) xs[na++] = e.x0 + (long long)(y - e.yt) * e.dxf;
    }
    for (int i = 1; i < na; ++i) {
        long long v = xs[i];
        int k = i - 1;
        while (k >= 0 && xs[k] > v) { xs[k + 1] = xs[k]; --k; }
        xs[k + 1] = v;
    }
    for (int k = 0; k + 1 < na; k += 2) {
        long long l = (xs[k] + 65535LL) >> 16, h = xs[k + 1] >> 16;
        if (l < W && h >= 0) {
            if (l < 0) l = 0;
            if (h >= W) h = W - 1;
            if (l <= h) { lo[cnt] = (int)l; hi[cnt] = (int)h; ++cnt; }
        }
    }
    return cnt;
}

// bits of 32-pixel word `w` (pixels 32w..32w+31, bit i = pixel 32w+i) covered by the ranges
__device__ __forceinline__ unsigned ranges_word(const int *lo, const int *hi, int cnt, int w) {
    unsigned bits = 0;
    const int base = w * 32;
    for (int k = 0; k < cnt; ++k) {
        int l = lo[k] - base, h = hi[k] - base;
        if (h < 0 || l > 31) continue;
        l = l < 0 ? 0 : l;
        h = h > 31 ? 31 : h;
        bits |= (0xffffffffu >> (31 - h)) & (0xffffffffu << l);
    }
    return bits;
}

// ---------------------------------------------------------------------------------------------
// project_points for one keypoint; P = K @ E[:3] and q = P @ [X,1] with k ascending, then /q[2].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void project_point(const double *K, const double *E, const double *X, double *u, double *v) {
    double q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double P[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = K[i * 3 + 0] * E[0 * 4 + j];
            s = s + K[i * 3 + 1] * E[1 * 4 + j];
            s = s + K[i * 3 + 2] * E[2 * 4 + j];
            P[j] = s;
        }
        double s = P[0] * X[0];
        s = s + P[1] * X[1];
        s = s + P[2] * X[2];
        s = s + P[3] * 1.0;
        q[i] = s;
    }
    *u = q[0] / q[2];
    *v = q[1] / q[2];
}

// camera_planes_dist for plane p: || (-R^T t) - mean(kp3d of plane) ||
__device__ __forceinline__ double plane_distance(const double *E, const double *kp3d, int p) {
    double c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double s = E[0 * 4 + i] * E[0 * 4 + 3];
        s = s + E[1 * 4 + i] * E[1 * 4 + 3];
        s = s + E[2 * 4 + i] * E[2 * 4 + 3];
        c[i] = -s;
    }
    const int n = c_plane_n[p];
    double mean[3] = {0, 0, 0};
    for (int k = 0; k < n; ++k)
        for (int a = 0; a < 3; ++a) mean[a] += kp3d[3 * c_plane_kp[p][k] + a];
    double s = 0;
    for (int a = 0; a < 3; ++a) {
        mean[a] /= n;
        const double d = c[a] - mean[a];
        s += d * d;
    }
    return sqrt(s);
}

// (cv2.findHomography lives in warp_solver.cuh)

// cv::invert of a 3x3 (closed form, OpenCV's operation order); zeros if singular.
__device__ __forceinline__ void invert3(const double *S, double *T) {
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.) { for (int i = 0; i < 9; ++i) T[i] = 0; return; }
    d = 1. / d;
    T[0] = (S[4] * S[8] - S[5] * S[7]) * d;
    T[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    T[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    T[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    T[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    T[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    T[6] = (S[3] * S[7] - S[4] * S[6]) * d;
    T[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    T[8] = (S[0] * S[4] - S[1] * S[3]) * d;
}

// gating + left/right symmetry remap of warp_unwarp_planes (planes_utils.py:57-68): target j or -1
__device__ __forceinline__ int plane_target(int i, const uint8_t *src_vis, const uint8_t *dst_vis) {
    if (!src_vis[i]) return -1;
    if (i >= 2 && !dst_vis[i]) return -1;
    if (i < 2 && !(dst_vis[0] == 1 || dst_vis[1] == 1)) return -1;
    int j = i;
    if (i < 2 && !dst_vis[i]) j = 1 - i;
    return j;
}

}  // namespace fusg
