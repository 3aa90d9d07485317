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

// ---------------------------------------------------------------------------------------------
// cv2.findHomography(src, dst), method 0 -- one WARP per point set.
//
// The arithmetic is OpenCV's, element for element (normalised DLT -> 9x9 LtL -> cv::eigen's Jacobi
// with its max-pivot bookkeeping -> de-normalisation -> LMSolver schedule for n > 4), so every
// matrix entry sees the same fp64 operations in the same order as the scalar oracle (whose n = 6 results are
// bit-identical to cv2 4.13.0, scripts/check_lm_vs_cv2.py); the warp only
// spreads *independent* entries over its lanes: the 45 LtL entries, the <= 27 element pairs of a
// Jacobi rotation, the 17 pivot candidates (shuffle arg-max), the 36+8 normal-equation entries and
// the per-point residual/Jacobian rows.  Matrices live in shared memory (dynamic indices).
// ---------------------------------------------------------------------------------------------
struct HomogScratch {
    double A[81], V[81], W[9];            // Jacobi: matrix (destroyed), eigenvectors (rows), eigenvalues (sorted)
    double J[108], N[81];                 // LM: Jacobian 12x9, normal matrix J^T J
    double r[12], rd[12], v[9], d[9], q[9], td[9];   // residual, trial residual, J^T r, step, back-substitution / scratch
    int indR[9], indC[9], perm[9], skip[9];
};

__device__ __forceinline__ double cv_hypot(double a, double b) {
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

#define FUSG_ROT(v0, v1) do { const double a0_ = (v0), b0_ = (v1); (v0) = a0_ * c - b0_ * s; (v1) = a0_ * s + b0_ * c; } while (0)

// first index of the strict maximum over the warp (ties -> lowest position), all lanes get it
__device__ __forceinline__ void warp_argmax_first(double &val, int &pos) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, val, off);
        const int op = __shfl_xor_sync(0xffffffffu, pos, off);
        if (ov > val || (ov == val && op < pos)) { val = ov; pos = op; }
    }
}

// row/column maxima bookkeeping of JacobiImpl_ for index idx (A upper triangle, n x n)
__device__ __forceinline__ int jacobi_row_max(const double *A, int n, int k) {
    int m = k + 1;
    double mv = fabs(A[n * k + m]);
    for (int i = k + 2; i < n; ++i) { const double val = fabs(A[n * k + i]); if (mv < val) { mv = val; m = i; } }
    return m;
}
__device__ __forceinline__ int jacobi_col_max(const double *A, int n, int k) {
    int m = 0;
    double mv = fabs(A[k]);
    for (int i = 1; i < k; ++i) { const double val = fabs(A[n * i + k]); if (mv < val) { mv = val; m = i; } }
    return m;
}

// cv::eigen (Jacobi) on sc.A (n = 9); leaves sc.perm = row order after OpenCV's descending sort.
__device__ inline void jacobi_eig_warp(HomogScratch &sc, const int lane) {
    constexpr int n = 9;
    const double eps = DBL_EPSILON;
    double *A = sc.A, *V = sc.V, *W = sc.W;
    for (int e = lane; e < n * n; e += 32) V[e] = (e / n == e % n) ? 1.0 : 0.0;
    if (lane < n) {
        W[lane] = A[(n + 1) * lane];
        if (lane < n - 1) sc.indR[lane] = jacobi_row_max(A, n, lane);
        if (lane > 0) sc.indC[lane] = jacobi_col_max(A, n, lane);
    }
    __syncwarp();
    const int maxIters = n * n * 30;
    for (int iters = 0; iters < maxIters; ++iters) {
        // pivot: first strict maximum over [row candidates 0..n-2, column candidates 1..n-1]
        double cand = -1.0;
        int pos = lane;
        if (lane < n - 1) cand = fabs(A[n * lane + sc.indR[lane]]);
        else if (lane < 2 * n - 2) { const int i = lane - (n - 1) + 1; cand = fabs(A[n * sc.indC[i] + i]); }
        warp_argmax_first(cand, pos);
        int k, l;
        if (pos < n - 1) { k = pos; l = sc.indR[pos]; }
        else { l = pos - (n - 1) + 1; k = sc.indC[l]; }
        const double p = A[n * k + l];
        if (fabs(p) <= eps) break;
        const double y = (W[l] - W[k]) * 0.5;
        double t = fabs(y) + cv_hypot(p, y);
        double s = cv_hypot(p, t);
        const double c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        __syncwarp();
        if (lane == 0) { A[n * k + l] = 0; W[k] -= t; W[l] += t; }
        if (lane < n) {
            const int i = lane;
            if (i < k) FUSG_ROT(A[n * i + k], A[n * i + l]);
            else if (i > k && i < l) FUSG_ROT(A[n * k + i], A[n * i + l]);
            else if (i > l) FUSG_ROT(A[n * k + i], A[n * l + i]);
        } else if (lane >= 16 && lane < 16 + n) {
            const int i = lane - 16;
            FUSG_ROT(V[n * k + i], V[n * l + i]);
        }
        __syncwarp();
        if (lane == 0 && k < n - 1) sc.indR[k] = jacobi_row_max(A, n, k);
        if (lane == 1 && k > 0) sc.indC[k] = jacobi_col_max(A, n, k);
        if (lane == 2 && l < n - 1) sc.indR[l] = jacobi_row_max(A, n, l);
        if (lane == 3 && l > 0) sc.indC[l] = jacobi_col_max(A, n, l);
        __syncwarp();
    }
    __syncwarp();
    if (lane == 0) {          // OpenCV's descending selection sort, tracked as a row permutation
        for (int i = 0; i < n; ++i) sc.perm[i] = i;
        for (int k = 0; k < n - 1; ++k) {
            int m = k;
            for (int i = k + 1; i < n; ++i) if (W[m] < W[i]) m = i;
            if (k != m) {
                const double tw = W[m]; W[m] = W[k]; W[k] = tw;
                const int tp = sc.perm[m]; sc.perm[m] = sc.perm[k]; sc.perm[k] = tp;
            }
        }
    }
    __syncwarp();
}

// entry `idx` of the two DLT rows of one correspondence (fundam.cpp runKernel)
__device__ __forceinline__ double dlt_lx(int idx, double X, double Y, double x) {
    switch (idx) { case 0: return X; case 1: return Y; case 2: return 1; case 6: return -x * X; case 7: return -x * Y; case 8: return -x; default: return 0; }
}
__device__ __forceinline__ double dlt_ly(int idx, double X, double Y, double y) {
    switch (idx) { case 3: return X; case 4: return Y; case 5: return 1; case 6: return -y * X; case 7: return -y * Y; case 8: return -y; default: return 0; }
}

// cv::solve(A, b, x, DECOMP_EIG) for the symmetric 9x9 in sc.A (destroyed): Jacobi factors, then
// SVBkSb with OpenCV's eigenvalue cut |w_i| <= 2 eps sum(w).  b: 9 doubles in shared memory; result in sc.d.
// With `diag_only` the routine instead returns max_c |(A^-1)_cc| of cv::invert(A, DECOMP_EIG) (LMSolver's
// lambda re-initialisation), seeded with DBL_EPSILON like the caller does.
__device__ inline double eig_solve_warp(HomogScratch &sc, const int lane, const double *b, bool diag_only) {
    constexpr int n = 9;
    jacobi_eig_warp(sc, lane);
    double threshold = 0;
    for (int i = 0; i < n; ++i) threshold += sc.W[i];
    threshold *= DBL_EPSILON * 2;
    if (!diag_only) {
        if (lane < n) {
            const double wi = sc.W[lane];
            const double *row = sc.V + n * sc.perm[lane];
            const bool cut = fabs(wi) <= threshold;
            double acc = 0;
            for (int j = 0; j < n; ++j) acc += row[j] * b[j];
            acc *= 1 / wi;
            sc.q[lane] = acc;
            sc.skip[lane] = cut;
        }
        __syncwarp();
        if (lane < n) {
            double x = 0;
            for (int i = 0; i < n; ++i)
                if (!sc.skip[i]) x = x + sc.q[i] * sc.V[n * sc.perm[i] + lane];
            sc.d[lane] = x;
        }
        __syncwarp();
        return 0;
    }
    double mx = 0;
    if (lane < n) {
        double x = 0;
        for (int i = 0; i < n; ++i) {
            const double wi = sc.W[i];
            if (fabs(wi) <= threshold) continue;
            const double vic = sc.V[n * sc.perm[i] + lane];
            const double sv = vic * (1 / wi);
            x = x + sv * vic;
        }
        mx = fabs(x);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    return fmax(DBL_EPSILON, mx);
}

// residual (+ Jacobian rows) of one correspondence at parameters h -- fundam.cpp HomographyRefineCallback
// as compiled into opencv-python 4.13.0: NINE parameters (h[8] in the denominator, 2n x 9 Jacobian).
__device__ __forceinline__ void lm_point(float Mxf, float Myf, float mxf, float myf, const double *h, double *err, double *Jrows) {
    const double Mx = Mxf, My = Myf;
    double ww = h[6] * Mx + h[7] * My + h[8];
    ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
    const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
    const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
    err[0] = xi - mxf;
    err[1] = yi - myf;
    if (Jrows) {
        Jrows[0] = Mx * ww; Jrows[1] = My * ww; Jrows[2] = ww;
        Jrows[3] = Jrows[4] = Jrows[5] = 0.;
        Jrows[6] = -Mx * ww * xi; Jrows[7] = -My * ww * xi; Jrows[8] = -ww * xi;
        Jrows[9] = Jrows[10] = Jrows[11] = 0.;
        Jrows[12] = Mx * ww; Jrows[13] = My * ww; Jrows[14] = ww;
        Jrows[15] = -Mx * ww * yi; Jrows[16] = -My * ww * yi; Jrows[17] = -ww * yi;
    }
}

// cv::Mat::dot (CV_64F, 9 elements) as the AVX2/FMA dispatch of the 4.13.0 wheel evaluates it:
// per group of four t = fma(a0,b0, a1*b1); t = fma(a2,b2,t); t = fma(a3,b3,t); res += t; tail res = fma(a,b,res).
__device__ __forceinline__ double cv_dot9(const double *a, const double *b) {
    double res = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 4) {
        double t = __fma_rn(a[i], b[i], a[i + 1] * b[i + 1]);
        t = __fma_rn(a[i + 2], b[i + 2], t);
        t = __fma_rn(a[i + 3], b[i + 3], t);
        res += t;
    }
    return __fma_rn(a[8], b[8], res);
}

// cv::gemm row . vector (GEMMSingleMul, one-column result): 4 interleaved accumulators, tail into s0
__device__ __forceinline__ double gemm_rowdot(const double *a, int astride, const double *b, int n) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int k = 0;
    for (; k <= n - 4; k += 4) {
        s0 += a[k * astride] * b[k];
        s1 += a[(k + 1) * astride] * b[k + 1];
        s2 += a[(k + 2) * astride] * b[k + 2];
        s3 += a[(k + 3) * astride] * b[k + 3];
    }
    for (; k < n; ++k) s0 += a[k * astride] * b[k];
    return ((s0 + s1) + s2) + s3;
}

// LM refinement (calib3d/levmarq.cpp LMSolverImpl::run, maxIters 10, epsx = epsf = FLT_EPSILON) of all nine
// entries of H; h9 uniform across the warp.  Mf/mf: this lane's correspondence (lane < count).
__device__ inline void lm_refine_warp(HomogScratch &sc, const int lane, const int count, float Mxf, float Myf, float mxf, float myf, double *h9) {
    constexpr int lx = 9;
    const int rows = 2 * count;
    const double epsx = FLT_EPSILON, epsf = FLT_EPSILON;
    double x[lx], xd[lx], D[lx];
#pragma unroll
    for (int i = 0; i < lx; ++i) x[i] = h9[i];
    auto residual = [&](const double *h, double *rdst, bool jac) {
        __syncwarp();
        if (lane < count) lm_point(Mxf, Myf, mxf, myf, h, rdst + 2 * lane, jac ? sc.J + 2 * lx * lane : nullptr);
        __syncwarp();
    };
    auto sumsq = [&](const double *rv) { double S = 0; for (int i = 0; i < rows; ++i) S += rv[i] * rv[i]; return S; };
    auto normal_eq = [&]() {
        // N = J^T J (cv::mulTransposed, sequential over rows), v = J^T r (cv::gemm) -- one entry per lane
        for (int e = lane; e < 45 + lx; e += 32) {
            if (e < 45) {
                int i = 0, rem = e;
                while (rem >= lx - i) { rem -= lx - i; ++i; }
                const int j = i + rem;
                double acc = 0;
                for (int k = 0; k < rows; ++k) acc += sc.J[k * lx + i] * sc.J[k * lx + j];
                sc.N[i * lx + j] = acc; sc.N[j * lx + i] = acc;
            } else {
                const int i = e - 45;
                sc.v[i] = gemm_rowdot(sc.J + i, lx, sc.r, rows);
            }
        }
        __syncwarp();
    };
    residual(x, sc.r, true);
    double S = sumsq(sc.r);
    normal_eq();
#pragma unroll
    for (int i = 0; i < lx; ++i) D[i] = sc.N[i * lx + i];
    const double Rlo = 0.25, Rhi = 0.75;
    double lambda = 1, lc = 0.75;
    int iter = 0;
    for (;;) {
        __syncwarp();
        for (int e = lane; e < lx * lx; e += 32) {
            const int ri = e / lx, ci = e - ri * lx;
            double a = sc.N[e];
            if (ri == ci) a += lambda * D[ri];
            sc.A[e] = a;
        }
        __syncwarp();
        eig_solve_warp(sc, lane, sc.v, false);               // step in sc.d
#pragma unroll
        for (int i = 0; i < lx; ++i) xd[i] = x[i] - sc.d[i];
        residual(xd, sc.rd, false);
        const double Sd = sumsq(sc.rd);
        // temp_d = -N d + 2 v  (cv::gemm(A, d, -1, v, 2))
        if (lane < lx) sc.td[lane] = -1. * gemm_rowdot(sc.N + lane * lx, 1, sc.d, lx) + 2. * sc.v[lane];
        __syncwarp();
        const double dS = cv_dot9(sc.d, sc.td);
        const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
        double nd = 0;
#pragma unroll
        for (int i = 0; i < lx; ++i) nd = fmax(nd, fabs(sc.d[i]));
        if (R > Rhi) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0;
        } else if (R < Rlo) {
            const double t = cv_dot9(sc.d, sc.v);
            double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
            nu = fmin(fmax(nu, 2.), 10.);
            if (lambda == 0) {
                __syncwarp();
                for (int e = lane; e < lx * lx; e += 32) sc.A[e] = sc.N[e];
                __syncwarp();
                const double maxval = eig_solve_warp(sc, lane, nullptr, true);
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        if (Sd < S) {
            S = Sd;
#pragma unroll
            for (int i = 0; i < lx; ++i) x[i] = xd[i];
            residual(x, sc.r, true);
            normal_eq();
        }
        iter++;
        double nr = 0;
        for (int i = 0; i < rows; ++i) nr = fmax(nr, fabs(sc.r[i]));
        if (!(iter < 10 && nd >= epsx && nr >= epsf)) break;
    }
#pragma unroll
    for (int i = 0; i < lx; ++i) h9[i] = x[i];
}

// Whole warp: s/d point to the 2*count int coordinates (uniform pointers).  Returns false where
// OpenCV returns None; H (9 doubles, identical on every lane) otherwise.
__device__ inline bool find_homography_warp(HomogScratch &sc, const int lane, const int *s, const int *d, const int count, double *H) {
    // centroids / mean-absolute-deviation scales: uniform, sequential over points like runKernel
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
    for (int i = 0; i < count; ++i) {
        cmx += (float)d[2 * i]; cmy += (float)d[2 * i + 1];
        cMx += (float)s[2 * i]; cMy += (float)s[2 * i + 1];
    }
    cmx /= count; cmy /= count; cMx /= count; cMy /= count;
    for (int i = 0; i < count; ++i) {
        smx += fabs((float)d[2 * i] - cmx); smy += fabs((float)d[2 * i + 1] - cmy);
        sMx += fabs((float)s[2 * i] - cMx); sMy += fabs((float)s[2 * i + 1] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) return false;
    smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
    // LtL: 45 upper-triangle entries over the lanes, each summed over the points in order
    __syncwarp();
    for (int e = lane; e < 45; e += 32) {
        int j = 0, rem = e;
        while (rem >= 9 - j) { rem -= 9 - j; ++j; }
        const int k = j + rem;
        double acc = 0;
        for (int i = 0; i < count; ++i) {
            const double x = ((float)d[2 * i] - cmx) * smx, y = ((float)d[2 * i + 1] - cmy) * smy;
            const double X = ((float)s[2 * i] - cMx) * sMx, Y = ((float)s[2 * i + 1] - cMy) * sMy;
            acc += dlt_lx(j, X, Y, x) * dlt_lx(k, X, Y, x) + dlt_ly(j, X, Y, y) * dlt_ly(k, X, Y, y);
        }
        sc.A[j * 9 + k] = acc;
        sc.A[k * 9 + j] = acc;
    }
    __syncwarp();
    jacobi_eig_warp(sc, lane);
    const double *H0 = sc.V + 9 * sc.perm[8];           // eigenvector of the smallest eigenvalue
    const double invHnorm[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
    const double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double Ht[9], H1[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc += invHnorm[i * 3 + k] * H0[k * 3 + j];
            Ht[i * 3 + j] = acc;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc += Ht[i * 3 + k] * Hnorm2[k * 3 + j];
            H1[i * 3 + j] = acc;
        }
    const double scl = 1. / H1[8];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = H1[i] * scl;
    if (count > 4) {
        const int li = lane < count ? lane : 0;
        lm_refine_warp(sc, lane, count, (float)s[2 * li], (float)s[2 * li + 1], (float)d[2 * li], (float)d[2 * li + 1], H);
        // H.convertTo(H, H.type(), scaleFor(H(2,2)))
        const double sc2 = fabs(H[8]) > DBL_EPSILON ? 1. / H[8] : 1.;
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] = H[i] * sc2;
    }
    return true;
}

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
