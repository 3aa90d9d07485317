/*
 * oracle/warp_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
 *
 * Plain-C, single-threaded restatement of the warp half of the reference's
 * novel-view-completion path:
 *     warp_learn/online_visibility.py:28-150   (projection, plane distances, visibility)
 *     warp_learn/planes_utils.py:11-82         (get_planes, warp_unwarp_planes)
 * and of the three OpenCV routines those files call.  OpenCV is a third-party
 * dependency that is NOT vendored under /root/reference (requirements.txt:5,
 * "opencv-python", unpinned; the container has 4.13.0), so its published
 * algorithms are restated here and pinned by tests/test_oracle_vs_cv2.py against
 * the installed cv2 and by tests/golden/ vectors produced by the imported
 * reference (scripts/make_golden_warp.py):
 *     cv2.fillPoly          planes_utils.py:29, online_visibility.py:84
 *     cv2.findHomography    planes_utils.py:71-72   (method 0: normalised DLT, Jacobi
 *                                                   eigen-solver, LM refinement if n > 4)
 *     cv2.warpPerspective   planes_utils.py:76-77   (INTER_LINEAR, BORDER_CONSTANT 0)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off: no FMA contraction, so
 * that fp64 results are a function of the written operation order only).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

#define N_KP 12
#define N_TEX 5   /* left right roof front back */
#define N_VIS 7   /* + front_bt back_bt          */

/* ------------------------------------------------------------------------- */
/* plane -> keypoint tables (online_visibility.py:9-25,110-114), indices into */
/* _KP_NAMES order (utils/keypoint_utils.py:9-13)                             */
/*  0 left_back_trunk 1 left_back_wheel 2 left_front_light 3 left_front_wheel */
/*  4 right_back_trunk 5 right_back_wheel 6 right_front_light                 */
/*  7 right_front_wheel 8 upper_left_rearwindow 9 upper_left_windshield       */
/* 10 upper_right_rearwindow 11 upper_right_windshield                        */
/* ------------------------------------------------------------------------- */
static const int PLANE_N[N_VIS] = {6, 6, 4, 4, 4, 4, 4};
static const int PLANE_KP[N_VIS][6] = {
    {0, 1, 3, 2, 9, 8},     /* left  */
    {4, 5, 7, 6, 11, 10},   /* right */
    {8, 9, 11, 10, -1, -1}, /* roof  */
    {2, 6, 11, 9, -1, -1},  /* front */
    {0, 4, 10, 8, -1, -1},  /* back  */
    {2, 6, 7, 3, -1, -1},   /* front_bt */
    {0, 4, 5, 1, -1, -1},   /* back_bt  */
};

int orc_plane_table(int plane, int *idx) {
    for (int k = 0; k < PLANE_N[plane]; ++k) idx[k] = PLANE_KP[plane][k];
    return PLANE_N[plane];
}

/* ========================================================================= */
/* cv2.fillPoly (single contour, 8-connected, shift 0): outline U interior   */
/* mask is H*W bytes; filled pixels are set to `val`.                         */
/* ========================================================================= */

/* OpenCV clipLine (imgproc/drawing.cpp), integer Cohen-Sutherland variant. */
static int clip_line(int64_t width, int64_t height, int64_t *px1, int64_t *py1, int64_t *px2, int64_t *py2) {
    int64_t x1 = *px1, y1 = *py1, x2 = *px2, y2 = *py2;
    int64_t right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return 0;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    *px1 = x1; *py1 = y1; *px2 = x2; *py2 = y2;
    return (c1 | c2) == 0;
}

/* 8-connected Bresenham, always traced left-to-right (OpenCV LineIterator). */
static void draw_line8(uint8_t *mask, int H, int W, int64_t x0, int64_t y0, int64_t x1, int64_t y1, uint8_t val) {
    if (!clip_line(W, H, &x0, &y0, &x1, &y1)) return;
    int dx = (int)(x1 - x0), dy = (int)(y1 - y0);
    int px = (int)x0, py = (int)y0;
    if (dx < 0) { dx = -dx; dy = -dy; px = (int)x1; py = (int)y1; }
    int ystep = dy < 0 ? -1 : 1;
    if (dy < 0) dy = -dy;
    if (dy > dx) { /* y-major */
        int err = dy - 2 * dx, count = dy + 1;
        for (int i = 0; i < count; ++i) {
            mask[(size_t)py * W + px] = val;
            int m = err < 0 ? -1 : 0;
            err += -2 * dx + ((2 * dy) & m);
            py += ystep;
            px += 1 & m;
        }
    } else { /* x-major */
        int err = dx - 2 * dy, count = dx + 1;
        for (int i = 0; i < count; ++i) {
            mask[(size_t)py * W + px] = val;
            int m = err < 0 ? -1 : 0;
            err += -2 * dy + ((2 * dx) & m);
            px += 1;
            py += ystep & m;
        }
    }
}

typedef struct { int y0, y1; int64_t x, dx; } PolyEdge;

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

#define XY_SHIFT 16
#define XY_ONE (1 << XY_SHIFT)

void orc_fill_poly(uint8_t *mask, int H, int W, const int32_t *pts, int n, uint8_t val) {
    PolyEdge edges[16];
    int ne = 0;
    if (n > 16) return;
    int64_t p0x = (int64_t)pts[2 * (n - 1)] << XY_SHIFT, p0y = pts[2 * (n - 1) + 1];
    for (int i = 0; i < n; ++i) {
        int64_t p1x = (int64_t)pts[2 * i] << XY_SHIFT, p1y = pts[2 * i + 1];
        int64_t t0x = (p0x + (XY_ONE >> 1)) >> XY_SHIFT, t0y = p0y;
        int64_t t1x = (p1x + (XY_ONE >> 1)) >> XY_SHIFT, t1y = p1y;
        draw_line8(mask, H, W, t0x, t0y, t1x, t1y, val);
        int64_t c0x = p0x, c0y = p0y, c1x = p1x, c1y = p1y;
        if ((uint64_t)t0x >= (uint64_t)W || (uint64_t)t1x >= (uint64_t)W ||
            (uint64_t)t0y >= (uint64_t)H || (uint64_t)t1y >= (uint64_t)H) {
            clip_line(W, H, &t0x, &t0y, &t1x, &t1y);
            if (t0y != t1y) { c0y = t0y; c1y = t1y; }
            c0x = t0x << XY_SHIFT; c1x = t1x << XY_SHIFT;
        }
        if (p0y != p1y) {
            PolyEdge e;
            e.dx = (c1x - c0x) / (c1y - c0y);
            if (p0y < p1y) { e.y0 = (int)p0y; e.y1 = (int)p1y; e.x = c0x + (p0y - c0y) * e.dx; }
            else           { e.y0 = (int)p1y; e.y1 = (int)p0y; e.x = c1x + (p1y - c1y) * e.dx; }
            edges[ne++] = e;
        }
        p0x = p1x; p0y = p1y;
    }
    if (ne < 2) return;
    int y_min = INT_MAX, y_max = INT_MIN;
    for (int i = 0; i < ne; ++i) {
        if (edges[i].y0 < y_min) y_min = edges[i].y0;
        if (edges[i].y1 > y_max) y_max = edges[i].y1;
    }
    if (y_max > H) y_max = H;
    for (int y = y_min; y < y_max; ++y) {
        int64_t xs[16];
        int na = 0;
        for (int i = 0; i < ne; ++i)
            if (edges[i].y0 <= y && y < edges[i].y1)
                xs[na++] = edges[i].x + (int64_t)(y - edges[i].y0) * edges[i].dx;
        if (y < 0) continue;
        qsort(xs, na, sizeof(int64_t), cmp_i64);
        for (int k = 0; k + 1 < na; k += 2) {
            int64_t xa = (xs[k] + XY_ONE - 1) >> XY_SHIFT, xb = xs[k + 1] >> XY_SHIFT;
            if (xa < W && xb >= 0) {
                if (xa < 0) xa = 0;
                if (xb >= W) xb = W - 1;
                for (int64_t x = xa; x <= xb; ++x) mask[(size_t)y * W + x] = val;
            }
        }
    }
}

/* ========================================================================= */
/* OpenCV's symmetric Jacobi eigen-solver (core/lapack.cpp JacobiImpl_),      */
/* reached by cv::eigen, cv::solve(DECOMP_EIG), cv::invert(DECOMP_EIG).       */
/* A (n x n, row-major, destroyed), W eigenvalues (descending), V rows =      */
/* eigenvectors.                                                              */
/* ========================================================================= */
static double cv_hypot(double a, double b) {
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

#define ROT(v0, v1) do { double a0_ = (v0), b0_ = (v1); (v0) = a0_ * c - b0_ * s; (v1) = a0_ * s + b0_ * c; } while (0)

void orc_jacobi(double *A, double *W, double *V, int n) {
    const double eps = DBL_EPSILON;
    int indR[16], indC[16];
    int i, j, k, m;
    double mv;
    for (i = 0; i < n; ++i) { for (j = 0; j < n; ++j) V[i * n + j] = 0; V[i * n + i] = 1; }
    for (k = 0; k < n; ++k) {
        W[k] = A[(n + 1) * k];
        if (k < n - 1) {
            for (m = k + 1, mv = fabs(A[n * k + m]), i = k + 2; i < n; ++i) {
                double val = fabs(A[n * k + i]);
                if (mv < val) { mv = val; m = i; }
            }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = fabs(A[k]), i = 1; i < k; ++i) {
                double val = fabs(A[n * i + k]);
                if (mv < val) { mv = val; m = i; }
            }
            indC[k] = m;
        }
    }
    int maxIters = n * n * 30;
    if (n > 1) for (int iters = 0; iters < maxIters; ++iters) {
        for (k = 0, mv = fabs(A[indR[0]]), i = 1; i < n - 1; ++i) {
            double val = fabs(A[n * i + indR[i]]);
            if (mv < val) { mv = val; k = i; }
        }
        int l = indR[k];
        for (i = 1; i < n; ++i) {
            double val = fabs(A[n * indC[i] + i]);
            if (mv < val) { mv = val; k = indC[i]; l = i; }
        }
        double p = A[n * k + l];
        if (fabs(p) <= eps) break;
        double y = (W[l] - W[k]) * 0.5;
        double t = fabs(y) + cv_hypot(p, y);
        double s = cv_hypot(p, t);
        double c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        A[n * k + l] = 0;
        W[k] -= t;
        W[l] += t;
        for (i = 0; i < k; ++i) ROT(A[n * i + k], A[n * i + l]);
        for (i = k + 1; i < l; ++i) ROT(A[n * k + i], A[n * i + l]);
        for (i = l + 1; i < n; ++i) ROT(A[n * k + i], A[n * l + i]);
        for (i = 0; i < n; ++i) ROT(V[n * k + i], V[n * l + i]);
        for (j = 0; j < 2; ++j) {
            int idx = j == 0 ? k : l;
            if (idx < n - 1) {
                for (m = idx + 1, mv = fabs(A[n * idx + m]), i = idx + 2; i < n; ++i) {
                    double val = fabs(A[n * idx + i]);
                    if (mv < val) { mv = val; m = i; }
                }
                indR[idx] = m;
            }
            if (idx > 0) {
                for (m = 0, mv = fabs(A[idx]), i = 1; i < idx; ++i) {
                    double val = fabs(A[n * i + idx]);
                    if (mv < val) { mv = val; m = i; }
                }
                indC[idx] = m;
            }
        }
    }
    for (k = 0; k < n - 1; ++k) {
        m = k;
        for (i = k + 1; i < n; ++i) if (W[m] < W[i]) m = i;
        if (k != m) {
            double tw = W[m]; W[m] = W[k]; W[k] = tw;
            for (i = 0; i < n; ++i) { double tv = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = tv; }
        }
    }
}

/* cv::solve(A, b, x, DECOMP_EIG) for a symmetric n x n A and one right-hand side
 * (core/lapack.cpp: SVD-style back substitution on the Jacobi factors,
 *  threshold = DBL_EPSILON*2 * sum|w|).                                      */
static void eig_backsubst(const double *W, const double *V, int n, const double *b, double *x) {
    /* u == v == V^T (columns are eigenvectors); x = sum_i (v_i . b / w_i) v_i */
    double threshold = 0;
    for (int i = 0; i < n; ++i) threshold += W[i];
    threshold *= DBL_EPSILON * 2;   /* SVBkSbImpl_: eps*2 * sum(w) */
    for (int i = 0; i < n; ++i) x[i] = 0;
    for (int i = 0; i < n; ++i) {
        double wi = W[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < n; ++j) s += V[i * n + j] * b[j];
        s *= wi;
        for (int j = 0; j < n; ++j) x[j] = x[j] + s * V[i * n + j];
    }
}

void orc_solve_eig(const double *A, const double *b, double *x, int n) {
    double a[81], w[9], v[81];
    memcpy(a, A, sizeof(double) * n * n);
    orc_jacobi(a, w, v, n);
    eig_backsubst(w, v, n, b, x);
}

/* cv::invert(A, Ainv, DECOMP_EIG) */
void orc_invert_eig(const double *A, double *Ainv, int n) {
    double a[81], w[9], v[81], e[9], col[9];
    memcpy(a, A, sizeof(double) * n * n);
    orc_jacobi(a, w, v, n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) e[i] = (i == c);
        eig_backsubst(w, v, n, e, col);
        for (int i = 0; i < n; ++i) Ainv[i * n + c] = col[i];
    }
}

/* ========================================================================= */
/* cv2.findHomography(src, dst) method 0 (calib3d/fundam.cpp)                 */
/* returns 1 and fills H[9] (H[8] == 1), or 0 when OpenCV returns None.       */
/*                                                                           */
/* The refinement of n > 4 points is pinned against the installed             */
/* opencv-python 4.13.0 binary (scripts/check_lm_vs_cv2.py):    */
/* that build optimises ALL NINE entries of H (HomographyRefineCallback reads */
/* h[8] in the denominator and emits a 2n x 9 Jacobian), so J^T J is singular */
/* along the scale gauge and cv::solve(DECOMP_EIG)'s eigenvalue cut decides   */
/* the step; H is rescaled by 1/H[8] afterwards.                              */
/* ========================================================================= */
#define LM_NP 9
static void lm_compute(const float *M, const float *m, int count, const double *h, double *err, double *J) {
    for (int i = 0; i < count; ++i) {
        double Mx = M[2 * i], My = M[2 * i + 1];
        double ww = h[6] * Mx + h[7] * My + h[8];
        ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
        double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
        double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
        err[2 * i] = xi - m[2 * i];
        err[2 * i + 1] = yi - m[2 * i + 1];
        if (J) {
            double *Jp = J + 2 * LM_NP * i;
            Jp[0] = Mx * ww; Jp[1] = My * ww; Jp[2] = ww;
            Jp[3] = Jp[4] = Jp[5] = 0.;
            Jp[6] = -Mx * ww * xi; Jp[7] = -My * ww * xi; Jp[8] = -ww * xi;
            Jp[9] = Jp[10] = Jp[11] = 0.;
            Jp[12] = Mx * ww; Jp[13] = My * ww; Jp[14] = ww;
            Jp[15] = -Mx * ww * yi; Jp[16] = -My * ww * yi; Jp[17] = -ww * yi;
        }
    }
}

/* cv::gemm row . vector (GEMMSingleMul, one-column result): 4 interleaved accumulators */
static double gemm_rowdot(const double *a, int astride, const double *b, int n) {
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

static void jtj_jtr(const double *J, const double *r, int rows, double *A, double *v) {
    /* A = J^T J (cv::mulTransposed, sequential over rows); v = J^T r (cv::gemm GEMM_1_T) */
    for (int i = 0; i < LM_NP; ++i)
        for (int j = i; j < LM_NP; ++j) {
            double s = 0;
            for (int k = 0; k < rows; ++k) s += J[k * LM_NP + i] * J[k * LM_NP + j];
            A[i * LM_NP + j] = s; A[j * LM_NP + i] = s;
        }
    for (int i = 0; i < LM_NP; ++i) v[i] = gemm_rowdot(J + i, LM_NP, r, rows);
}

/* cv::Mat::dot on CV_64F as the 4.13.0 wheel executes it on an FMA-capable CPU (the AVX2
 * dispatch of dotProd_ is compiled with contraction): per group of four
 * t = fma(a0,b0, a1*b1); t = fma(a2,b2,t); t = fma(a3,b3,t); res += t, tail res = fma(a,b,res). */
static double cv_dot(const double *a, const double *b, int n) {
    double res = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        double t = fma(a[i], b[i], a[i + 1] * b[i + 1]);
        t = fma(a[i + 2], b[i + 2], t);
        t = fma(a[i + 3], b[i + 3], t);
        res += t;
    }
    for (; i < n; ++i) res = fma(a[i], b[i], res);
    return res;
}

static double norm_l2sqr(const double *r, int n) { double s = 0; for (int i = 0; i < n; ++i) s += r[i] * r[i]; return s; }
static double norm_inf(const double *r, int n) { double s = 0; for (int i = 0; i < n; ++i) { double a = fabs(r[i]); if (a > s) s = a; } return s; }

/* LM trace hook for scripts/check_lm_vs_cv2.py: number of outer iterations of the last call */
int orc_lm_last_iters = 0;
int orc_lm_iters(void) { return orc_lm_last_iters; }

static void lm_refine(const float *M, const float *m, int count, double *h9) {
    /* calib3d/levmarq.cpp LMSolverImpl::run, maxIters = 10, epsx = epsf = FLT_EPSILON */
    const int lx = LM_NP, rows = 2 * count;
    const int maxIters = 10;
    const double epsx = FLT_EPSILON, epsf = FLT_EPSILON;
    double x[LM_NP], xd[LM_NP], r[32], rd[32], J[32 * LM_NP], A[LM_NP * LM_NP], Ap[LM_NP * LM_NP], v[LM_NP], d[LM_NP], D[LM_NP], temp_d[LM_NP];
    memcpy(x, h9, sizeof(x));
    lm_compute(M, m, count, x, r, J);
    double S = norm_l2sqr(r, rows);
    jtj_jtr(J, r, rows, A, v);
    for (int i = 0; i < lx; ++i) D[i] = A[i * lx + i];
    const double Rlo = 0.25, Rhi = 0.75;
    double lambda = 1, lc = 0.75;
    int iter = 0;
    for (;;) {
        memcpy(Ap, A, sizeof(A));
        for (int i = 0; i < lx; ++i) Ap[i * lx + i] += lambda * D[i];
        orc_solve_eig(Ap, v, d, lx);
        for (int i = 0; i < lx; ++i) xd[i] = x[i] - d[i];
        lm_compute(M, m, count, xd, rd, NULL);
        double Sd = norm_l2sqr(rd, rows);
        /* temp_d = -A d + 2 v  (cv::gemm(A, d, -1, v, 2)) */
        for (int i = 0; i < lx; ++i) temp_d[i] = -1. * gemm_rowdot(A + i * lx, 1, d, lx) + 2. * v[i];
        double dS = cv_dot(d, temp_d, lx);
        double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
        if (R > Rhi) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0;
        } else if (R < Rlo) {
            double t = cv_dot(d, v, lx);
            double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
            nu = fmin(fmax(nu, 2.), 10.);
            if (lambda == 0) {
                double maxval = DBL_EPSILON;
                orc_invert_eig(A, Ap, lx);
                for (int i = 0; i < lx; ++i) maxval = fmax(maxval, fabs(Ap[i * lx + i]));
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        if (Sd < S) {
            S = Sd;
            memcpy(x, xd, sizeof(x));
            lm_compute(M, m, count, x, r, J);
            jtj_jtr(J, r, rows, A, v);
        }
        iter++;
        int proceed = iter < maxIters && norm_inf(d, lx) >= epsx && norm_inf(r, rows) >= epsf;
        if (!proceed) break;
    }
    orc_lm_last_iters = iter;
    memcpy(h9, x, sizeof(x));
}

int orc_find_homography(const int32_t *src, const int32_t *dst, int count, double *H, int refine) {
    float M[16 * 2], m[16 * 2];
    if (count < 4 || count > 16) return 0;
    for (int i = 0; i < 2 * count; ++i) { M[i] = (float)src[i]; m[i] = (float)dst[i]; }
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
    for (int i = 0; i < count; ++i) {
        cmx += m[2 * i]; cmy += m[2 * i + 1];
        cMx += M[2 * i]; cMy += M[2 * i + 1];
    }
    cmx /= count; cmy /= count; cMx /= count; cMy /= count;
    for (int i = 0; i < count; ++i) {
        smx += fabs(m[2 * i] - cmx); smy += fabs(m[2 * i + 1] - cmy);
        sMx += fabs(M[2 * i] - cMx); sMy += fabs(M[2 * i + 1] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON)
        return 0;
    smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
    double invHnorm[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
    double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double LtL[81], W[9], V[81];
    memset(LtL, 0, sizeof(LtL));
    for (int i = 0; i < count; ++i) {
        double x = (m[2 * i] - cmx) * smx, y = (m[2 * i + 1] - cmy) * smy;
        double X = (M[2 * i] - cMx) * sMx, Y = (M[2 * i + 1] - cMy) * sMy;
        double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
        double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
        for (int j = 0; j < 9; ++j)
            for (int k = j; k < 9; ++k)
                LtL[j * 9 + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
    }
    for (int j = 0; j < 9; ++j) for (int k = 0; k < j; ++k) LtL[j * 9 + k] = LtL[k * 9 + j];
    orc_jacobi(LtL, W, V, 9);
    const double *H0 = V + 72; /* eigenvector of the smallest eigenvalue */
    double Ht[9], H1[9];
    /* Htemp = invHnorm * H0 ; H0 = Htemp * Hnorm2  (cv::gemm 3x3, sequential k) */
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += invHnorm[i * 3 + k] * H0[k * 3 + j];
        Ht[i * 3 + j] = s;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += Ht[i * 3 + k] * Hnorm2[k * 3 + j];
        H1[i * 3 + j] = s;
    }
    double sc = 1. / H1[8];
    for (int i = 0; i < 9; ++i) H[i] = H1[i] * sc;
    if (count > 4 && refine) {
        lm_refine(M, m, count, H);
        /* H.convertTo(H, H.type(), scaleFor(H(2,2))) */
        double sc2 = fabs(H[8]) > DBL_EPSILON ? 1. / H[8] : 1.;
        for (int i = 0; i < 9; ++i) H[i] = H[i] * sc2;
    }
    return 1;
}

/* ========================================================================= */
/* cv2.warpPerspective(src u8 HxWx3, H, (W,H)), INTER_LINEAR, BORDER_CONSTANT */
/* (imgproc/imgwarp.cpp WarpPerspectiveInvoker + remapBilinear, fixed point)  */
/* tapmask (optional, H*W bytes): a source tap is read as 0 where tapmask==0  */
/* -- this fuses planes_utils.py:31 (image*mask) into the gather.             */
/* ========================================================================= */
static int invert3(const double *S, double *T) {
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.) return 0;
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
    return 1;
}

int orc_invert3(const double *S, double *T) { return invert3(S, T); }

void orc_warp_perspective(const uint8_t *src, const uint8_t *tapmask, int H, int W, const double *Hm, uint8_t *dst) {
    double M[9];
    if (!invert3(Hm, M)) memset(M, 0, sizeof(M));  /* cv::invert leaves zeros on a singular matrix */
    /* block decomposition of WarpPerspectiveInvoker (BLOCK_SZ = 32) */
    int bh0 = 16 < H ? 16 : H;
    int bw0 = (32 * 32 / bh0) < W ? (32 * 32 / bh0) : W;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            int bx = (x / bw0) * bw0, x1 = x - bx;
            double X0 = M[0] * bx + M[1] * y + M[2];
            double Y0 = M[3] * bx + M[4] * y + M[5];
            double W0 = M[6] * bx + M[7] * y + M[8];
            double Wv = W0 + M[6] * x1;
            Wv = Wv ? 32. / Wv : 0;
            double fX = fmax((double)INT_MIN, fmin((double)INT_MAX, (X0 + M[0] * x1) * Wv));
            double fY = fmax((double)INT_MIN, fmin((double)INT_MAX, (Y0 + M[3] * x1) * Wv));
            int X = (int)lrint(fX), Y = (int)lrint(fY);   /* cvRound: round-half-even */
            int sx = X >> 5, sy = Y >> 5, a = X & 31, b = Y & 31;
            /* saturate_cast<short> on the integer part */
            if (sx < -32768) sx = -32768; if (sx > 32767) sx = 32767;
            if (sy < -32768) sy = -32768; if (sy > 32767) sy = 32767;
            int w00 = (32 - a) * (32 - b) * 32, w01 = a * (32 - b) * 32, w10 = (32 - a) * b * 32, w11 = a * b * 32;
            uint8_t *o = dst + ((size_t)y * W + x) * 3;
            int in00 = sx >= 0 && sx < W && sy >= 0 && sy < H;
            int in01 = sx + 1 >= 0 && sx + 1 < W && sy >= 0 && sy < H;
            int in10 = sx >= 0 && sx < W && sy + 1 >= 0 && sy + 1 < H;
            int in11 = sx + 1 >= 0 && sx + 1 < W && sy + 1 >= 0 && sy + 1 < H;
            if (tapmask) {
                if (in00) in00 = tapmask[(size_t)sy * W + sx] != 0;
                if (in01) in01 = tapmask[(size_t)sy * W + sx + 1] != 0;
                if (in10) in10 = tapmask[(size_t)(sy + 1) * W + sx] != 0;
                if (in11) in11 = tapmask[(size_t)(sy + 1) * W + sx + 1] != 0;
            }
            for (int c = 0; c < 3; ++c) {
                int v00 = in00 ? src[((size_t)sy * W + sx) * 3 + c] : 0;
                int v01 = in01 ? src[((size_t)sy * W + sx + 1) * 3 + c] : 0;
                int v10 = in10 ? src[((size_t)(sy + 1) * W + sx) * 3 + c] : 0;
                int v11 = in11 ? src[((size_t)(sy + 1) * W + sx + 1) * 3 + c] : 0;
                o[c] = (uint8_t)((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15);
            }
        }
    }
}

/* ========================================================================= */
/* online_visibility.py                                                       */
/* ========================================================================= */

/* project_points (:28-56) for one point.  Spec'd evaluation order (numpy's BLAS
 * order is not defined): P = K @ E[:3] with k ascending, q = P @ [X Y Z 1] with k
 * ascending, then q/q[2].  E is 3x4 row-major. */
static void project_point(const double *K, const double *E, const double *X, double *uv) {
    double P[12], q[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) {
        double s = K[i * 3 + 0] * E[0 * 4 + j];
        s = s + K[i * 3 + 1] * E[1 * 4 + j];
        s = s + K[i * 3 + 2] * E[2 * 4 + j];
        P[i * 4 + j] = s;
    }
    for (int i = 0; i < 3; ++i) {
        double s = P[i * 4 + 0] * X[0];
        s = s + P[i * 4 + 1] * X[1];
        s = s + P[i * 4 + 2] * X[2];
        s = s + P[i * 4 + 3] * 1.0;
        q[i] = s;
    }
    uv[0] = q[0] / q[2];
    uv[1] = q[1] / q[2];
}

void orc_project_points(const double *K, const double *E, const double *kp3d, int n, double *uv) {
    for (int i = 0; i < n; ++i) project_point(K, E, kp3d + 3 * i, uv + 2 * i);
}

/* camera_planes_dist (:59-75): camera centre = inv(E)[:3,3] = -R^T t for a rigid
 * E (spec'd closed form; numpy's LAPACK inverse differs only in the last ulps). */
void orc_plane_distances(const double *E, const double *kp3d, double *dist) {
    double c[3];
    for (int i = 0; i < 3; ++i) {
        double s = E[0 * 4 + i] * E[0 * 4 + 3];
        s = s + E[1 * 4 + i] * E[1 * 4 + 3];
        s = s + E[2 * 4 + i] * E[2 * 4 + 3];
        c[i] = -s;
    }
    for (int p = 0; p < N_VIS; ++p) {
        double mean[3] = {0, 0, 0};
        for (int k = 0; k < PLANE_N[p]; ++k)
            for (int a = 0; a < 3; ++a) mean[a] += kp3d[3 * PLANE_KP[p][k] + a];
        double s = 0;
        for (int a = 0; a < 3; ++a) {
            mean[a] /= PLANE_N[p];
            double d = c[a] - mean[a];
            s += d * d;
        }
        dist[p] = sqrt(s);
    }
}

/* visibility from already-truncated vertices + distances (:105-150) */
void orc_visibility_from_pts(const int32_t *pts /*12x2*/, const double *dist /*7*/, int h, int w, uint8_t *vis /*7*/,
                             int32_t *areas /*optional 14: abs,occ per plane*/) {
    uint8_t *plane = (uint8_t *)malloc((size_t)h * w);
    uint8_t *occl = (uint8_t *)malloc((size_t)h * w);
    for (int p = 0; p < N_VIS; ++p) {
        int32_t poly[12], q[12];
        for (int k = 0; k < PLANE_N[p]; ++k) { poly[2 * k] = pts[2 * PLANE_KP[p][k]]; poly[2 * k + 1] = pts[2 * PLANE_KP[p][k] + 1]; }
        memset(plane, 0, (size_t)h * w);
        orc_fill_poly(plane, h, w, poly, PLANE_N[p], 255);
        memcpy(occl, plane, (size_t)h * w);
        for (int o = 0; o < N_VIS; ++o) {
            if (dist[o] < dist[p]) {
                for (int k = 0; k < PLANE_N[o]; ++k) { q[2 * k] = pts[2 * PLANE_KP[o][k]]; q[2 * k + 1] = pts[2 * PLANE_KP[o][k] + 1]; }
                orc_fill_poly(occl, h, w, q, PLANE_N[o], 0);
            }
        }
        long a_abs = 0, a_occ = 0;
        for (size_t i = 0; i < (size_t)h * w; ++i) { a_abs += plane[i] > 0; a_occ += occl[i] > 0; }
        vis[p] = (double)a_occ > 0.9 * (double)a_abs;
        if (areas) { areas[2 * p] = (int32_t)a_abs; areas[2 * p + 1] = (int32_t)a_occ; }
    }
    free(plane); free(occl);
}

/* compute_visibility(extrinsic, intrinsic, kpoints_3d, h, w) (:105-150).
 * E 3x4 row-major, K 3x3, kp3d 12x3 in _KP_NAMES order.  Also returns the int()
 * truncated projections so tests can compare them with numpy's. */
void orc_compute_visibility(const double *E, const double *K, const double *kp3d, int h, int w, uint8_t *vis,
                            int32_t *pts_out /*optional 24*/) {
    double uv[24], dist[N_VIS];
    int32_t pts[24];
    orc_project_points(K, E, kp3d, N_KP, uv);
    for (int i = 0; i < 24; ++i) pts[i] = (int32_t)uv[i];   /* int(): truncate toward zero */
    orc_plane_distances(E, kp3d, dist);
    orc_visibility_from_pts(pts, dist, h, w, vis, NULL);
    if (pts_out) memcpy(pts_out, pts, sizeof(pts));
}

/* ========================================================================= */
/* planes_utils.py                                                            */
/* ========================================================================= */

/* get_planes (:11-37): kp is the already int32-truncated 12x2 array in _KP_NAMES
 * order; writes planes (5,H,W,3) = image * fillPoly mask. */
void orc_get_planes(const uint8_t *img, int H, int W, const int32_t *kp, uint8_t *planes) {
    uint8_t *mask = (uint8_t *)malloc((size_t)H * W);
    for (int p = 0; p < N_TEX; ++p) {
        int32_t poly[12];
        for (int k = 0; k < PLANE_N[p]; ++k) { poly[2 * k] = kp[2 * PLANE_KP[p][k]]; poly[2 * k + 1] = kp[2 * PLANE_KP[p][k] + 1]; }
        memset(mask, 0, (size_t)H * W);
        orc_fill_poly(mask, H, W, poly, PLANE_N[p], 1);
        uint8_t *o = planes + (size_t)p * H * W * 3;
        for (size_t i = 0; i < (size_t)H * W; ++i)
            for (int c = 0; c < 3; ++c) o[3 * i + c] = mask[i] ? img[3 * i + c] : 0;
    }
    free(mask);
}

/* gating + symmetry remap of warp_unwarp_planes (:57-68): returns target j or -1 */
static int plane_target(int i, const uint8_t *src_vis, const uint8_t *dst_vis) {
    if (!src_vis[i]) return -1;
    if (i >= 2 && !dst_vis[i]) return -1;
    if (i < 2 && !(dst_vis[0] == 1 || dst_vis[1] == 1)) return -1;
    int j = i;
    if (i < 2 && !dst_vis[i]) j = 1 - i;
    return j;
}

/* warp_unwarp_planes (:40-82) on explicit plane images (the literal drop-in).
 * src_kp/dst_kp: 12x2 int32 (_KP_NAMES order).  plane_j[i] = target index or -1. */
void orc_warp_unwarp_planes(const uint8_t *src_planes, int H, int W, const int32_t *src_kp, const int32_t *dst_kp,
                            const uint8_t *src_vis, const uint8_t *dst_vis, uint8_t *warped, uint8_t *unwarped,
                            int8_t *plane_j, double *H12_out /*5x9 optional*/) {
    size_t psz = (size_t)H * W * 3;
    memset(warped, 0, psz * N_TEX);
    if (unwarped) memset(unwarped, 0, psz * N_TEX);
    uint8_t *tmp = (uint8_t *)malloc(psz);
    for (int i = 0; i < N_TEX; ++i) {
        plane_j[i] = -1;
        if (H12_out) for (int k = 0; k < 9; ++k) H12_out[9 * i + k] = 0;
        int j = plane_target(i, src_vis, dst_vis);
        if (j < 0) continue;
        int32_t s[12], d[12];
        int n = PLANE_N[i];
        for (int k = 0; k < n; ++k) {
            s[2 * k] = src_kp[2 * PLANE_KP[i][k]]; s[2 * k + 1] = src_kp[2 * PLANE_KP[i][k] + 1];
            d[2 * k] = dst_kp[2 * PLANE_KP[j][k]]; d[2 * k + 1] = dst_kp[2 * PLANE_KP[j][k] + 1];
        }
        double H12[9], H21[9];
        int ok12 = orc_find_homography(s, d, n, H12, 1);
        int ok21 = orc_find_homography(d, s, n, H21, 1);
        if (!(ok12 && ok21)) continue;
        plane_j[i] = (int8_t)j;
        if (H12_out) memcpy(H12_out + 9 * i, H12, sizeof(H12));
        orc_warp_perspective(src_planes + psz * i, NULL, H, W, H12, tmp);
        memcpy(warped + psz * j, tmp, psz);
        if (unwarped) orc_warp_perspective(tmp, NULL, H, W, H21, unwarped + psz * i);
    }
    free(tmp);
}

/* The fused batch item: visibility(src), visibility(dst), get_planes masks, gating,
 * homographies and the masked forward warp, without materialising the 5 source
 * planes or the (discarded) unwarp.  Equivalent to
 *   compute_visibility x2 -> get_planes -> warp_unwarp_planes()[0]
 * (trajectory_inference.py:165-174).  Returns 0, or -1 if a keypoint is outside
 * the frame (the clipped-polygon regime is not covered by the fused path). */
int orc_warp_fused(const uint8_t *src, int H, int W, const int32_t *src_kp, const int32_t *dst_kp,
                   const double *K, const double *E_src, const double *E_dst, const double *kp3d,
                   uint8_t *warped, uint8_t *vis_out /*2x7*/, int8_t *plane_j /*5*/, double *H12_out /*5x9*/) {
    orc_compute_visibility(E_src, K, kp3d, H, W, vis_out, NULL);
    orc_compute_visibility(E_dst, K, kp3d, H, W, vis_out + N_VIS, NULL);
    size_t psz = (size_t)H * W * 3;
    memset(warped, 0, psz * N_TEX);
    uint8_t *mask = (uint8_t *)malloc((size_t)H * W);
    for (int i = 0; i < N_TEX; ++i) {
        plane_j[i] = -1;
        for (int k = 0; k < 9; ++k) H12_out[9 * i + k] = 0;
        int j = plane_target(i, vis_out, vis_out + N_VIS);
        if (j < 0) continue;
        int32_t s[12], d[12];
        int n = PLANE_N[i];
        for (int k = 0; k < n; ++k) {
            s[2 * k] = src_kp[2 * PLANE_KP[i][k]]; s[2 * k + 1] = src_kp[2 * PLANE_KP[i][k] + 1];
            d[2 * k] = dst_kp[2 * PLANE_KP[j][k]]; d[2 * k + 1] = dst_kp[2 * PLANE_KP[j][k] + 1];
        }
        double H12[9], H21[9];
        if (!(orc_find_homography(s, d, n, H12, 1) && orc_find_homography(d, s, n, H21, 1))) continue;
        plane_j[i] = (int8_t)j;
        memcpy(H12_out + 9 * i, H12, sizeof(H12));
        memset(mask, 0, (size_t)H * W);
        orc_fill_poly(mask, H, W, s, n, 1);
        orc_warp_perspective(src, mask, H, W, H12, warped + psz * j);
    }
    free(mask);
    return 0;
}
