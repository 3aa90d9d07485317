// warp_geom_thread.cuh -- thread-per-point-set variant of cv2.findHomography (same arithmetic as
// warp_geom.cuh's warp-cooperative solver, one thread per solve).  The large matrices (9x9 LtL / V of the
// Jacobi solve, then J / A / Ap / L of the LM refinement, 288 doubles per thread) live in SHARED memory,
// element i of thread t at base[i * 32 + t]: with 255 registers and 3.2 KB of local arrays per thread the
// first version kept only 8 warps per SM resident and still thrashed L1 (ncu: 11 long-scoreboard stalls per
// issue, 485 MB of local-memory write-back per launch); the functions are templates over the array type so a
// plain `double *` (local memory) instantiates the same arithmetic.
// Higher throughput when tens of thousands of solves are in flight (BASELINE config 3); the
// warp-cooperative variant has ~3x lower latency for small batches.  Both are bit-identical to the
// oracle (tests/test_warp_gpu.py).
#pragma once
#include "warp_geom.cuh"

namespace fusg {

// array of doubles interleaved over the S threads of a block: element i at p[i * S]
template <int S>
struct StridedArr {
    double *p;
    __device__ __forceinline__ double &operator[](int i) const { return p[(size_t)i * S]; }
    __device__ __forceinline__ StridedArr operator+(int off) const { return StridedArr{p + (size_t)off * S}; }
};

// ---------------------------------------------------------------------------------------------
// OpenCV's Jacobi eigen-solver (cv::eigen on a symmetric matrix), n <= 9.
// A is destroyed; W = eigenvalues (descending); rows of V = eigenvectors.
// ---------------------------------------------------------------------------------------------

template <class AT, class VT>
__device__ inline void jacobi_eig(AT A, double *W, VT V, const int n) {
    const double eps = DBL_EPSILON;
    int indR[9], indC[9];
    int i, j, k, m;
    double mv;
    for (i = 0; i < n; ++i) { for (j = 0; j < n; ++j) V[i * n + j] = 0; V[i * n + i] = 1; }
    for (k = 0; k < n; ++k) {
        W[k] = A[(n + 1) * k];
        if (k < n - 1) {
            for (m = k + 1, mv = fabs(A[n * k + m]), i = k + 2; i < n; ++i) {
                const double val = fabs(A[n * k + i]);
                if (mv < val) { mv = val; m = i; }
            }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = fabs(A[k]), i = 1; i < k; ++i) {
                const double val = fabs(A[n * i + k]);
                if (mv < val) { mv = val; m = i; }
            }
            indC[k] = m;
        }
    }
    const int maxIters = n * n * 30;
    if (n > 1) for (int iters = 0; iters < maxIters; ++iters) {
        for (k = 0, mv = fabs(A[indR[0]]), i = 1; i < n - 1; ++i) {
            const double val = fabs(A[n * i + indR[i]]);
            if (mv < val) { mv = val; k = i; }
        }
        int l = indR[k];
        for (i = 1; i < n; ++i) {
            const double val = fabs(A[n * indC[i] + i]);
            if (mv < val) { mv = val; k = indC[i]; l = i; }
        }
        const double p = A[n * k + l];
        if (fabs(p) <= eps) break;
        const double y = (W[l] - W[k]) * 0.5;
        double t = fabs(y) + cv_hypot(p, y);
        double s = cv_hypot(p, t);
        const double c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        A[n * k + l] = 0;
        W[k] -= t;
        W[l] += t;
        for (i = 0; i < k; ++i) FUSG_ROT(A[n * i + k], A[n * i + l]);
        for (i = k + 1; i < l; ++i) FUSG_ROT(A[n * k + i], A[n * i + l]);
        for (i = l + 1; i < n; ++i) FUSG_ROT(A[n * k + i], A[n * l + i]);
        for (i = 0; i < n; ++i) FUSG_ROT(V[n * k + i], V[n * l + i]);
        for (j = 0; j < 2; ++j) {
            const int idx = j == 0 ? k : l;
            if (idx < n - 1) {
                for (m = idx + 1, mv = fabs(A[n * idx + m]), i = idx + 2; i < n; ++i) {
                    const double val = fabs(A[n * idx + i]);
                    if (mv < val) { mv = val; m = i; }
                }
                indR[idx] = m;
            }
            if (idx > 0) {
                for (m = 0, mv = fabs(A[idx]), i = 1; i < idx; ++i) {
                    const double val = fabs(A[n * i + idx]);
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
            const double tw = W[m]; W[m] = W[k]; W[k] = tw;
            for (i = 0; i < n; ++i) { const double tv = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = tv; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LM refinement of the 8 free homography parameters (OpenCV LMSolver schedule, maxIters 10,
// eps FLT_EPSILON); linear systems by square-root-free Cholesky.
// ---------------------------------------------------------------------------------------------
template <class JT>
__device__ inline void lm_residual(const float *M, const float *m, int count, const double *h, double *err, JT J, bool want_J) {
    for (int i = 0; i < count; ++i) {
        const double Mx = M[2 * i], My = M[2 * i + 1];
        double ww = h[6] * Mx + h[7] * My + 1.;
        ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
        const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
        const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
        err[2 * i] = xi - m[2 * i];
        err[2 * i + 1] = yi - m[2 * i + 1];
        if (want_J) {
            JT Jp = J + 16 * i;
            Jp[0] = Mx * ww; Jp[1] = My * ww; Jp[2] = ww;
            Jp[3] = Jp[4] = Jp[5] = 0.;
            Jp[6] = -Mx * ww * xi; Jp[7] = -My * ww * xi;
            Jp[8] = Jp[9] = Jp[10] = 0.;
            Jp[11] = Mx * ww; Jp[12] = My * ww; Jp[13] = ww;
            Jp[14] = -Mx * ww * yi; Jp[15] = -My * ww * yi;
        }
    }
}

template <class JT, class AT>
__device__ inline void lm_normal_eq(JT J, const double *r, int rows, AT A, double *v) {
    for (int i = 0; i < 8; ++i)
        for (int j = i; j < 8; ++j) {
            double s = 0;
            for (int k = 0; k < rows; ++k) s += J[k * 8 + i] * J[k * 8 + j];
            A[i * 8 + j] = s; A[j * 8 + i] = s;
        }
    for (int i = 0; i < 8; ++i) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int k = 0;
        for (; k <= rows - 4; k += 4) {
            s0 += J[k * 8 + i] * r[k];
            s1 += J[(k + 1) * 8 + i] * r[k + 1];
            s2 += J[(k + 2) * 8 + i] * r[k + 2];
            s3 += J[(k + 3) * 8 + i] * r[k + 3];
        }
        for (; k < rows; ++k) s0 += J[k * 8 + i] * r[k];
        v[i] = ((s0 + s1) + s2) + s3;
    }
}

__device__ inline double dot4(const double *a, const double *b, int n) {
    double res = 0;
    int i = 0;
    for (; i <= n - 4; i += 4)
        res += a[i] * b[i] + a[i + 1] * b[i + 1] + a[i + 2] * b[i + 2] + a[i + 3] * b[i + 3];
    for (; i < n; ++i) res += a[i] * b[i];
    return res;
}

template <class AT, class LT>
__device__ inline void ldl_solve8_serial(AT A, const double *b, double *x, LT L) {
    double Dg[8], y[8];
    for (int j = 0; j < 8; ++j) {
        double dj = A[j * 8 + j];
        for (int k = 0; k < j; ++k) dj -= L[j * 8 + k] * L[j * 8 + k] * Dg[k];
        Dg[j] = dj;
        const double inv = dj > 0 ? 1. / dj : 0.;
        for (int i = j + 1; i < 8; ++i) {
            double s = A[i * 8 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 8 + k] * L[j * 8 + k] * Dg[k];
            L[i * 8 + j] = s * inv;
        }
    }
    for (int i = 0; i < 8; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * 8 + k] * y[k];
        y[i] = s;
    }
    for (int i = 0; i < 8; ++i) y[i] = Dg[i] > 0 ? y[i] / Dg[i] : 0.;
    for (int i = 7; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < 8; ++k) s -= L[k * 8 + i] * x[k];
        x[i] = s;
    }
}

template <class MT>
__device__ inline void lm_refine(const float *M, const float *m, int count, double *h8, MT mem) {
    const int lx = 8, rows = 2 * count;
    const int maxIters = 10;
    const double epsx = FLT_EPSILON, epsf = FLT_EPSILON;
    double x[8], xd[8], r[12], rd[12], v[8], d[8], D[8], temp_d[8];
    MT J = mem, A = mem + 96, Ap = mem + 160, L = mem + 224;        // 12x8 | 8x8 | 8x8 | 8x8
    for (int i = 0; i < 8; ++i) x[i] = h8[i];
    lm_residual(M, m, count, x, r, J, true);
    double S = 0;
    for (int i = 0; i < rows; ++i) S += r[i] * r[i];
    lm_normal_eq(J, r, rows, A, v);
    for (int i = 0; i < lx; ++i) D[i] = A[i * 8 + i];
    const double Rlo = 0.25, Rhi = 0.75;
    double lambda = 1, lc = 0.75;
    int iter = 0;
    for (;;) {
        for (int i = 0; i < 64; ++i) Ap[i] = A[i];
        for (int i = 0; i < lx; ++i) Ap[i * 8 + i] += lambda * D[i];
        ldl_solve8_serial(Ap, v, d, L);
        for (int i = 0; i < lx; ++i) xd[i] = x[i] - d[i];
        lm_residual(M, m, count, xd, rd, J, false);
        double Sd = 0;
        for (int i = 0; i < rows; ++i) Sd += rd[i] * rd[i];
        for (int i = 0; i < lx; ++i) {
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int k = 0; k < lx; k += 4) {
                s0 += A[i * 8 + k] * d[k];
                s1 += A[i * 8 + k + 1] * d[k + 1];
                s2 += A[i * 8 + k + 2] * d[k + 2];
                s3 += A[i * 8 + k + 3] * d[k + 3];
            }
            temp_d[i] = -1. * (((s0 + s1) + s2) + s3) + 2. * v[i];
        }
        const double dS = dot4(d, temp_d, lx);
        const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
        if (R > Rhi) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0;
        } else if (R < Rlo) {
            const double t = dot4(d, v, lx);
            double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
            nu = fmin(fmax(nu, 2.), 10.);
            if (lambda == 0) {
                double maxval = DBL_EPSILON;
                for (int c = 0; c < lx; ++c) {
                    double e[8] = {0, 0, 0, 0, 0, 0, 0, 0}, col[8];
                    e[c] = 1.;
                    ldl_solve8_serial(A, e, col, L);
                    maxval = fmax(maxval, fabs(col[c]));
                }
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        if (Sd < S) {
            S = Sd;
            for (int i = 0; i < 8; ++i) x[i] = xd[i];
            lm_residual(M, m, count, x, r, J, true);
            lm_normal_eq(J, r, rows, A, v);
        }
        iter++;
        double nd = 0, nr = 0;
        for (int i = 0; i < lx; ++i) nd = fmax(nd, fabs(d[i]));
        for (int i = 0; i < rows; ++i) nr = fmax(nr, fabs(r[i]));
        if (!(iter < maxIters && nd >= epsx && nr >= epsf)) break;
    }
    for (int i = 0; i < 8; ++i) h8[i] = x[i];
}

// OpenCV's "returns None" test of findHomography: all src or all dst points share an x or a y.
__device__ inline bool homography_degenerate(const int *s, const int *d, int count) {
    float M[12], m[12];
    for (int i = 0; i < 2 * count; ++i) { M[i] = (float)s[i]; m[i] = (float)d[i]; }
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
    return fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON;
}

// cv2.findHomography(src, dst), method 0, count in {4..6}.  Returns false where OpenCV returns None.
// mem: 288 doubles of scratch (StridedArr over shared memory, or a plain double * to local memory).
template <class MT>
__device__ inline bool find_homography_thread(const int *s, const int *d, int count, double *H, MT mem) {
    float M[12], m[12];
    for (int i = 0; i < 2 * count; ++i) { M[i] = (float)s[i]; m[i] = (float)d[i]; }
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
        return false;
    smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
    const double invHnorm[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
    const double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double W[9];
    MT LtL = mem, V = mem + 81;
    for (int i = 0; i < 81; ++i) LtL[i] = 0;
    for (int i = 0; i < count; ++i) {
        const double x = (m[2 * i] - cmx) * smx, y = (m[2 * i + 1] - cmy) * smy;
        const double X = (M[2 * i] - cMx) * sMx, Y = (M[2 * i + 1] - cMy) * sMy;
        const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
        const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
        for (int j = 0; j < 9; ++j)
            for (int k = j; k < 9; ++k)
                LtL[j * 9 + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
    }
    for (int j = 0; j < 9; ++j) for (int k = 0; k < j; ++k) LtL[j * 9 + k] = LtL[k * 9 + j];
    jacobi_eig(LtL, W, V, 9);
    double H0[9], Ht[9], H1[9];
    for (int i = 0; i < 9; ++i) H0[i] = V[72 + i];                 // last eigenvector, before the scratch is reused
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += invHnorm[i * 3 + k] * H0[k * 3 + j];
        Ht[i * 3 + j] = acc;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += Ht[i * 3 + k] * Hnorm2[k * 3 + j];
        H1[i * 3 + j] = acc;
    }
    const double sc = 1. / H1[8];
    for (int i = 0; i < 9; ++i) H[i] = H1[i] * sc;
    if (count > 4) {
        lm_refine(M, m, count, H, mem);
        H[8] = 1.;
    }
    return true;
}


}  // namespace fusg
