// warp_solver.cuh -- cv2.findHomography(src, dst) (method 0) on the device, ONE THREAD PER POINT SET.
//
// Reference behaviour: warp_learn/planes_utils.py:71-72 calls cv2.findHomography on the 4 or 6 vertices of a texture
// plane.  OpenCV 4.13.0 runs: normalised DLT (9x9 L^T L, cv::eigen = Jacobi) and, for more than four points, ten
// iterations at most of LMSolver on all nine entries of H, each of which solves the (singular) 9x9 normal equations
// with cv::solve(DECOMP_EIG) -- another Jacobi eigen-decomposition -- and occasionally inverts them the same way
// (oracle/warp_oracle.c restates this and is pinned bit for bit against cv2, scripts/check_lm_vs_cv2.py).  One 6-point
// solve is therefore a chain of ~10 Jacobi runs (~1150 dependent plane rotations); every fp64 operation has to happen
// in OpenCV's order, so the only parallelism is ACROSS point sets.
//
// Layout of the work:
//   * a thread owns a point set; the 32 lanes of a warp run the same instruction stream on 32 point sets;
//   * every per-thread matrix lives in shared memory as element e of lane t at base[e * 32 + t]: any dynamic index
//     (the Jacobi pivot differs per lane) stays bank-conflict free;
//   * the Jacobi sweep is branch-free (symmetric min/max indexing, predicated stores), so lanes with different pivots
//     stay converged; a lane that has converged idles until the slowest lane of its warp has;
//   * the solve is a small STATE MACHINE whose unit of work is "one Jacobi run on the matrix in A": the DLT
//     decomposition, each LM solve and each LM inversion are such jobs, so a lane that needs an inversion while its
//     neighbours need a solve still shares the Jacobi code with them.
//
// Compiled with -fmad=false: no contraction; the two places where OpenCV's binary itself uses FMA (Mat::dot) call
// __fma_rn explicitly.
#pragma once
#include <cfloat>
#include <cstdint>

namespace fusg {

// per-warp scratch, in doubles per lane.  JacobiImpl_ only ever touches the upper triangle of its (symmetric) input and keeps
// the diagonal in W, and J^T J is symmetric: both are stored packed, which is what lets FOUR solver warps share an SM
// (4 x 52 KB; with full 9x9 matrices it was three, and the 547 six-point warps of a 16k-crop batch needed two waves).
constexpr int SOLVER_V = 0;                    // 81: eigenvectors (rows)
constexpr int SOLVER_A = 81;                   // 36: Jacobi matrix, strict upper triangle: A[r][c], r < c, at tri_T(r) + c (destroyed)
constexpr int SOLVER_W = 117;                  // 9 : eigenvalues; on entry of sv_jacobi the diagonal of the matrix
constexpr int SOLVER_N = 126;                  // 45: LM normal matrix J^T J, upper triangle incl. diagonal (sym_idx)
constexpr int SOLVER_v = 171;                  // 9 : J^T r
constexpr int SOLVER_x = 180;                  // 9 : current parameters
constexpr int SOLVER_xd = 189;                 // 9 : trial parameters
constexpr int SOLVER_D = 198;                  // 9 : diag(J^T J) of the first iterate
constexpr int SOLVER_DOUBLES = 207;            // per lane
constexpr int SOLVER_SMEM_BYTES = SOLVER_DOUBLES * 32 * 8;      // 52,992 B per warp
constexpr int SOLVER_WARPS_PER_SM = 4;

// strict upper triangle of a 9x9 matrix, row-major: element (r, c), r < c, lives at tri_T(r) + c.  (r == c == 0 maps to -1:
// SOLVER_A is not the first block, so the unconditional loads of the branch-free rotation stay inside the allocation.)
__host__ __device__ constexpr int tri_T(int r) { return (r * (15 - r)) / 2 - 1; }
__device__ __forceinline__ int tri_Td(int r) { return (int)((unsigned)(r * (15 - r)) >> 1) - 1; }       // run-time r in 0..8
// upper triangle incl. diagonal of a symmetric 9x9 matrix
__host__ __device__ constexpr int sym_idx(int i, int j) { return i <= j ? i * 9 - (i * (i - 1)) / 2 + (j - i) : j * 9 - (j * (j - 1)) / 2 + (i - j); }

struct LaneMem {
    double *p;                                  // already offset by the lane
    __device__ __forceinline__ double &operator()(int base, int e) const { return p[(base + e) * 32]; }
};

// OpenCV's hypot: (a > b) ? a * sqrt(1 + (b/a)^2) : (b > 0 ? b * sqrt(1 + (a/b)^2) : 0) on the magnitudes -- written
// branch-free (larger / smaller magnitude; a == b takes the same values either way), the operations are the same
__device__ __forceinline__ double sv_hypot(double a, double b) {
    a = fabs(a); b = fabs(b);
    const double mx = a > b ? a : b, mn = a > b ? b : a;
    const double r = mn / (mx > 0 ? mx : 1.0);
    return mx > 0 ? mx * sqrt(1 + r * r) : 0.0;
}

__device__ __forceinline__ int nib_get(unsigned long long pk, int i) { return (int)((pk >> (4 * i)) & 15ull); }
__device__ __forceinline__ unsigned long long nib_set(unsigned long long pk, int i, int v) {
    return (pk & ~(15ull << (4 * i))) | ((unsigned long long)v << (4 * i));
}
__device__ __forceinline__ int nib_get32(unsigned pk, int i) { return (int)((pk >> (4 * i)) & 15u); }
__device__ __forceinline__ unsigned nib_set32(unsigned pk, int i, int v) { return (pk & ~(15u << (4 * i))) | ((unsigned)v << (4 * i)); }

// "first strict maximum" of JacobiImpl_'s scans (m = first candidate; a later one replaces it only if strictly
// greater) as a balanced tree: of two candidates in scan order the later wins only if strictly greater.  Candidates
// that are out of range carry -1 (every real |a| is >= 0).  Depth 4 instead of a 9- or 16-long dependent chain.
struct Cand { double v; int pos; };
__device__ __forceinline__ Cand cand_first_max(const Cand &a, const Cand &b) { return b.v > a.v ? b : a; }   // a precedes b
__device__ __forceinline__ int first_max9(const double *v) {
    Cand c[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { c[i].v = v[i]; c[i].pos = i; }
    const Cand c01 = cand_first_max(c[0], c[1]), c23 = cand_first_max(c[2], c[3]), c45 = cand_first_max(c[4], c[5]), c67 = cand_first_max(c[6], c[7]);
    const Cand c03 = cand_first_max(c01, c23), c47 = cand_first_max(c45, c67);
    return cand_first_max(cand_first_max(c03, c47), c[8]).pos;
}
__device__ __forceinline__ int first_max16(const double *v) {
    Cand c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i].v = v[i]; c[i].pos = i; }
#pragma unroll
    for (int w = 1; w < 16; w *= 2)
#pragma unroll
        for (int i = 0; i < 16; i += 2 * w) c[i] = cand_first_max(c[i], c[i + w]);
    return c[0].pos;
}

// index of the first maximum of |A[k][i]|, i = k+1..8 (JacobiImpl_'s indR[k]); k <= 7 -- from memory (initial pass)
__device__ __forceinline__ int sv_row_max(const LaneMem &m, int k) {
    double v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = i > k ? fabs(m(SOLVER_A, tri_T(k) + i)) : -1.0;
    return first_max9(v);
}
// index of the first maximum of |A[i][k]|, i = 0..k-1 (indC[k]); k >= 1
__device__ __forceinline__ int sv_col_max(const LaneMem &m, int k) {
    double v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = i < k ? fabs(m(SOLVER_A, tri_T(i) + k)) : -1.0;
    return first_max9(v);
}

// cv::eigen (JacobiImpl_, n = 9) on the matrix whose strict upper triangle is in A and whose diagonal is in W
// -> W (sorted descending), V (rows, NOT permuted), perm = row of V holding
// eigenvector i (nibble-packed).  `active`: lanes without a job skip the sweep but take part in the warp votes.
// The dependent chain of one rotation is what bounds a solve (~1150 rotations in sequence), so: the pivot search and
// the four index scans are depth-4 trees, and the scans run on the freshly rotated values still in registers (the
// elements of rows / columns k and l ARE the rotation's outputs) instead of re-reading them from shared memory.
__device__ inline unsigned long long sv_jacobi(const LaneMem &m, bool active) {
    constexpr int n = 9;
    unsigned indR = 0, indC = 0;           // indR: nibble k = indR[k], k = 0..7;  indC: nibble i-1 = indC[i], i = 1..8
    if (active) {
#pragma unroll
        for (int e = 0; e < n * n; ++e) m(SOLVER_V, e) = (e / n == e % n) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < n; ++k) {
            if (k < n - 1) indR = nib_set32(indR, k, sv_row_max(m, k));
            if (k > 0) indC = nib_set32(indC, k - 1, sv_col_max(m, k));
        }
    }
    const int maxIters = n * n * 30;
    for (int iters = 0; iters < maxIters; ++iters) {
        int k = 0, l = 1;
        double p = 0, Wk = 0, Wl = 0;
        if (active) {
            // pivot: first strict maximum over [row candidates k = 0..7, then column candidates i = 1..8]
            double cv[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) cv[i] = fabs(m(SOLVER_A, tri_T(i) + nib_get32(indR, i)));
#pragma unroll
            for (int i = 1; i < 9; ++i) cv[7 + i] = fabs(m(SOLVER_A, tri_Td(nib_get32(indC, i - 1)) + i));
            const int pos = first_max16(cv);
            if (pos < 8) { k = pos; l = nib_get32(indR, pos); }
            else { l = pos - 7; k = nib_get32(indC, l - 1); }
            p = m(SOLVER_A, tri_Td(k) + l);
            Wk = m(SOLVER_W, k); Wl = m(SOLVER_W, l);      // (issued with the pivot load, ahead of the vote)
            if (fabs(p) <= DBL_EPSILON) active = false;
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            const double y = (Wl - Wk) * 0.5;
            double t = fabs(y) + sv_hypot(p, y);
            double s = sv_hypot(p, t);
            const double c = t / s;
            s = p / s; t = (p / t) * p;
            if (y < 0) { s = -s; t = -t; }
            const int Tk = tri_Td(k), Tl = tri_Td(l);
            m(SOLVER_A, Tk + l) = 0;
            m(SOLVER_W, k) = Wk - t;
            m(SOLVER_W, l) = Wl + t;
            // rotate rows/columns k and l of the (upper-triangular) matrix: for every i != k, l the pair is
            // (S[i][k], S[i][l]) of the symmetric matrix S, i.e. A[min][max] -- JacobiImpl_'s three loops in one.
            // r0[i] = |new S[i][k]|, r1[i] = |new S[i][l]| stay in registers for the index scans.
            // (all 36 operands are loaded before the first store: the compiler cannot move a shared-memory load above a store
            // that might alias it, and a load-rotate-store sequence per pair exposes the load latency 18 times per rotation)
            double r0[n], r1[n], a0[n], b0[n], va[n], vb[n];
            int e0[n], e1[n];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                e0[i] = i < k ? tri_T(i) + k : Tk + i;      // (i == k / i == l: some other valid slot, loaded but never stored)
                e1[i] = i < l ? tri_T(i) + l : Tl + i;
                a0[i] = m(SOLVER_A, e0[i]); b0[i] = m(SOLVER_A, e1[i]);
                va[i] = m(SOLVER_V, n * k + i); vb[i] = m(SOLVER_V, n * l + i);
            }
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const bool on = i != k && i != l;
                const double n0 = a0[i] * c - b0[i] * s, n1 = a0[i] * s + b0[i] * c;
                if (on) { m(SOLVER_A, e0[i]) = n0; m(SOLVER_A, e1[i]) = n1; }
                r0[i] = fabs(n0); r1[i] = fabs(n1);
                m(SOLVER_V, n * k + i) = va[i] * c - vb[i] * s;
                m(SOLVER_V, n * l + i) = va[i] * s + vb[i] * c;
            }
            // indR[k]: i > k of S[k][i] (S[k][l] is now 0);  indC[k]: i < k of S[i][k];
            // indR[l]: i > l of S[l][i];                      indC[l]: i < l of S[i][l] (S[k][l] = 0)
            double sc[n];
            if (k < n - 1) {
#pragma unroll
                for (int i = 0; i < n; ++i) sc[i] = i > k ? (i == l ? 0.0 : r0[i]) : -1.0;
                indR = nib_set32(indR, k, first_max9(sc));
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) sc[i] = i < k ? r0[i] : -1.0;
                indC = nib_set32(indC, k - 1, first_max9(sc));
            }
            if (l < n - 1) {
#pragma unroll
                for (int i = 0; i < n; ++i) sc[i] = i > l ? r1[i] : -1.0;
                indR = nib_set32(indR, l, first_max9(sc));
            }
            {
#pragma unroll
                for (int i = 0; i < n; ++i) sc[i] = i < l ? (i == k ? 0.0 : r1[i]) : -1.0;
                indC = nib_set32(indC, l - 1, first_max9(sc));          // l >= 1 always
            }
        }
    }
    // OpenCV's descending selection sort of the eigenvalues, tracked as a row permutation
    unsigned long long perm = 0x876543210ull;
#pragma unroll
    for (int k = 0; k < n - 1; ++k) {
        int mi = k;
        double wm = m(SOLVER_W, k);
#pragma unroll
        for (int i = k + 1; i < n; ++i) {
            const double wi = m(SOLVER_W, i);
            if (wm < wi) { wm = wi; mi = i; }
        }
        if (mi != k) {
            const double wk = m(SOLVER_W, k);
            m(SOLVER_W, mi) = wk;
            m(SOLVER_W, k) = wm;
            const int pk = nib_get(perm, k), pm = nib_get(perm, mi);
            perm = nib_set(nib_set(perm, k, pm), mi, pk);
        }
    }
    return perm;
}

// SVBkSb after the Jacobi factors: x = sum_i (v_i . b / w_i) v_i over the eigenvalues above OpenCV's cut
__device__ __forceinline__ void sv_backsubst(const LaneMem &m, unsigned long long perm, const double *b, double *x) {
    constexpr int n = 9;
    double threshold = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) threshold += m(SOLVER_W, i);
    threshold *= DBL_EPSILON * 2;
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        double wi = m(SOLVER_W, i);
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        const int row = n * nib_get(perm, i);
        double vr[n];
        double acc = 0;
#pragma unroll
        for (int j = 0; j < n; ++j) { vr[j] = m(SOLVER_V, row + j); acc += vr[j] * b[j]; }
        acc *= wi;
#pragma unroll
        for (int j = 0; j < n; ++j) x[j] = x[j] + acc * vr[j];
    }
}

// max_c |(A^-1)_cc| of cv::invert(A, DECOMP_EIG), seeded with DBL_EPSILON (LMSolver's lambda re-initialisation)
__device__ __forceinline__ double sv_inverse_diag_max(const LaneMem &m, unsigned long long perm) {
    constexpr int n = 9;
    double threshold = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) threshold += m(SOLVER_W, i);
    threshold *= DBL_EPSILON * 2;
    double dg[n];
#pragma unroll
    for (int j = 0; j < n; ++j) dg[j] = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        double wi = m(SOLVER_W, i);
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        const int row = n * nib_get(perm, i);
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const double vic = m(SOLVER_V, row + j);
            dg[j] = dg[j] + (vic * wi) * vic;
        }
    }
    double mx = DBL_EPSILON;
#pragma unroll
    for (int j = 0; j < n; ++j) mx = fmax(mx, fabs(dg[j]));
    return mx;
}

// cv::Mat::dot (CV_64F, 9 elements) as the FMA-contracted AVX2 dispatch of the 4.13.0 wheel evaluates it
__device__ __forceinline__ double sv_dot9(const double *a, const double *b) {
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

struct PointSet {
    float Mx[6], My[6], mx[6], my[6];           // source (M) and destination (m) points as OpenCV's Point2f
    int count;
};

// HomographyRefineCallback::compute at parameters h: residuals r[2*count]; with `normal` also N = J^T J
// (cv::mulTransposed: every entry summed over the rows in order) and v = J^T r (cv::gemm: four interleaved
// accumulators over the rows) into shared memory.  Returns |r|^2 (cv::norm NORM_L2SQR) and max|r|.
__device__ __forceinline__ double sv_residual(const LaneMem &m, const PointSet &ps, const double *h, bool normal, double &rinf) {
    constexpr int lx = 9;
    double S = 0;
    rinf = 0;
    double vacc[lx][4];
    if (normal) {
#pragma unroll
        for (int e = 0; e < 45; ++e) m(SOLVER_N, e) = 0;
#pragma unroll
        for (int i = 0; i < lx; ++i) vacc[i][0] = vacc[i][1] = vacc[i][2] = vacc[i][3] = 0;
    }
#pragma unroll
    for (int pt = 0; pt < 6; ++pt) {
        if (pt < ps.count) {
            const double Mx = ps.Mx[pt], My = ps.My[pt];
            double ww = h[6] * Mx + h[7] * My + h[8];
            ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
            const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
            const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
            const double r0 = xi - ps.mx[pt], r1 = yi - ps.my[pt];
            S += r0 * r0;
            S += r1 * r1;
            rinf = fmax(rinf, fmax(fabs(r0), fabs(r1)));
            if (normal) {
                const double J0[lx] = {Mx * ww, My * ww, ww, 0., 0., 0., -Mx * ww * xi, -My * ww * xi, -ww * xi};
                const double J1[lx] = {0., 0., 0., Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi, -ww * yi};
#pragma unroll
                for (int i = 0; i < lx; ++i)
#pragma unroll
                    for (int j = i; j < lx; ++j) {
                        double a = m(SOLVER_N, sym_idx(i, j));
                        a += J0[i] * J0[j];
                        a += J1[i] * J1[j];
                        m(SOLVER_N, sym_idx(i, j)) = a;
                    }
                // rows 2*pt and 2*pt+1 of J go to accumulator (row & 3); the tail rows of a 10-row system to accumulator 0
#pragma unroll
                for (int i = 0; i < lx; ++i) {
                    vacc[i][(2 * pt) & 3] += J0[i] * r0;
                    if (ps.count == 5 && pt == 4) vacc[i][0] += J1[i] * r1;          // 10 rows: rows 8, 9 are gemm's tail
                    else vacc[i][(2 * pt + 1) & 3] += J1[i] * r1;
                }
            }
        }
    }
    if (normal) {
#pragma unroll
        for (int i = 0; i < lx; ++i) m(SOLVER_v, i) = ((vacc[i][0] + vacc[i][1]) + vacc[i][2]) + vacc[i][3];     // (N's lower triangle is the mirror: sym_idx)
    }
    return S;
}

// OpenCV's "returns None" test: all source or all destination points share an x or a y
__device__ __forceinline__ bool sv_degenerate_or_setup(const PointSet &ps, double &cMx, double &cMy, double &cmx, double &cmy, double &sMx,
                                                       double &sMy, double &smx, double &smy) {
    const int count = ps.count;
    cMx = cMy = cmx = cmy = sMx = sMy = smx = smy = 0;
    for (int i = 0; i < count; ++i) {
        cmx += ps.mx[i]; cmy += ps.my[i];
        cMx += ps.Mx[i]; cMy += ps.My[i];
    }
    cmx /= count; cmy /= count; cMx /= count; cMy /= count;
    for (int i = 0; i < count; ++i) {
        smx += fabs(ps.mx[i] - cmx); smy += fabs(ps.my[i] - cmy);
        sMx += fabs(ps.Mx[i] - cMx); sMy += fabs(ps.My[i] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) return true;
    smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
    return false;
}

// The whole solve for this lane's point set (ps.count in {4, 5, 6}; 0 = no task).  Must be called by all 32 lanes
// of the warp.  Returns false where OpenCV returns None (H untouched); H[8] == 1 otherwise.
__device__ inline bool sv_find_homography(const LaneMem &m, const PointSet &ps, double *H) {
    constexpr int lx = 9;
    enum { DONE = 0, DLT = 1, SOLVE = 2, INVERT = 3 };
    int phase = DONE;
    bool good = false;
    double cMx, cMy, cmx, cmy, sMx, sMy, smx, smy;
    if (ps.count >= 4 && !sv_degenerate_or_setup(ps, cMx, cMy, cmx, cmy, sMx, sMy, smx, smy)) {
        good = true;
        phase = DLT;
        // L^T L, upper triangle accumulated point by point (fundam.cpp runKernel), then mirrored
#pragma unroll
        for (int e = 0; e < 36; ++e) m(SOLVER_A, e) = 0;
#pragma unroll
        for (int e = 0; e < 9; ++e) m(SOLVER_W, e) = 0;
#pragma unroll
        for (int pt = 0; pt < 6; ++pt) {
            if (pt < ps.count) {
                const double x = (ps.mx[pt] - cmx) * smx, y = (ps.my[pt] - cmy) * smy;
                const double X = (ps.Mx[pt] - cMx) * sMx, Y = (ps.My[pt] - cMy) * sMy;
                const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
                const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
#pragma unroll
                for (int j = 0; j < 9; ++j)
#pragma unroll
                    for (int k = j; k < 9; ++k) {
                        if (k == j) m(SOLVER_W, j) += Lx[j] * Lx[k] + Ly[j] * Ly[k];
                        else m(SOLVER_A, tri_T(j) + k) += Lx[j] * Lx[k] + Ly[j] * Ly[k];
                    }
            }
        }
    }
    // LM registers that persist across Jacobi jobs
    double S = 0, lambda = 1, lc = 0.75, Sd = 0, nu = 0, nd = 0, rinf = 0;
    int iter = 0;
    for (;;) {
        if (!__any_sync(0xffffffffu, phase != DONE)) break;
        const unsigned long long perm = sv_jacobi(m, phase != DONE);
        if (phase == DONE) continue;
        bool finish = false;               // run the tail of the LM iteration (lambda update done, accept / stop test)
        if (phase == DLT) {
            const int row = 9 * nib_get(perm, 8);           // eigenvector of the smallest eigenvalue
            double H0[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) H0[e] = m(SOLVER_V, row + e);
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
            double x0[lx];
#pragma unroll
            for (int i = 0; i < 9; ++i) { x0[i] = H1[i] * scl; m(SOLVER_x, i) = x0[i]; }
            if (ps.count == 4) {
                phase = DONE;
            } else {
                // LMSolverImpl::run prologue
                S = sv_residual(m, ps, x0, true, rinf);
#pragma unroll
                for (int i = 0; i < lx; ++i) m(SOLVER_D, i) = m(SOLVER_N, sym_idx(i, i));
                lambda = 1; lc = 0.75; iter = 0;
                phase = SOLVE;
            }
        } else if (phase == SOLVE) {
            double vv[lx], d[lx], xd[lx], td[lx];
#pragma unroll
            for (int i = 0; i < lx; ++i) vv[i] = m(SOLVER_v, i);
            sv_backsubst(m, perm, vv, d);
            nd = 0;
#pragma unroll
            for (int i = 0; i < lx; ++i) {
                xd[i] = m(SOLVER_x, i) - d[i];
                m(SOLVER_xd, i) = xd[i];
                nd = fmax(nd, fabs(d[i]));
            }
            double dummy;
            Sd = sv_residual(m, ps, xd, false, dummy);
            // temp_d = -N d + 2 v (cv::gemm(A, d, -1, v, 2): four interleaved accumulators, tail into the first)
#pragma unroll
            for (int i = 0; i < lx; ++i) {
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
                for (int k = 0; k < 8; k += 4) {
                    s0 += m(SOLVER_N, sym_idx(i, k)) * d[k];
                    s1 += m(SOLVER_N, sym_idx(i, k + 1)) * d[k + 1];
                    s2 += m(SOLVER_N, sym_idx(i, k + 2)) * d[k + 2];
                    s3 += m(SOLVER_N, sym_idx(i, k + 3)) * d[k + 3];
                }
                s0 += m(SOLVER_N, sym_idx(i, 8)) * d[8];
                td[i] = -1. * (((s0 + s1) + s2) + s3) + 2. * vv[i];
            }
            const double dS = sv_dot9(d, td);
            const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
            finish = true;
            if (R > 0.75) {
                lambda *= 0.5;
                if (lambda < lc) lambda = 0;
            } else if (R < 0.25) {
                const double t = sv_dot9(d, vv);
                nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
                nu = fmin(fmax(nu, 2.), 10.);
                if (lambda == 0) {
                    // needs invert(A, DECOMP_EIG): one more Jacobi job on N itself, then the tail below
#pragma unroll
                    for (int i = 0; i < lx; ++i)
#pragma unroll
                        for (int j = i; j < lx; ++j) {
                            const double a = m(SOLVER_N, sym_idx(i, j));
                            if (i == j) m(SOLVER_W, i) = a;
                            else m(SOLVER_A, tri_T(i) + j) = a;
                        }
                    phase = INVERT;
                    finish = false;
                } else {
                    lambda *= nu;
                }
            }
        } else {   // INVERT
            const double maxval = sv_inverse_diag_max(m, perm);
            lambda = lc = 1. / maxval;
            nu *= 0.5;
            lambda *= nu;
            finish = true;
        }
        if (finish) {
            if (Sd < S) {
                S = Sd;
                double xn[lx];
#pragma unroll
                for (int i = 0; i < lx; ++i) { xn[i] = m(SOLVER_xd, i); m(SOLVER_x, i) = xn[i]; }
                sv_residual(m, ps, xn, true, rinf);
            }
            iter++;
            const bool proceed = iter < 10 && nd >= (double)FLT_EPSILON && rinf >= (double)FLT_EPSILON;
            phase = proceed ? SOLVE : DONE;
        }
        if (phase == SOLVE) {
            // Ap = N + lambda * D for the next solve
#pragma unroll
            for (int i = 0; i < lx; ++i)
#pragma unroll
                for (int j = i; j < lx; ++j) {
                    const double a = m(SOLVER_N, sym_idx(i, j));
                    if (i == j) m(SOLVER_W, i) = a + lambda * m(SOLVER_D, i);
                    else m(SOLVER_A, tri_T(i) + j) = a;
                }
        }
    }
    if (good) {
        double x[lx];
#pragma unroll
        for (int i = 0; i < lx; ++i) x[i] = m(SOLVER_x, i);
        if (ps.count > 4) {
            // H.convertTo(H, H.type(), scaleFor(H(2,2)))
            const double sc2 = fabs(x[8]) > DBL_EPSILON ? 1. / x[8] : 1.;
#pragma unroll
            for (int i = 0; i < lx; ++i) x[i] = x[i] * sc2;
        }
#pragma unroll
        for (int i = 0; i < lx; ++i) H[i] = x[i];
    }
    return good;
}

}  // namespace fusg
