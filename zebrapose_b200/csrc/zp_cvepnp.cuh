// Exact replay of the arithmetic of cv2.solvePnP(SOLVEPNP_EPNP) for minimal samples (OpenCV is the un-vendored library
// behind the reference's cv2.solvePnPRansac call, /root/reference/zebrapose/binary_code_helper/CNN_output_to_pose.py:155-157).
//
// Why bit-exact: for a 5-point sample M^T M (12x12) has a 2-dimensional null space; which basis of it a solver returns is
// decided by rounding, and one of EPnP's three beta initialisations depends on that basis.  A solver that is merely
// accurate picks another basis, hence another pose, on a large fraction of outlier-bearing samples, and the RANSAC
// winner then differs from cv2's on 6-22 % of crops (round 1).  So this file performs the same IEEE-754 double
// operations, in the same order, as OpenCV's calib3d EPnP + its small-matrix one-sided Jacobi SVD (published algorithms:
// Lepetit/Moreno-Noguer/Fua 2009; Hestenes 1958), pinned by oracle/cv_epnp.c against cv2 itself:
//   * no fused multiply-add anywhere (this header must be compiled with nvcc -fmad=false / gcc -ffp-contract=off);
//   * sums run sequentially in index order; the Jacobi's hypot is |a|*sqrt(1+(b/a)^2); division and sqrt are IEEE;
//   * image points pass through float32 normalised coordinates: u' = double(float((u-cx)*(1/fx)))*fx + cx.
//
// What is NOT the same is the schedule, because the FP64 pipe of an SM (16 lanes per scheduler) is the resource this
// work is bound by and a warp instruction costs the same whether 1 or 32 of its lanes do something useful:
//   stage A  "prep"  one THREAD per hypothesis: staging, centroid, 3x3 PCA, control points, their SVD inverse, the
//                    barycentric coordinates (all short serial chains)
//   stage B  "null"  SIX LANES per hypothesis: the 78 sums of M^T M, 13 per lane, then the 12x12 Jacobi SVD.  The cyclic order visits the pairs (0,1),(0,2)..
//                    (10,11) one after the other, but a pair only depends on the last earlier pairs that touched its two
//                    rows, so pair (i,j) of sweep s runs at step T = 12 s + i + j: six pairs per step, a sweep every 12
//                    steps, and a result bit-identical to the serial order because every rotation sees exactly the
//                    operands it would have seen (tests/test_oracle_cv_epnp.py).  The squared row norms the serial code
//                    carries along are sequential sums of the stored row, so a lane recomputes them from the rows it loads.
//                    Then L (6x10) and rho, a row per lane.
//   stage C  "cand"  one THREAD per (beta initialisation, hypothesis): the 6xN least-squares problem through the same
//                    Jacobi SVD (serial order), 5 Gauss-Newton steps, camera-frame points, 3x3 alignment SVD, error
//   stage D  "pick"  EPnP's choice among the three candidates
// Arrays of stages A and C are strided views (element k at p[k * s]): thread-private data is interleaved over the threads
// of a CTA (conflict-free shared memory) and the hand-off between the stages is structure-of-arrays in global memory
// (coalesced), with one piece of code for both and for the host harness (tests/native/cvepnp_host.cpp, stride 1).
#pragma once
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifndef ZP_HD
#ifdef __CUDACC__
#define ZP_HD __host__ __device__
#else
#define ZP_HD
#endif
#endif

#define CVE_G 6              // lanes per hypothesis in stage B
#define CVE_MAXM 8           // largest minimal-sample size
#define CVE_RS 14            // row stride (doubles) of the 12x12 matrix in stage B: even, so that rows are 16-byte aligned (128-bit shared-memory accesses)

struct CveCam { double fu, fv, uc, vc; };

// strided view of a small array
struct Dv {
    double* p;
    int s;
    ZP_HD double& operator[](int k) const { return p[(size_t)k * s]; }
    ZP_HD Dv at(int k) const { Dv v; v.p = p + (size_t)k * s; v.s = s; return v; }
};
ZP_HD inline Dv cve_dv(double* p, int s) { Dv v; v.p = p; v.s = s; return v; }

// hand-off record of one hypothesis between the stages (offsets in doubles; stored as [field][hypothesis])
#define CVH_PW 0             // [8][3]
#define CVH_US 24            // [8][2]
#define CVH_AL 40            // [8][4]
#define CVH_CW 72            // [4][3]
#define CVH_V4 84            // [4][12] rows 11,10,9,8 of U^T
#define CVH_L 132            // [6][10]
#define CVH_RHO 192          // [6]
#define CVH_OUT 198          // [3][13] per candidate: R[9] t[3] err
#define CVH_DOUBLES 237

ZP_HD inline double cve_hypot(double a, double b) {
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// one Jacobi pair: rows Ai, Aj of length M (and rows Vi, Vj of length n of the accumulated rotations)
// ------------------------------------------------------------------------------------------------------------------
// The library's control flow (skip test, two-branch hypot, two-branch c/s) is evaluated here WITHOUT divergent branches:
// the lanes of a warp work on different pairs / hypotheses, and a divergent branch around a division + square root makes
// the warp pay both sides.  Every lane still performs exactly the operations of its own branch:
//   hypot(p, beta)  = big * sqrt(1 + (small/big)^2) with big/small = max/min(|p|, |beta|)  (0 when both are 0)
//   beta < 0:  s = sqrt(((gamma-beta)*0.5) / gamma), c = p / (gamma*s*2);  else c = sqrt((gamma+beta) / (gamma*2)), s = p / (gamma*c*2)
//   skip test |p| <= eps*sqrt(a*b): decided from p*p against eps^2*(a*b) when the two differ by more than 1e-9 relative
//   (the rounding of the exact expression is 4 ulp at most); the square root is only taken in between or near underflow.
// the arithmetic of one pair on rows held in registers: returns false (rows untouched) when the pair is skipped, else the
// rotation factors in c, s and the rotated rows in ri, rj
template <int M>
ZP_HD inline bool cve_pair_core(double* ri, double* rj, double& c, double& s) {
    const double eps = DBL_EPSILON * 10;
    double p = 0, a = 0, b = 0;
#pragma unroll
    for (int k = 0; k < M; k++) { p += ri[k] * rj[k]; a += ri[k] * ri[k]; b += rj[k] * rj[k]; }
    {
        const double ab = a * b, pp = p * p, lim = ab * (eps * eps);
        bool skip;
        if (lim > 1e-280 && ab < 1e300 && pp < 1e300 && pp > lim * (1 + 1e-9)) skip = false;
        else if (lim > 1e-280 && ab < 1e300 && pp < lim * (1 - 1e-9)) skip = true;
        else skip = fabs(p) <= eps * sqrt(ab);
        if (skip) return false;
    }
    p *= 2;
    const double beta = a - b;
    double gamma;
    {
        const double x = fabs(p), y = fabs(beta);
        const bool xg = x > y;
        const double big = xg ? x : y, small = xg ? y : x;
        const double q = small / big;
        const double r = big * sqrt(1 + q * q);
        gamma = (xg || y > 0) ? r : 0.0;
    }
    const bool neg = beta < 0;
    const double num = neg ? (gamma - beta) * 0.5 : gamma + beta;
    const double den = neg ? gamma : gamma * 2;
    const double first = sqrt(num / den);
    const double second = p / (gamma * first * 2);
    c = neg ? second : first; s = neg ? first : second;
#pragma unroll
    for (int k = 0; k < M; k++) {
        const double t0 = c * ri[k] + s * rj[k];
        const double t1 = -s * ri[k] + c * rj[k];
        ri[k] = t0; rj[k] = t1;
    }
    return true;
}

template <int M, bool HASV>
ZP_HD inline bool cve_pair(Dv Ai, Dv Aj, Dv Vi, Dv Vj, int n) {
    double ri[M], rj[M], c, s;
#pragma unroll
    for (int k = 0; k < M; k++) { ri[k] = Ai[k]; rj[k] = Aj[k]; }
    if (!cve_pair_core<M>(ri, rj, c, s)) return false;
#pragma unroll
    for (int k = 0; k < M; k++) { Ai[k] = ri[k]; Aj[k] = rj[k]; }
    if (HASV) {
        for (int k = 0; k < n; k++) {
            const double t0 = c * Vi[k] + s * Vj[k];
            const double t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0; Vj[k] = t1;
        }
    }
    return true;
}

// the same for two contiguous, 16-byte aligned rows (stage B): on the device the rows move as 128-bit shared-memory accesses
// (half the load / store instructions of the pair step, whose shared-memory queue was a top stall: ncu mio_throttle)
template <int M>
ZP_HD inline bool cve_pair_rows(double* Ai, double* Aj) {
    static_assert(M % 2 == 0, "rows are moved two doubles at a time");
    double ri[M], rj[M], c, s;
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int k = 0; k < M; k += 2) {
        const double2 u = *reinterpret_cast<const double2*>(Ai + k), v = *reinterpret_cast<const double2*>(Aj + k);
        ri[k] = u.x; ri[k + 1] = u.y; rj[k] = v.x; rj[k + 1] = v.y;
    }
#else
    for (int k = 0; k < M; k++) { ri[k] = Ai[k]; rj[k] = Aj[k]; }
#endif
    if (!cve_pair_core<M>(ri, rj, c, s)) return false;
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int k = 0; k < M; k += 2) {
        *reinterpret_cast<double2*>(Ai + k) = make_double2(ri[k], ri[k + 1]);
        *reinterpret_cast<double2*>(Aj + k) = make_double2(rj[k], rj[k + 1]);
    }
#else
    for (int k = 0; k < M; k++) { Ai[k] = ri[k]; Aj[k] = rj[k]; }
#endif
    return true;
}

// serial cyclic Jacobi (stages A and C): n rows of length M, row r of At at At.at(r * astep); Vt likewise
template <int M, bool HASV>
ZP_HD inline void cve_jserial(Dv At, int astep, Dv Vt, int vstep, int n) {
    if (HASV)
        for (int i = 0; i < n; i++)
            for (int k = 0; k < n; k++) Vt[i * vstep + k] = i == k ? 1.0 : 0.0;
    const int max_iter = M > 30 ? M : 30;
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++)
                changed |= cve_pair<M, HASV>(At.at(i * astep), At.at(j * astep), HASV ? Vt.at(i * vstep) : Vt,
                                             HASV ? Vt.at(j * vstep) : Vt, n);
        if (!changed) break;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// wave-front Jacobi (stage B): one lane's view of the problem, worked on by nl lanes of which this one is number l;
// chg = 4 ints shared by the problem's lanes ("sweep s rotated something", slot s & 3)
// ------------------------------------------------------------------------------------------------------------------
struct CveJ {
    double* At; int* chg;
    int astep, n, l, nl, max_iter;
    bool done;
};

ZP_HD inline CveJ cve_j_make(double* At, int astep, int n, int m_cols, int l, int nl, int* chg, bool active) {
    CveJ j;
    j.At = At; j.chg = chg; j.astep = astep; j.n = n; j.l = l; j.nl = nl;
    j.max_iter = m_cols > 30 ? m_cols : 30;
    j.done = !(active && l < nl);
    return j;
}

// cleared flags; called by lane l == 0 of the problem before the first step (then a barrier)
ZP_HD inline void cve_j_init(const CveJ& j) {
    if (j.done || j.l != 0) return;
    j.chg[0] = j.chg[1] = j.chg[2] = j.chg[3] = 0;
}

// step T (1, 2, ...): this lane's pair, if it has one.  Pairs of sweep s = (T - tau)/n with i + j = tau; two sweeps overlap.
template <int M, int N>
ZP_HD inline void cve_jstep_a(CveJ& j, int T) {
    constexpr int n = N;
    const int ta = (T - 1) % n + 1, sa = (T - ta) / n;
    const int tb = ta + n, sb = sa - 1;
    const int lo_a = ta - n + 1 > 0 ? ta - n + 1 : 0;
    int cnt_a = (ta + 1) / 2 - lo_a;
    if (cnt_a < 0 || sa >= j.max_iter) cnt_a = 0;
    const int lo_b = tb - n + 1;
    int cnt_b = (tb + 1) / 2 - lo_b;
    if (cnt_b < 0 || sb < 0 || tb > 2 * n - 3) cnt_b = 0;
    int ii, jj, s;
    if (j.l < cnt_a) { ii = lo_a + j.l; jj = ta - ii; s = sa; }
    else if (j.l - cnt_a < cnt_b) { ii = lo_b + (j.l - cnt_a); jj = tb - ii; s = sb; }
    else return;
    if (jj >= n || ii >= jj) return;
    if (cve_pair_rows<M>(j.At + ii * j.astep, j.At + jj * j.astep)) j.chg[s & 3] = 1;     // At 16-byte aligned, astep even
}

// after the barrier that follows step T: if a sweep completed at T, stop when it rotated nothing (or at the sweep cap)
template <int N>
ZP_HD inline void cve_jstep_c(CveJ& j, int T) {
    constexpr int n = N;
    const int num = T - (2 * n - 3);
    if (num < 0 || num % n != 0) return;
    const int s = num / n;
    const int c = j.chg[s & 3];
    if (j.l == 0) j.chg[(s + 2) & 3] = 0;
    if (!c || s + 1 >= j.max_iter) j.done = true;
}

// ------------------------------------------------------------------------------------------------------------------
// the tail of the SVD routine: singular values = row norms, selection sort (descending, rows swapped physically),
// rows normalised (an exactly-zero singular value gets a pseudo-random row orthogonalised against the rows above it).
// have_v = false: no Vt storage (the rows are still swapped: every call replayed here has a Vt in the library).
// ------------------------------------------------------------------------------------------------------------------
ZP_HD inline void cve_finish(Dv At, int astep, Dv W, Dv Vt, int vstep, bool have_v, int m, int n, int n1) {
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    int i, j, k;
    double sd;
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { const double t = At[i * astep + k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (k = 0; k < m; k++) { t = At[i * astep + k]; At[i * astep + k] = At[j * astep + k]; At[j * astep + k] = t; }
            if (have_v) for (k = 0; k < n; k++) { t = Vt[i * vstep + k]; Vt[i * vstep + k] = Vt[j * vstep + k]; Vt[j * vstep + k] = t; }
        }
    }
    uint64_t rng = 0x12345678;
    for (i = 0; i < n1; i++) {
        sd = i < n ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1. / m;
            for (k = 0; k < m; k++) {
                rng = (uint64_t)(uint32_t)rng * 4164903690ull + (uint32_t)(rng >> 32);
                At[i * astep + k] = ((uint32_t)rng & 256) != 0 ? val0 : -val0;
            }
            for (int it2 = 0; it2 < 2; it2++)
                for (j = 0; j < i; j++) {
                    sd = 0;
                    for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
                    double asum = 0;
                    for (k = 0; k < m; k++) {
                        const double t = At[i * astep + k] - sd * At[j * astep + k];
                        At[i * astep + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (k = 0; k < m; k++) At[i * astep + k] *= asum;
                }
            sd = 0;
            for (k = 0; k < m; k++) { const double t = At[i * astep + k]; sd += t * t; }
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1 / sd : 0.;
        for (k = 0; k < m; k++) At[i * astep + k] *= s;
    }
}

// x = pinv(A) b from the finished SVD of a 6 x NC system (At rows = left vectors, Vt): singular values at or below
// 2 eps sum(w) are dropped
ZP_HD inline void cve_backsubst6(Dv At, Dv w, Dv Vt, int nc, const double* b, double* x) {
    double thr = 0;
    for (int i = 0; i < nc; i++) { x[i] = 0; thr += w[i]; }
    thr *= DBL_EPSILON * 2;
    for (int i = 0; i < nc; i++) {
        double wi = w[i];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < 6; j++) s += At[i * 6 + j] * b[j];
        s *= wi;
        for (int j = 0; j < nc; j++) x[j] = x[j] + s * Vt[i * nc + j];
    }
}

// Householder least squares of the 6x4 Gauss-Newton system in the published EPnP code's evaluation order (the column
// scale is the largest magnitude among rows k .. nr-2: its scan stops one row early)
ZP_HD inline void cve_qr_solve64(double* A, double* b, double* X) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
#pragma unroll
    for (int k = 0; k < nc; k++) {
        double eta = fabs(A[k * nc + k]);
#pragma unroll
        for (int i = k + 1; i < nr; i++) {
            const double elt = fabs(A[(i - 1) * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) { A1[k] = A2[k] = 0.0; return; }
        double sum2 = 0.0;
        const double inv_eta = 1. / eta;
#pragma unroll
        for (int i = k; i < nr; i++) { A[i * nc + k] *= inv_eta; sum2 += A[i * nc + k] * A[i * nc + k]; }
        double sigma = sqrt(sum2);
        if (A[k * nc + k] < 0) sigma = -sigma;
        A[k * nc + k] += sigma;
        A1[k] = sigma * A[k * nc + k];
        A2[k] = -eta * sigma;
#pragma unroll
        for (int j = k + 1; j < nc; j++) {
            double sum = 0;
#pragma unroll
            for (int i = k; i < nr; i++) sum += A[i * nc + k] * A[i * nc + j];
            const double tau = sum / A1[k];
#pragma unroll
            for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
        }
    }
#pragma unroll
    for (int j = 0; j < nc; j++) {
        double tau = 0;
#pragma unroll
        for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
        tau /= A1[j];
#pragma unroll
        for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
#pragma unroll
    for (int i = nc - 2; i >= 0; i--) {
        double sum = 0;
#pragma unroll
        for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
        X[i] = (b[i] - sum) / A2[i];
    }
}

// 5 Gauss-Newton steps on the betas; L (6x10) and rho (6) are strided views (read-only)
ZP_HD inline void cve_gauss_newton(Dv L, Dv rho, double* betas) {
    double a[24], b[6], x[4] = {0, 0, 0, 0};
    for (int it = 0; it < 5; it++) {
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double rowL[10];
#pragma unroll
            for (int q = 0; q < 10; q++) rowL[q] = L[10 * i + q];
            double* rowA = a + i * 4;
            rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
            rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
            rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
            rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
            b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                             rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                             rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                             rowL[9] * betas[3] * betas[3]);
        }
        cve_qr_solve64(a, b, x);
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

ZP_HD inline double cve_dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ------------------------------------------------------------------------------------------------------------------
// stage A: everything up to the barycentric coordinates.  Views: pw [m][3], us [m][2], al [m][4], cw [4][3] (kept for the
// later stages), wk = 33 doubles of scratch (A3 9 | V3 9 | W 3 | CI 9 | spare 3).
// corr: the five float planes of the crop (u | v | X | Y | Z, `cap` apart); idx: the m sample indices (all valid)
// ------------------------------------------------------------------------------------------------------------------
ZP_HD inline void cve_stage_a(const float* corr, int cap, const int32_t* idx, int m, const CveCam& cam,
                              Dv pw, Dv us, Dv al, Dv cw, Dv wk) {
    const Dv A3 = wk, V3 = wk.at(9), W = wk.at(18), ci = wk.at(21);
    const double ifx = 1. / cam.fu, ify = 1. / cam.fv;
    for (int p = 0; p < m; p++) {
        const int i = idx[p];
        const double u = (double)corr[i], v = (double)corr[(size_t)cap + i];
        pw[3 * p] = (double)corr[2 * (size_t)cap + i];
        pw[3 * p + 1] = (double)corr[3 * (size_t)cap + i];
        pw[3 * p + 2] = (double)corr[4 * (size_t)cap + i];
        // undistortPoints (no distortion) -> float32 normalised coordinates -> back to pixels
        const double x = (u - cam.uc) * ifx, y = (v - cam.vc) * ify;
        us[2 * p] = (double)(float)x * cam.fu + cam.uc;
        us[2 * p + 1] = (double)(float)y * cam.fv + cam.vc;
    }
    // control points: centroid + PCA axes of the object points
    double c0[3] = {0, 0, 0};
    for (int p = 0; p < m; p++) for (int j = 0; j < 3; j++) c0[j] += pw[3 * p + j];
    for (int j = 0; j < 3; j++) { c0[j] /= m; cw[j] = c0[j]; }
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
            double s = 0;
            for (int p = 0; p < m; p++) s += (pw[3 * p + i] - c0[i]) * (pw[3 * p + j] - c0[j]);
            A3[i * 3 + j] = s; A3[j * 3 + i] = s;
        }
    cve_jserial<3, true>(A3, 3, V3, 3, 3);
    cve_finish(A3, 3, W, V3, 3, true, 3, 3, 3);
    for (int i = 1; i < 4; i++) {
        const double k = sqrt(W[i - 1] / m);
        for (int j = 0; j < 3; j++) cw[3 * i + j] = c0[j] + k * A3[3 * (i - 1) + j];
    }
    // inverse of the control-point basis through its SVD: cc[3*i + j-1] = cw[j][i] - cw[0][i]; At = cc^T
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A3[r * 3 + c] = cw[3 * (r + 1) + c] - c0[c];
    cve_jserial<3, true>(A3, 3, V3, 3, 3);
    cve_finish(A3, 3, W, V3, 3, true, 3, 3, 3);
    const double thr = (W[0] + W[1] + W[2]) * (DBL_EPSILON * 2);
    for (int i = 0; i < 9; i++) ci[i] = 0;
    for (int i = 0; i < 3; i++) {           // V diag(1/w) U^T, one singular value at a time
        double wi = W[i];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        double buf[3];
        for (int j = 0; j < 3; j++) buf[j] = A3[i * 3 + j] * wi;
        for (int r = 0; r < 3; r++) {
            const double s = V3[i * 3 + r];
            for (int j = 0; j < 3; j++) ci[r * 3 + j] = ci[r * 3 + j] + s * buf[j];
        }
    }
    // barycentric coordinates
    for (int p = 0; p < m; p++) {
        const double d0 = pw[3 * p] - c0[0], d1 = pw[3 * p + 1] - c0[1], d2 = pw[3 * p + 2] - c0[2];
        double a1 = ci[0] * d0 + ci[1] * d1 + ci[2] * d2;
        double a2 = ci[3] * d0 + ci[4] * d1 + ci[5] * d2;
        double a3 = ci[6] * d0 + ci[7] * d1 + ci[8] * d2;
        al[4 * p + 1] = a1; al[4 * p + 2] = a2; al[4 * p + 3] = a3;
        al[4 * p] = 1.0f - a1 - a2 - a3;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// stage B pieces (A = the hypothesis' [12][13] matrix in shared memory, unit stride)
// ------------------------------------------------------------------------------------------------------------------
// M^T M by the six lanes: entry (i, j >= i) = sequential sum over the 2m rows of M (per point the rows [a fu, 0, a (uc-u)]
// and [0, a fv, a (vc-v)]), 13 entries per lane.  al [m][4] and us [m][2] in shared memory, unit stride.
// Two steps with a group barrier between them: (1) the rows of M are tabulated once -- in the A region itself, which is
// free until the entries are written -- as X1[x][p][c] / X2[x][p][c] (x = column type 0..2, explicit zero planes), so an
// entry costs 4 loads + 2 multiplies + 2 additions per point and no selects; (2) entries accumulated in registers, then
// (after the barrier of the caller) written.  m > 6 does not fit the region and takes the direct form.
ZP_HD inline bool cve_b_mtm_tabulated(int m) { return 24 * m <= 12 * CVE_RS; }

ZP_HD inline void cve_b_mtm_table(double* A, int lane, const double* al, const double* us, int m, const CveCam& cam) {
    if (!cve_b_mtm_tabulated(m)) return;
    double* X1 = A; double* X2 = A + 12 * m;
    for (int t = lane; t < 4 * m; t += CVE_G) {
        const int p = t >> 2;
        const double a = al[t];
        const double du = cam.uc - us[2 * p], dv = cam.vc - us[2 * p + 1];
        X1[t] = a * cam.fu; X1[4 * m + t] = 0.0; X1[8 * m + t] = a * du;
        X2[t] = 0.0; X2[4 * m + t] = a * cam.fv; X2[8 * m + t] = a * dv;
    }
}

// returns this lane's 13 sums in out[13] (entry e = lane + 6 t)
ZP_HD inline void cve_b_mtm_sums(const double* A, int lane, const double* al, const double* us, int m, const CveCam& cam,
                                 double* out) {
    const bool tab = cve_b_mtm_tabulated(m);
    const double* X1 = A; const double* X2 = A + 12 * m;
#pragma unroll
    for (int t = 0; t < 13; t++) {
        const int e = lane + CVE_G * t;
        int i = 0, r = e;
        while (r >= 12 - i) { r -= 12 - i; i++; }
        const int j = i + r;
        const int ci = i / 3, xi = i - 3 * ci, cj = j / 3, xj = j - 3 * cj;
        double s = 0;
        if (tab) {
            const double *p1i = X1 + 4 * m * xi + ci, *p1j = X1 + 4 * m * xj + cj, *p2i = X2 + 4 * m * xi + ci, *p2j = X2 + 4 * m * xj + cj;
            for (int p = 0; p < m; p++) {
                s += p1i[4 * p] * p1j[4 * p];
                s += p2i[4 * p] * p2j[4 * p];
            }
        } else {
            for (int p = 0; p < m; p++) {
                const double ai = al[4 * p + ci], aj = al[4 * p + cj];
                const double du = cam.uc - us[2 * p], dv = cam.vc - us[2 * p + 1];
                const double r1i = xi == 0 ? ai * cam.fu : xi == 1 ? 0.0 : ai * du;
                const double r1j = xj == 0 ? aj * cam.fu : xj == 1 ? 0.0 : aj * du;
                const double r2i = xi == 0 ? 0.0 : xi == 1 ? ai * cam.fv : ai * dv;
                const double r2j = xj == 0 ? 0.0 : xj == 1 ? aj * cam.fv : aj * dv;
                s += r1i * r1j;
                s += r2i * r2j;
            }
        }
        out[t] = s;
    }
}

ZP_HD inline void cve_b_mtm_store(double* A, int lane, const double* in) {
#pragma unroll
    for (int t = 0; t < 13; t++) {
        const int e = lane + CVE_G * t;
        int i = 0, r = e;
        while (r >= 12 - i) { r -= 12 - i; i++; }
        const int j = i + r;
        A[i * CVE_RS + j] = in[t]; A[j * CVE_RS + i] = in[t];
    }
}

// the tail of the 12x12 SVD, of which only rows 11, 10, 9, 8 of the sorted, normalised U^T are needed:
//   (1) lane l: singular values of rows l and l + 6 (sequential sums) -> W          [group barrier]
//   (2) every lane: the library's selection sort replayed on (W, index) in registers; V4 row q = row perm[11-q] * (1/W)
//       written by lanes 0..3.  Returns false when a singular value is not > DBL_MIN (zero or NaN): the pseudo-random-row
//       branch needs the whole sorted matrix, so the caller then runs cve_b_finish on one lane instead.
ZP_HD inline void cve_b_norms(const double* A, int lane, double* W) {
    for (int r = lane; r < 12; r += CVE_G) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) { const double t = A[r * CVE_RS + k]; sd += t * t; }
        W[r] = sqrt(sd);
    }
}

ZP_HD inline bool cve_b_tail(const double* A, int lane, const double* W, double* V4) {
    double w[12];
    int idx[12];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 12; i++) { w[i] = W[i]; idx[i] = i; ok = ok && w[i] > DBL_MIN; }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < 11; i++) {
        double wj = w[i];
        int j = i;
#pragma unroll
        for (int k = i + 1; k < 12; k++) { const bool lt = wj < w[k]; wj = lt ? w[k] : wj; j = lt ? k : j; }
        // swap positions i and j (j is dynamic: a select per later position)
        const double wi = w[i];
        const int ii = idx[i];
        int ij = ii;
#pragma unroll
        for (int k = i + 1; k < 12; k++) {
            const bool hit = k == j;
            ij = hit ? idx[k] : ij;
            w[k] = hit ? wi : w[k];
            idx[k] = hit ? ii : idx[k];
        }
        w[i] = wj; idx[i] = ij;
    }
    if (lane < 4) {
        int src = 0;
        double ws = 0;
#pragma unroll
        for (int i = 8; i < 12; i++) if (11 - i == lane) { src = idx[i]; ws = w[i]; }
        const double s = 1 / ws;
#pragma unroll
        for (int k = 0; k < 12; k++) V4[lane * 12 + k] = A[src * CVE_RS + k] * s;
    }
    return true;
}

// one lane: the generic tail (any singular values); keeps rows 11, 10, 9, 8 of U^T in V4 [4][12]; W: 12 doubles
ZP_HD inline void cve_b_finish(double* A, double* W, double* V4) {
    cve_finish(cve_dv(A, 1), CVE_RS, cve_dv(W, 1), cve_dv(nullptr, 1), 0, false, 12, 12, 12);
    for (int q = 0; q < 4; q++)
        for (int k = 0; k < 12; k++) V4[q * 12 + k] = A[(11 - q) * CVE_RS + k];
}

// lane r: row r of L (6x10) and rho[r]; V4 in shared memory, cw / L / rho strided views
ZP_HD inline void cve_b_L_rho(const double* V4, int r, Dv cw, Dv L, Dv rho) {
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    double dv[4][3];
    for (int q = 0; q < 4; q++)
        for (int e = 0; e < 3; e++) dv[q][e] = V4[q * 12 + 3 * pa[r] + e] - V4[q * 12 + 3 * pb[r] + e];
    L[10 * r + 0] = cve_dot3(dv[0], dv[0]);
    L[10 * r + 1] = 2.0f * cve_dot3(dv[0], dv[1]);
    L[10 * r + 2] = cve_dot3(dv[1], dv[1]);
    L[10 * r + 3] = 2.0f * cve_dot3(dv[0], dv[2]);
    L[10 * r + 4] = 2.0f * cve_dot3(dv[1], dv[2]);
    L[10 * r + 5] = cve_dot3(dv[2], dv[2]);
    L[10 * r + 6] = 2.0f * cve_dot3(dv[0], dv[3]);
    L[10 * r + 7] = 2.0f * cve_dot3(dv[1], dv[3]);
    L[10 * r + 8] = 2.0f * cve_dot3(dv[2], dv[3]);
    L[10 * r + 9] = cve_dot3(dv[3], dv[3]);
    const double a0 = cw[3 * pa[r]] - cw[3 * pb[r]], a1 = cw[3 * pa[r] + 1] - cw[3 * pb[r] + 1], a2 = cw[3 * pa[r] + 2] - cw[3 * pb[r] + 2];
    rho[r] = a0 * a0 + a1 * a1 + a2 * a2;
}

// ------------------------------------------------------------------------------------------------------------------
// stage C: one beta initialisation c (0, 1, 2 = OpenCV's N = 1, 2, 3) of one hypothesis.  Read-only views: L, rho, V4,
// al, pw, us.  wk: 60 doubles of scratch (least squares: At 30 | Vt 25 | w 5; then A3 9 | V3 9 | W3 3 | - | pcs 24).
// out: R[9] t[3] err (13 doubles).
// ------------------------------------------------------------------------------------------------------------------
template <int NC>
ZP_HD inline void cve_c_solve(Dv L, Dv rho, Dv wk, int c, double* x) {
    const Dv At = wk, Vt = wk.at(30), w = wk.at(55);
    const int cols0[4] = {0, 1, 3, 6};
    for (int q = 0; q < NC; q++) {
        const int col = c == 0 ? cols0[q] : q;
        for (int r = 0; r < 6; r++) At[q * 6 + r] = L[10 * r + col];
    }
    cve_jserial<6, true>(At, 6, Vt, NC, NC);
    cve_finish(At, 6, w, Vt, NC, true, 6, NC, NC);
    double rh[6];
    for (int r = 0; r < 6; r++) rh[r] = rho[r];
    cve_backsubst6(At, w, Vt, NC, rh, x);
}

ZP_HD inline void cve_stage_c(int c, int m, const CveCam& cam, Dv L, Dv rho, Dv V4, Dv al, Dv pws, Dv us, Dv wk, Dv out) {
    double x[5], be[4];
    if (c == 0) cve_c_solve<4>(L, rho, wk, c, x);
    else if (c == 1) cve_c_solve<3>(L, rho, wk, c, x);
    else cve_c_solve<5>(L, rho, wk, c, x);
    if (c == 0) {
        if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = -x[1] / be[0]; be[2] = -x[2] / be[0]; be[3] = -x[3] / be[0]; }
        else { be[0] = sqrt(x[0]); be[1] = x[1] / be[0]; be[2] = x[2] / be[0]; be[3] = x[3] / be[0]; }
    } else {
        if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = (x[2] < 0) ? sqrt(-x[2]) : 0.0; }
        else { be[0] = sqrt(x[0]); be[1] = (x[2] > 0) ? sqrt(x[2]) : 0.0; }
        if (x[1] < 0) be[0] = -be[0];
        be[2] = c == 2 ? x[3] / be[0] : 0.0;
        be[3] = 0.0;
    }
    cve_gauss_newton(L, rho, be);
    // camera-frame control points and points, sign
    const Dv A3 = wk, V3 = wk.at(9), W3 = wk.at(18), pcs = wk.at(24);
    double ccs[4][3];
    for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int q = 0; q < 3; q++) ccs[j][q] += be[i] * V4[12 * i + 3 * j + q];
    for (int p = 0; p < m; p++) {
        const double a0 = al[4 * p], a1 = al[4 * p + 1], a2 = al[4 * p + 2], a3 = al[4 * p + 3];
        for (int j = 0; j < 3; j++) pcs[3 * p + j] = a0 * ccs[0][j] + a1 * ccs[1][j] + a2 * ccs[2][j] + a3 * ccs[3][j];
    }
    if (pcs[2] < 0.0)
        for (int i = 0; i < 3 * m; i++) pcs[i] = -pcs[i];
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int p = 0; p < m; p++)
        for (int j = 0; j < 3; j++) { pc0[j] += pcs[3 * p + j]; pw0[j] += pws[3 * p + j]; }
    for (int j = 0; j < 3; j++) { pc0[j] /= m; pw0[j] /= m; }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < m; p++) {
        const double w0 = pws[3 * p] - pw0[0], w1 = pws[3 * p + 1] - pw0[1], w2 = pws[3 * p + 2] - pw0[2];
        for (int j = 0; j < 3; j++) {
            const double d = pcs[3 * p + j] - pc0[j];
            abt[3 * j] += d * w0;
            abt[3 * j + 1] += d * w1;
            abt[3 * j + 2] += d * w2;
        }
    }
    for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) A3[r * 3 + q] = abt[q * 3 + r];
    cve_jserial<3, true>(A3, 3, V3, 3, 3);
    cve_finish(A3, 3, W3, V3, 3, true, 3, 3, 3);
    double R[3][3], t[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[i][j] = A3[0 * 3 + i] * V3[0 * 3 + j] + A3[1 * 3 + i] * V3[1 * 3 + j] + A3[2 * 3 + i] * V3[2 * 3 + j];
    const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                       R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
    if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
    t[0] = pc0[0] - cve_dot3(R[0], pw0);
    t[1] = pc0[1] - cve_dot3(R[1], pw0);
    t[2] = pc0[2] - cve_dot3(R[2], pw0);
    double sum2 = 0.0;
    for (int p = 0; p < m; p++) {
        const double pw[3] = {pws[3 * p], pws[3 * p + 1], pws[3 * p + 2]};
        const double Xc = cve_dot3(R[0], pw) + t[0];
        const double Yc = cve_dot3(R[1], pw) + t[1];
        const double inv_Zc = 1.0 / (cve_dot3(R[2], pw) + t[2]);
        const double ue = cam.uc + cam.fu * Xc * inv_Zc;
        const double ve = cam.vc + cam.fv * Yc * inv_Zc;
        const double u = us[2 * p], v = us[2 * p + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[3 * i + j] = R[i][j];
    for (int i = 0; i < 3; i++) out[9 + i] = t[i];
    out[12] = sum2 / m;
}

// stage D: EPnP's choice among the three candidates ([3][13] view); returns the candidate index
ZP_HD inline int cve_pick(Dv o) {
    int N = 0;
    if (o[13 + 12] < o[12]) N = 1;
    if (o[26 + 12] < o[13 * N + 12]) N = 2;
    return N;
}
