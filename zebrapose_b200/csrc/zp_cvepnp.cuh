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
// What is NOT the same is the schedule.  The cyclic Jacobi visits the pairs (0,1),(0,2)..(n-2,n-1) one after the other;
// a pair only depends on the last earlier pairs that touched its two rows, so pair (i,j) of sweep s can run at time
// T = n*s + i + j: up to n/2 pairs per step, a sweep every n steps, and a result that is bit-identical to the serial
// order because every rotation sees exactly the operands it would have seen (tests/test_cvepnp_host.py).  Six lanes
// share one hypothesis: the 12x12 problem keeps all six busy (66 pairs in 12 steps), the three beta candidates run on
// lanes 0-4.  The squared row norms the serial code carries along are sequential sums of the stored row, so a lane
// recomputes them from the row it loads (same values, same order) instead of passing them between lanes.
//
// The code is written as per-lane "phases" separated by group barriers, so that the kernel (zp_cvsolve.cu) and the host
// harness (tests/native/cvepnp_host.cu, lanes emulated by a loop) execute the same source.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#ifndef ZP_HD
#ifdef __CUDACC__
#define ZP_HD __host__ __device__
#else
#define ZP_HD
#endif
#endif

#ifdef __CUDA_ARCH__
#define CVE_NOINLINE __noinline__      // the SVD tail and the pair step are called from many phases: one copy each
#else
#define CVE_NOINLINE
#endif

#define CVE_G 6              // lanes per hypothesis
#define CVE_MAXM 8           // largest minimal-sample size
#define CVE_RS 13            // row stride (doubles) of the 12x12 matrix: rows of one hypothesis fall on distinct banks

// per-hypothesis scratch (offsets in doubles)
#define CVE_A 0              // [12][13] M^T M -> rotated rows; later: the three least-squares systems, then the pose slots
#define CVE_PW 156           // [m][3] object points
#define CVE_US 180           // [m][2] image points after the float32 staging
#define CVE_AL 196           // [m][4] barycentric coordinates
#define CVE_CW 228           // [4][3] control points
#define CVE_V4 240           // [4][12] rows 11,10,9,8 of U^T
#define CVE_L 288            // [6][10]
#define CVE_RHO 348          // [6]
#define CVE_A3 354           // [3][3] small Jacobi problem (PCA, control-point inverse)
#define CVE_V3 363           // [3][3]
#define CVE_W 372            // [12] singular values
#define CVE_CI 384           // [3][3] inverse of the control-point basis
#define CVE_OUT 393          // [3][13] per candidate: R[9] t[3] err
#define CVE_FLAGS 432        // 3 x 4 ints: "sweep s of problem q rotated something", slot s & 3
#define CVE_HB 439           // doubles per hypothesis (odd: consecutive hypotheses start on different banks)

// sub-layout of CVE_A during the beta stage: candidate c solves a 6 x NC[c] least-squares problem (rows of At = columns)
//   c = 0 (N = 1): NC 4   At @0   Vt @24   w @40   x @134
//   c = 1 (N = 2): NC 3   At @44  Vt @62   w @71   x @138
//   c = 2 (N = 3): NC 5   At @74  Vt @104  w @129  x @141
// and during the pose stage candidate c owns 48 doubles @48c: pcs[m][3] @0, At3 @24, Vt3 @33, W3 @42

struct CveCam { double fu, fv, uc, vc; };

ZP_HD inline double cve_hypot(double a, double b) {
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// one Jacobi pair: rows Ai, Aj of length M (and rows Vi, Vj of length n of the accumulated rotations)
// ------------------------------------------------------------------------------------------------------------------
template <int M, bool HASV>
ZP_HD inline bool cve_pair(double* Ai, double* Aj, double* Vi, double* Vj, int n) {
    const double eps = DBL_EPSILON * 10;
    double ri[M], rj[M];
#pragma unroll
    for (int k = 0; k < M; k++) { ri[k] = Ai[k]; rj[k] = Aj[k]; }
    double p = 0, a = 0, b = 0;
#pragma unroll
    for (int k = 0; k < M; k++) { p += ri[k] * rj[k]; a += ri[k] * ri[k]; b += rj[k] * rj[k]; }
    if (fabs(p) <= eps * sqrt(a * b)) return false;
    p *= 2;
    const double beta = a - b, gamma = cve_hypot(p, beta);
    double c, s;
    if (beta < 0) {
        const double delta = (gamma - beta) * 0.5;
        s = sqrt(delta / gamma);
        c = p / (gamma * s * 2);
    } else {
        c = sqrt((gamma + beta) / (gamma * 2));
        s = p / (gamma * c * 2);
    }
#pragma unroll
    for (int k = 0; k < M; k++) {
        const double t0 = c * ri[k] + s * rj[k];
        const double t1 = -s * ri[k] + c * rj[k];
        Ai[k] = t0; Aj[k] = t1;
    }
    if (HASV) {
        for (int k = 0; k < n; k++) {
            const double t0 = c * Vi[k] + s * Vj[k];
            const double t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0; Vj[k] = t1;
        }
    }
    return true;
}

// one lane's view of one Jacobi problem: n rows of length M at At (row stride astep), worked on by nl lanes of which this
// one is number l.  chg = 4 ints shared by the problem's lanes.
struct CveJ {
    double* At; double* Vt; int* chg;
    int astep, vstep, n, l, nl, max_iter;
    bool active, done;
};

ZP_HD inline CveJ cve_j_none() {
    CveJ j;
    j.At = nullptr; j.Vt = nullptr; j.chg = nullptr; j.astep = j.vstep = 0; j.n = 2; j.l = 0; j.nl = 0; j.max_iter = 0;
    j.active = false; j.done = true;
    return j;
}

ZP_HD inline CveJ cve_j_make(double* At, int astep, double* Vt, int vstep, int n, int m_cols, int l, int nl, int* chg) {
    CveJ j;
    j.At = At; j.Vt = Vt; j.chg = chg; j.astep = astep; j.vstep = vstep; j.n = n; j.l = l; j.nl = nl;
    j.max_iter = m_cols > 30 ? m_cols : 30;
    j.active = l < nl; j.done = !j.active;
    return j;
}

// identity in Vt and cleared flags; called by lane l == 0 of the problem before the first step (then a barrier)
ZP_HD inline void cve_j_init(const CveJ& j) {
    if (!j.active || j.l != 0) return;
    if (j.Vt)
        for (int i = 0; i < j.n; i++)
            for (int k = 0; k < j.n; k++) j.Vt[i * j.vstep + k] = i == k ? 1.0 : 0.0;
    j.chg[0] = j.chg[1] = j.chg[2] = j.chg[3] = 0;
}

// step T (1, 2, ...): this lane's pair, if it has one.  Pairs of sweep s = (T - tau)/n with i + j = tau; two sweeps overlap.
template <int M, bool HASV>
ZP_HD inline void cve_jstep_a(CveJ& j, int T) {
    const int n = j.n;
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
    double* Vi = HASV ? j.Vt + ii * j.vstep : nullptr;
    double* Vj = HASV ? j.Vt + jj * j.vstep : nullptr;
    if (cve_pair<M, HASV>(j.At + ii * j.astep, j.At + jj * j.astep, Vi, Vj, n)) j.chg[s & 3] = 1;
}

// after the barrier that follows step T: if a sweep completed at T, stop when it rotated nothing (or at the sweep cap)
ZP_HD inline void cve_jstep_c(CveJ& j, int T) {
    const int n = j.n;
    const int num = T - (2 * n - 3);
    if (num < 0 || num % n != 0) return;
    const int s = num / n;
    const int c = j.chg[s & 3];
    if (j.l == 0) j.chg[(s + 2) & 3] = 0;
    if (!c || s + 1 >= j.max_iter) j.done = true;
}

// ------------------------------------------------------------------------------------------------------------------
// the tail of the SVD routine: singular values = row norms, selection sort (descending, rows swapped physically),
// rows normalised (an exactly-zero singular value gets a pseudo-random row orthogonalised against the rows above it)
// ------------------------------------------------------------------------------------------------------------------
ZP_HD CVE_NOINLINE inline void cve_finish(double* At, int astep, double* W, double* Vt, int vstep, int m, int n, int n1) {
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
            // (the library swaps the rows "if Vt"; every call replayed here has a Vt, stored or not)
            for (k = 0; k < m; k++) { t = At[i * astep + k]; At[i * astep + k] = At[j * astep + k]; At[j * astep + k] = t; }
            if (Vt) for (k = 0; k < n; k++) { t = Vt[i * vstep + k]; Vt[i * vstep + k] = Vt[j * vstep + k]; Vt[j * vstep + k] = t; }
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
ZP_HD inline void cve_backsubst6(const double* At, const double* w, const double* Vt, int nc, const double* b, double* x) {
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
        for (int i = k + 1; i < nr; i++) {
            const double elt = fabs(A[(i - 1) * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) { A1[k] = A2[k] = 0.0; return; }
        double sum2 = 0.0;
        const double inv_eta = 1. / eta;
        for (int i = k; i < nr; i++) { A[i * nc + k] *= inv_eta; sum2 += A[i * nc + k] * A[i * nc + k]; }
        double sigma = sqrt(sum2);
        if (A[k * nc + k] < 0) sigma = -sigma;
        A[k * nc + k] += sigma;
        A1[k] = sigma * A[k * nc + k];
        A2[k] = -eta * sigma;
#pragma unroll
        for (int j = k + 1; j < nc; j++) {
            double sum = 0;
            for (int i = k; i < nr; i++) sum += A[i * nc + k] * A[i * nc + j];
            const double tau = sum / A1[k];
            for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
        }
    }
#pragma unroll
    for (int j = 0; j < nc; j++) {
        double tau = 0;
        for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
        tau /= A1[j];
        for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
#pragma unroll
    for (int i = nc - 2; i >= 0; i--) {
        double sum = 0;
        for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
        X[i] = (b[i] - sum) / A2[i];
    }
}

ZP_HD inline void cve_gauss_newton(const double* L, const double* rho, double* betas) {
    double a[24], b[6], x[4] = {0, 0, 0, 0};
    for (int it = 0; it < 5; it++) {
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double* rowL = L + i * 10;
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
ZP_HD inline double cve_dist2(const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

// ------------------------------------------------------------------------------------------------------------------
// phases.  S = the hypothesis' scratch block, lane = 0..CVE_G-1; a group barrier separates consecutive phases.
// ------------------------------------------------------------------------------------------------------------------

// phase 0 (lane 0): stage the m sampled correspondences, centroid, scatter matrix -> A3 (transposed = itself)
// uvxyz: the five float planes of the crop (u | v | X | Y | Z, `cap` apart); idx: the m sample indices (all valid)
ZP_HD inline void cve_ph0(double* S, int lane, const float* corr, int cap, const int32_t* idx, int m, const CveCam& cam) {
    if (lane != 0) return;
    double* pw = S + CVE_PW; double* us = S + CVE_US; double* cws = S + CVE_CW;
    const double ifx = 1. / cam.fu, ify = 1. / cam.fv;
    for (int p = 0; p < m; p++) {
        const int i = idx[p];
        const double u = (double)corr[i], v = (double)corr[(size_t)cap + i];
        pw[3 * p] = (double)corr[2 * (size_t)cap + i];
        pw[3 * p + 1] = (double)corr[3 * (size_t)cap + i];
        pw[3 * p + 2] = (double)corr[4 * (size_t)cap + i];
        const double x = (u - cam.uc) * ifx, y = (v - cam.vc) * ify;
        us[2 * p] = (double)(float)x * cam.fu + cam.uc;
        us[2 * p + 1] = (double)(float)y * cam.fv + cam.vc;
    }
    cws[0] = cws[1] = cws[2] = 0;
    for (int p = 0; p < m; p++) for (int j = 0; j < 3; j++) cws[j] += pw[3 * p + j];
    for (int j = 0; j < 3; j++) cws[j] /= m;
    double* A3 = S + CVE_A3;
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
            double s = 0;
            for (int p = 0; p < m; p++) s += (pw[3 * p + i] - cws[i]) * (pw[3 * p + j] - cws[j]);
            A3[i * 3 + j] = s; A3[j * 3 + i] = s;
        }
}

// phase 1 (lane 0): PCA finished -> control points -> their basis matrix, transposed, into A3 for the SVD inverse
ZP_HD inline void cve_ph1(double* S, int lane, int m) {
    if (lane != 0) return;
    double* A3 = S + CVE_A3; double* V3 = S + CVE_V3; double* W = S + CVE_W; double* cws = S + CVE_CW;
    cve_finish(A3, 3, W, V3, 3, 3, 3, 3);
    for (int i = 1; i < 4; i++) {
        const double k = sqrt(W[i - 1] / m);
        for (int j = 0; j < 3; j++) cws[3 * i + j] = cws[j] + k * A3[3 * (i - 1) + j];
    }
    // cc[3*i + j-1] = cws[j][i] - cws[0][i]; the SVD works on the transpose: At[r][c] = cc[c][r]
    double cc[9];
    for (int i = 0; i < 3; i++) for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[3 * j + i] - cws[i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A3[r * 3 + c] = cc[c * 3 + r];
}

// phase 2 (lane 0): inverse of the control-point basis from its SVD: V diag(1/w) U^T, one singular value at a time
ZP_HD inline void cve_ph2(double* S, int lane) {
    if (lane != 0) return;
    double* A3 = S + CVE_A3; double* V3 = S + CVE_V3; double* W = S + CVE_W; double* ci = S + CVE_CI;
    cve_finish(A3, 3, W, V3, 3, 3, 3, 3);
    const double thr = (W[0] + W[1] + W[2]) * (DBL_EPSILON * 2);
    for (int i = 0; i < 9; i++) ci[i] = 0;
    for (int i = 0; i < 3; i++) {
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
}

// phase 3 (all lanes): barycentric coordinates, point p on lane p % G
ZP_HD inline void cve_ph3(double* S, int lane, int m) {
    const double* pw = S + CVE_PW; const double* cws = S + CVE_CW; const double* ci = S + CVE_CI;
    for (int p = lane; p < m; p += CVE_G) {
        const double* pi = pw + 3 * p;
        double* a = S + CVE_AL + 4 * p;
        for (int j = 0; j < 3; j++)
            a[1 + j] = ci[3 * j] * (pi[0] - cws[0]) + ci[3 * j + 1] * (pi[1] - cws[1]) + ci[3 * j + 2] * (pi[2] - cws[2]);
        a[0] = 1.0f - a[1] - a[2] - a[3];
    }
}

// phase 4 (all lanes): M^T M, entry (i, j >= i) = sequential sum over the 2m rows of M, 13 entries per lane
ZP_HD inline void cve_ph4(double* S, int lane, int m, const CveCam& cam) {
    const double* al = S + CVE_AL; const double* us = S + CVE_US;
    double* A = S + CVE_A;
    for (int e = lane; e < 78; e += CVE_G) {
        int i = 0, r = e;
        while (r >= 12 - i) { r -= 12 - i; i++; }
        const int j = i + r;
        const int ci = i / 3, xi = i - 3 * ci, cj = j / 3, xj = j - 3 * cj;
        double s = 0;
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
        A[i * CVE_RS + j] = s; A[j * CVE_RS + i] = s;
    }
}

// phase 5 (lane 0): finish the 12x12 SVD, keep rows 11, 10, 9, 8 of U^T
ZP_HD inline void cve_ph5(double* S, int lane) {
    if (lane != 0) return;
    double* A = S + CVE_A; double* W = S + CVE_W; double* V4 = S + CVE_V4;
    cve_finish(A, CVE_RS, W, nullptr, 0, 12, 12, 12);
    for (int q = 0; q < 4; q++)
        for (int k = 0; k < 12; k++) V4[q * 12 + k] = A[(11 - q) * CVE_RS + k];
}

// phase 6 (lane r): row r of L (6x10) and rho[r]
ZP_HD inline void cve_ph6(double* S, int lane) {
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    const double* V4 = S + CVE_V4; const double* cws = S + CVE_CW;
    for (int r = lane; r < 6; r += CVE_G) {
        double dv[4][3];
        for (int q = 0; q < 4; q++)
            for (int e = 0; e < 3; e++) dv[q][e] = V4[q * 12 + 3 * pa[r] + e] - V4[q * 12 + 3 * pb[r] + e];
        double* row = S + CVE_L + 10 * r;
        row[0] = cve_dot3(dv[0], dv[0]);
        row[1] = 2.0f * cve_dot3(dv[0], dv[1]);
        row[2] = cve_dot3(dv[1], dv[1]);
        row[3] = 2.0f * cve_dot3(dv[0], dv[2]);
        row[4] = 2.0f * cve_dot3(dv[1], dv[2]);
        row[5] = cve_dot3(dv[2], dv[2]);
        row[6] = 2.0f * cve_dot3(dv[0], dv[3]);
        row[7] = 2.0f * cve_dot3(dv[1], dv[3]);
        row[8] = 2.0f * cve_dot3(dv[2], dv[3]);
        row[9] = cve_dot3(dv[3], dv[3]);
        S[CVE_RHO + r] = cve_dist2(cws + 3 * pa[r], cws + 3 * pb[r]);
    }
}

struct CveCand { int nc, at, vt, w, x, lane0, nl; };
ZP_HD inline CveCand cve_cand(int c) {
    CveCand k;
    if (c == 0) { k.nc = 4; k.at = 0; k.vt = 24; k.w = 40; k.x = 134; k.lane0 = 0; k.nl = 2; }
    else if (c == 1) { k.nc = 3; k.at = 44; k.vt = 62; k.w = 71; k.x = 138; k.lane0 = 2; k.nl = 1; }
    else { k.nc = 5; k.at = 74; k.vt = 104; k.w = 129; k.x = 141; k.lane0 = 3; k.nl = 2; }
    return k;
}
// candidate a lane works for during the beta stage (-1: none) and whether it is the candidate's owner lane
ZP_HD inline int cve_lane_cand(int lane) { return lane < 2 ? 0 : lane == 2 ? 1 : lane < 5 ? 2 : -1; }

// phase 7 (owner lanes): the three sub-systems of L, transposed (rows of At = columns of the 6 x NC matrix)
ZP_HD inline void cve_ph7(double* S, int lane) {
    const int c = cve_lane_cand(lane);
    if (c < 0) return;
    const CveCand k = cve_cand(c);
    if (lane != k.lane0) return;
    const int cols0[4] = {0, 1, 3, 6};
    const double* L = S + CVE_L;
    double* At = S + CVE_A + k.at;
    for (int q = 0; q < k.nc; q++) {
        const int col = c == 0 ? cols0[q] : q;
        for (int r = 0; r < 6; r++) At[q * 6 + r] = L[10 * r + col];
    }
}

// phase 8 (owner lanes): least-squares solution -> beta initialisation -> 5 Gauss-Newton steps -> camera-frame control
// points and points, sign, centroids, the 3x3 correlation matrix (transposed into the candidate's At3)
ZP_HD inline void cve_ph8(double* S, int lane, int m, double* betas_out) {
    const int c = cve_lane_cand(lane);
    if (c < 0) return;
    const CveCand k = cve_cand(c);
    if (lane != k.lane0) return;
    double* At = S + CVE_A + k.at; double* Vt = S + CVE_A + k.vt; double* w = S + CVE_A + k.w;
    cve_finish(At, 6, w, Vt, k.nc, 6, k.nc, k.nc);
    double x[5], be[4];
    cve_backsubst6(At, w, Vt, k.nc, S + CVE_RHO, x);
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
    double L[60], rho[6];
    for (int i = 0; i < 60; i++) L[i] = S[CVE_L + i];
    for (int i = 0; i < 6; i++) rho[i] = S[CVE_RHO + i];
    cve_gauss_newton(L, rho, be);
    for (int i = 0; i < 4; i++) betas_out[i] = be[i];
}

// phase 9 (owner lanes, after a barrier: the least-squares scratch is dead): pose slots
ZP_HD inline void cve_ph9(double* S, int lane, int m, const double* be) {
    const int c = cve_lane_cand(lane);
    if (c < 0) return;
    const CveCand k = cve_cand(c);
    if (lane != k.lane0) return;
    const double* V4 = S + CVE_V4; const double* al = S + CVE_AL; const double* pws = S + CVE_PW;
    double* slot = S + CVE_A + 48 * c;
    double* pcs = slot; double* At3 = slot + 24;
    double ccs[4][3];
    for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; i++) {
        const double* v = V4 + 12 * i;
        for (int j = 0; j < 4; j++)
            for (int q = 0; q < 3; q++) ccs[j][q] += be[i] * v[3 * j + q];
    }
    for (int p = 0; p < m; p++) {
        const double* a = al + 4 * p;
        for (int j = 0; j < 3; j++)
            pcs[3 * p + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }
    if (pcs[2] < 0.0)
        for (int i = 0; i < 3 * m; i++) pcs[i] = -pcs[i];
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int p = 0; p < m; p++)
        for (int j = 0; j < 3; j++) { pc0[j] += pcs[3 * p + j]; pw0[j] += pws[3 * p + j]; }
    for (int j = 0; j < 3; j++) { pc0[j] /= m; pw0[j] /= m; }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < m; p++) {
        const double* pc = pcs + 3 * p;
        const double* pw = pws + 3 * p;
        for (int j = 0; j < 3; j++) {
            abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
            abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
            abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
        }
    }
    for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) At3[r * 3 + q] = abt[q * 3 + r];
    // the centroids are needed again after the SVD: keep them where pcs of points >= 6 would be only if m <= 6;
    // recomputing them in phase 10 is cheaper than finding room
}

// phase 10 (owner lanes): R = U V^T, det fix, t, mean reprojection distance -> CVE_OUT slot of the candidate
ZP_HD inline void cve_ph10(double* S, int lane, int m, const CveCam& cam) {
    const int c = cve_lane_cand(lane);
    if (c < 0) return;
    const CveCand k = cve_cand(c);
    if (lane != k.lane0) return;
    const double* pws = S + CVE_PW; const double* us = S + CVE_US;
    double* slot = S + CVE_A + 48 * c;
    double* pcs = slot; double* ut3 = slot + 24; double* vt3 = slot + 33; double* W3 = slot + 42;
    cve_finish(ut3, 3, W3, vt3, 3, 3, 3, 3);
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int p = 0; p < m; p++)
        for (int j = 0; j < 3; j++) { pc0[j] += pcs[3 * p + j]; pw0[j] += pws[3 * p + j]; }
    for (int j = 0; j < 3; j++) { pc0[j] /= m; pw0[j] /= m; }
    double R[3][3], t[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[i][j] = ut3[0 * 3 + i] * vt3[0 * 3 + j] + ut3[1 * 3 + i] * vt3[1 * 3 + j] + ut3[2 * 3 + i] * vt3[2 * 3 + j];
    const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                       R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
    if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
    t[0] = pc0[0] - cve_dot3(R[0], pw0);
    t[1] = pc0[1] - cve_dot3(R[1], pw0);
    t[2] = pc0[2] - cve_dot3(R[2], pw0);
    double sum2 = 0.0;
    for (int p = 0; p < m; p++) {
        const double* pw = pws + 3 * p;
        const double Xc = cve_dot3(R[0], pw) + t[0];
        const double Yc = cve_dot3(R[1], pw) + t[1];
        const double inv_Zc = 1.0 / (cve_dot3(R[2], pw) + t[2]);
        const double ue = cam.uc + cam.fu * Xc * inv_Zc;
        const double ve = cam.vc + cam.fv * Yc * inv_Zc;
        const double u = us[2 * p], v = us[2 * p + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    double* out = S + CVE_OUT + 13 * c;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[3 * i + j] = R[i][j];
    for (int i = 0; i < 3; i++) out[9 + i] = t[i];
    out[12] = sum2 / m;
}

// phase 11 (any lane after a barrier): EPnP's choice among the three candidates; returns the slot (R[9] t[3] err)
ZP_HD inline const double* cve_pick(const double* S) {
    const double* o = S + CVE_OUT;
    int N = 0;
    if (o[13 + 12] < o[12]) N = 1;
    if (o[26 + 12] < o[13 * N + 12]) N = 2;
    return o + 13 * N;
}
