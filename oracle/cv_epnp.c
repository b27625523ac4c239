/* oracle/cv_epnp.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Operation-by-operation float64 restatement of the arithmetic behind
 *     cv2.solvePnP(obj, img, K, None, flags=cv2.SOLVEPNP_EPNP)
 * which the reference calls through cv2.solvePnPRansac (zebrapose/binary_code_helper/CNN_output_to_pose.py:155-157).
 * OpenCV is an un-vendored dependency of the reference (unpinned there; opencv-python-headless 4.13.0.92 in this
 * image), so this file restates the published algorithm (Lepetit, Moreno-Noguer, Fua: EPnP, IJCV 2009) in the
 * evaluation order OpenCV's calib3d uses, and is PINNED against cv2 itself: tests/test_oracle_cv_epnp.py checks that
 * every stage (A^T A sums, small-matrix one-sided Jacobi SVD, SVD inverse / least squares) is bit-identical to
 * cv2.mulTransposed / cv2.SVDecomp / cv2.invert / cv2.solve, and that whole poses of 5-point samples (M^T M has a
 * 2-dimensional null space whose basis is decided by rounding) equal cv2.solvePnP's.
 *
 * What makes the replay exact (each item measured in this image):
 *   - calib3d and the small-matrix SVD are compiled without FMA contraction: every a*b+c is two roundings;
 *   - sums run sequentially in index order (no pairwise / SIMD re-association);
 *   - the Jacobi SVD uses its own hypot: |a|>|b| ? |a|*sqrt(1+(b/a)^2) : |b|*sqrt(1+(a/b)^2), not libm's;
 *   - image points reach EPnP as float32 normalised coordinates: u' = double(float((u-cx)*(1/fx)))*fx + cx.
 *
 * Build: `make -C oracle` (gcc -O2 -ffp-contract=off; no -mfma) -> oracle/_build/libcv_epnp.so
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXN 16  /* largest matrix side that goes through the Jacobi SVD here */

/* arithmetic-operation census of the Jacobi SVDs (add, mul, div, sqrt each count 1), by problem size: [0] n = 12 (the
 * null space of M^T M), [1] every other size; [2] / [3] rotated / skipped pairs of the n = 12 problems.  bench.py turns
 * the per-hypothesis averages into the algorithmic FP64 work of the solver kernels. */
static double zpo_census[4];
void zpo_census_reset(void) { zpo_census[0] = zpo_census[1] = zpo_census[2] = zpo_census[3] = 0; }
void zpo_census_read(double* out4) { for (int i = 0; i < 4; i++) out4[i] = zpo_census[i]; }

static double cv_hypot(double a, double b) {
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

/* One-sided (Hestenes) Jacobi on the n rows (length m) of At; Vt n x n accumulates the rotations.
 * On return the rows of At are the left singular vectors (normalised), W descending.  n1 = rows to normalise. */
int zpo_jacobi_svd(double* At, int astep, double* Wout, double* Vt, int vstep, int m, int n, int n1) {
    double W[MAXN];
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    int i, j, k, iter, max_iter = m > 30 ? m : 30;
    double c, s, sd;
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
        W[i] = sd;
        if (Vt) { for (k = 0; k < n; k++) Vt[i * vstep + k] = 0; Vt[i * vstep + i] = 1; }
    }
    for (iter = 0; iter < max_iter; iter++) {
        int changed = 0;
        for (i = 0; i < n - 1; i++)
            for (j = i + 1; j < n; j++) {
                double *Ai = At + i * astep, *Aj = At + j * astep;
                double a = W[i], p = 0, b = W[j];
                for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
                zpo_census[n == 12 ? 0 : 1] += 2.0 * m + 3;
                if (fabs(p) <= eps * sqrt(a * b)) { if (n == 12) zpo_census[3] += 1; continue; }
                /* hypot 5, c/s 8, rotation 6m, norms 4m, Vt rotation 6n */
                zpo_census[n == 12 ? 0 : 1] += 14.0 + 10.0 * m + (Vt ? 6.0 * n : 0.0);
                if (n == 12) zpo_census[2] += 1;
                p *= 2;
                double beta = a - b, gamma = cv_hypot(p, beta);
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (k = 0; k < m; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                if (Vt) {
                    double *Vi = Vt + i * vstep, *Vj = Vt + j * vstep;
                    for (k = 0; k < n; k++) {
                        double t0 = c * Vi[k] + s * Vj[k];
                        double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0; Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            if (Vt) {
                for (k = 0; k < m; k++) { t = At[i * astep + k]; At[i * astep + k] = At[j * astep + k]; At[j * astep + k] = t; }
                for (k = 0; k < n; k++) { t = Vt[i * vstep + k]; Vt[i * vstep + k] = Vt[j * vstep + k]; Vt[j * vstep + k] = t; }
            }
        }
    }
    for (i = 0; i < n; i++) Wout[i] = W[i];
    if (!Vt) return iter;
    uint64_t rng = 0x12345678;
    for (i = 0; i < n1; i++) {
        sd = i < n ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            /* exactly-zero singular value: random +-1/m vector, orthogonalised against the previous rows */
            const double val0 = 1. / m;
            for (k = 0; k < m; k++) {
                rng = (uint64_t)(uint32_t)rng * 4164903690U + (uint32_t)(rng >> 32);
                At[i * astep + k] = ((uint32_t)rng & 256) != 0 ? val0 : -val0;
            }
            for (int it2 = 0; it2 < 2; it2++)
                for (j = 0; j < i; j++) {
                    sd = 0;
                    for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
                    double asum = 0;
                    for (k = 0; k < m; k++) {
                        double t = At[i * astep + k] - sd * At[j * astep + k];
                        At[i * astep + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (k = 0; k < m; k++) At[i * astep + k] *= asum;
                }
            sd = 0;
            for (k = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
            sd = sqrt(sd);
        }
        s = sd > minval ? 1 / sd : 0.;
        for (k = 0; k < m; k++) At[i * astep + k] *= s;
    }
    return iter;
}

/* dst (c x c) = src^T src for src r x c: per entry a sequential sum over the rows (upper triangle, mirrored). */
void zpo_mul_transposed(const double* src, int r, int c, double* dst) {
    for (int i = 0; i < c; i++)
        for (int j = i; j < c; j++) {
            double s = 0;
            for (int k = 0; k < r; k++) s += src[k * c + i] * src[k * c + j];
            dst[i * c + j] = s; dst[j * c + i] = s;
        }
}

/* SVD of a square n x n matrix A (row-major): w, Ut (rows = left vectors), Vt, as SVD::compute arranges it. */
void zpo_svd_square(const double* A, int n, double* w, double* Ut, double* Vt) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ut[i * n + j] = A[j * n + i];
    zpo_jacobi_svd(Ut, n, w, Vt, n, n, n, n);
}

/* x = pinv(A) b through the Jacobi SVD (A m x n, m >= n, one right-hand side): the DECOMP_SVD solve. */
void zpo_solve_svd(const double* A, int m, int n, const double* b, double* x) {
    double at[MAXN * MAXN], w[MAXN], vt[MAXN * MAXN];
    for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) at[i * m + j] = A[j * n + i];
    zpo_jacobi_svd(at, m, w, vt, n, m, n, n);
    double thr = 0;
    for (int i = 0; i < n; i++) { x[i] = 0; thr += w[i]; }
    thr *= DBL_EPSILON * 2;
    for (int i = 0; i < n; i++) {
        double wi = w[i];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < m; j++) s += at[i * m + j] * b[j];
        s *= wi;
        for (int j = 0; j < n; j++) x[j] = x[j] + s * vt[i * n + j];
    }
}

/* inverse of a 3x3 through the SVD (DECOMP_SVD invert): dst = V diag(1/w) U^T accumulated singular value by value. */
void zpo_invert3_svd(const double* A, double* dst) {
    double ut[9], w[3], vt[9];
    zpo_svd_square(A, 3, w, ut, vt);
    double thr = (w[0] + w[1] + w[2]) * (DBL_EPSILON * 2);
    for (int i = 0; i < 9; i++) dst[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = w[i];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        double buf[3];
        for (int j = 0; j < 3; j++) buf[j] = ut[i * 3 + j] * wi;
        for (int r = 0; r < 3; r++) {
            double s = vt[i * 3 + r];
            for (int j = 0; j < 3; j++) dst[r * 3 + j] = dst[r * 3 + j] + s * buf[j];
        }
    }
}

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dist2(const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

/* Householder QR least squares of the 6x4 Gauss-Newton system in the evaluation order of the published EPnP code
 * (column scaling by the largest magnitude among the rows k..nr-2 -- the scan stops one row early). */
static void qr_solve64(double* A, double* b, double* X) {
    const int nr = 6, nc = 4;
    double A1[6], A2[6];
    double *pA = A, *ppAkk = pA;
    for (int k = 0; k < nc; k++) {
        double *ppAik1 = ppAkk, eta = fabs(*ppAik1);
        for (int i = k + 1; i < nr; i++) {
            double elt = fabs(*ppAik1);
            if (eta < elt) eta = elt;
            ppAik1 += nc;
        }
        if (eta == 0) { A1[k] = A2[k] = 0.0; return; }
        double *ppAik2 = ppAkk, sum2 = 0.0, inv_eta = 1. / eta;
        for (int i = k; i < nr; i++) { *ppAik2 *= inv_eta; sum2 += *ppAik2 * *ppAik2; ppAik2 += nc; }
        double sigma = sqrt(sum2);
        if (*ppAkk < 0) sigma = -sigma;
        *ppAkk += sigma;
        A1[k] = sigma * *ppAkk;
        A2[k] = -eta * sigma;
        for (int j = k + 1; j < nc; j++) {
            double *ppAik = ppAkk, sum = 0;
            for (int i = k; i < nr; i++) { sum += *ppAik * ppAik[j - k]; ppAik += nc; }
            double tau = sum / A1[k];
            ppAik = ppAkk;
            for (int i = k; i < nr; i++) { ppAik[j - k] -= tau * *ppAik; ppAik += nc; }
        }
        ppAkk += nc + 1;
    }
    double *ppAjj = pA, *pb = b;
    for (int j = 0; j < nc; j++) {
        double *ppAij = ppAjj, tau = 0;
        for (int i = j; i < nr; i++) { tau += *ppAij * pb[i]; ppAij += nc; }
        tau /= A1[j];
        ppAij = ppAjj;
        for (int i = j; i < nr; i++) { pb[i] -= tau * *ppAij; ppAij += nc; }
        ppAjj += nc + 1;
    }
    X[nc - 1] = pb[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; i--) {
        double *ppAij = pA + i * nc + (i + 1), sum = 0;
        for (int j = i + 1; j < nc; j++) { sum += *ppAij * X[j]; ppAij++; }
        X[i] = (pb[i] - sum) / A2[i];
    }
}

static void gauss_newton(const double* L, const double* rho, double* betas) {
    double a[24], b[6], x[4] = {0, 0, 0, 0};
    for (int it = 0; it < 5; it++) {
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
        qr_solve64(a, b, x);
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

typedef struct {
    int n;
    const double *pws, *us, *alphas;
    double fu, fv, uc, vc;
    double* pcs;
} prob_t;

static double compute_R_and_t(const prob_t* P, const double* ut, const double* betas, double R[3][3], double t[3]) {
    const int n = P->n;
    double ccs[4][3];
    for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; i++) {
        const double* v = ut + 12 * (11 - i);
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
    }
    double* pcs = P->pcs;
    for (int i = 0; i < n; i++) {
        const double* a = P->alphas + 4 * i;
        for (int j = 0; j < 3; j++)
            pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }
    if (pcs[2] < 0.0) {
        for (int i = 0; i < 4; i++) for (int j = 0; j < 3; j++) ccs[i][j] = -ccs[i][j];
        for (int i = 0; i < 3 * n; i++) pcs[i] = -pcs[i];
    }
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) { pc0[j] += pcs[3 * i + j]; pw0[j] += P->pws[3 * i + j]; }
    for (int j = 0; j < 3; j++) { pc0[j] /= n; pw0[j] /= n; }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, d[3], ut3[9], vt3[9];
    for (int i = 0; i < n; i++) {
        const double* pc = pcs + 3 * i;
        const double* pw = P->pws + 3 * i;
        for (int j = 0; j < 3; j++) {
            abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
            abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
            abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
        }
    }
    zpo_svd_square(abt, 3, d, ut3, vt3);
    /* R = U V^T with U = ut3^T, V = vt3^T */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[i][j] = ut3[0 * 3 + i] * vt3[0 * 3 + j] + ut3[1 * 3 + i] * vt3[1 * 3 + j] + ut3[2 * 3 + i] * vt3[2 * 3 + j];
    const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                       R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
    if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
    t[0] = pc0[0] - dot3(R[0], pw0);
    t[1] = pc0[1] - dot3(R[1], pw0);
    t[2] = pc0[2] - dot3(R[2], pw0);
    double sum2 = 0.0;
    for (int i = 0; i < n; i++) {
        const double* pw = P->pws + 3 * i;
        double Xc = dot3(R[0], pw) + t[0];
        double Yc = dot3(R[1], pw) + t[1];
        double inv_Zc = 1.0 / (dot3(R[2], pw) + t[2]);
        double ue = P->uc + P->fu * Xc * inv_Zc;
        double ve = P->vc + P->fv * Yc * inv_Zc;
        double u = P->us[2 * i], v = P->us[2 * i + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
}

/* EPnP of n correspondences.  pw [n,3] double, uv [n,2] pixels (double values of the float32 image points),
 * K = fu, fv, uc, vc.  f32_stage: 1 = image points were float32 (hypotheses), 2 = float64 (final solve), 0 = use uv as is.
 * scratch: 9n doubles (alphas 4n | pcs 3n | us 2n).  Out: R[9] row-major, t[3];
 * dbg (nullable, 12*12 + 12 + 12 + 3 doubles): Ut | singular values | the three candidates' betas | errors.
 * Returns the index (1..3) of the winning beta initialisation. */
int zpo_cv_epnp(const double* pw, const double* uv, int n, const double* K4, int f32_stage, double* scratch,
                double* Rout, double* tout, double* dbg) {
    const double fu = K4[0], fv = K4[1], uc = K4[2], vc = K4[3];
    double* alphas = scratch;
    double* pcs = scratch + 4 * n;
    double* us = scratch + 7 * n;
    const double ifx = 1. / fu, ify = 1. / fv;
    for (int i = 0; i < n; i++) {
        if (f32_stage) {
            /* undistortPoints (no distortion) -> normalised coordinates in the image points' own type (float32 for the
             * RANSAC hypotheses, float64 for the final solve on the inliers) -> back to pixels */
            double x = (uv[2 * i] - uc) * ifx, y = (uv[2 * i + 1] - vc) * ify;
            if (f32_stage == 1) { x = (double)(float)x; y = (double)(float)y; }
            us[2 * i] = x * fu + uc;
            us[2 * i + 1] = y * fv + vc;
        } else { us[2 * i] = uv[2 * i]; us[2 * i + 1] = uv[2 * i + 1]; }
    }
    /* control points: centroid + PCA axes */
    double cws[4][3] = {{0}};
    for (int i = 0; i < n; i++) for (int j = 0; j < 3; j++) cws[0][j] += pw[3 * i + j];
    for (int j = 0; j < 3; j++) cws[0][j] /= n;
    double pw0tpw0[9], dc[3], uct[9], vtmp[9];
    {
        /* PW0^T PW0 with PW0 rows pw_i - c0 (pcs used as scratch for PW0) */
        for (int i = 0; i < n; i++) for (int j = 0; j < 3; j++) pcs[3 * i + j] = pw[3 * i + j] - cws[0][j];
        zpo_mul_transposed(pcs, n, 3, pw0tpw0);
        zpo_svd_square(pw0tpw0, 3, dc, uct, vtmp);
    }
    for (int i = 1; i < 4; i++) {
        double k = sqrt(dc[i - 1] / n);
        for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + k * uct[3 * (i - 1) + j];
    }
    /* barycentric coordinates */
    double cc[9], ci[9];
    for (int i = 0; i < 3; i++) for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
    zpo_invert3_svd(cc, ci);
    for (int i = 0; i < n; i++) {
        const double* pi = pw + 3 * i;
        double* a = alphas + 4 * i;
        for (int j = 0; j < 3; j++)
            a[1 + j] = ci[3 * j] * (pi[0] - cws[0][0]) + ci[3 * j + 1] * (pi[1] - cws[0][1]) + ci[3 * j + 2] * (pi[2] - cws[0][2]);
        a[0] = 1.0f - a[1] - a[2] - a[3];
    }
    /* M^T M accumulated without materialising M when n is large: the sums are per entry, sequential over rows */
    double mtm[144];
    {
        double row1[12], row2[12];
        for (int i = 0; i < 144; i++) mtm[i] = 0;
        for (int p = 0; p < n; p++) {
            const double* as = alphas + 4 * p;
            const double u = us[2 * p], v = us[2 * p + 1];
            for (int i = 0; i < 4; i++) {
                row1[3 * i] = as[i] * fu; row1[3 * i + 1] = 0.0; row1[3 * i + 2] = as[i] * (uc - u);
                row2[3 * i] = 0.0; row2[3 * i + 1] = as[i] * fv; row2[3 * i + 2] = as[i] * (vc - v);
            }
            for (int i = 0; i < 12; i++)
                for (int j = i; j < 12; j++) {
                    double s = mtm[i * 12 + j];
                    s += row1[i] * row1[j];
                    s += row2[i] * row2[j];
                    mtm[i * 12 + j] = s;
                }
        }
        for (int i = 0; i < 12; i++) for (int j = 0; j < i; j++) mtm[i * 12 + j] = mtm[j * 12 + i];
    }
    double d[12], ut[144], vt[144];
    zpo_svd_square(mtm, 12, d, ut, vt);
    /* L (6x10) and rho */
    double L[60], rho[6];
    {
        const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
        double dv[4][6][3];
        for (int i = 0; i < 4; i++) {
            int a = 0, b = 1;
            for (int j = 0; j < 6; j++) {
                dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
                dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
                dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
                b++;
                if (b > 3) { a++; b = a + 1; }
            }
        }
        for (int i = 0; i < 6; i++) {
            double* row = L + 10 * i;
            row[0] = dot3(dv[0][i], dv[0][i]);
            row[1] = 2.0f * dot3(dv[0][i], dv[1][i]);
            row[2] = dot3(dv[1][i], dv[1][i]);
            row[3] = 2.0f * dot3(dv[0][i], dv[2][i]);
            row[4] = 2.0f * dot3(dv[1][i], dv[2][i]);
            row[5] = dot3(dv[2][i], dv[2][i]);
            row[6] = 2.0f * dot3(dv[0][i], dv[3][i]);
            row[7] = 2.0f * dot3(dv[1][i], dv[3][i]);
            row[8] = 2.0f * dot3(dv[2][i], dv[3][i]);
            row[9] = dot3(dv[3][i], dv[3][i]);
        }
        rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);
        rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
    }
    double Betas[4][4], rep[4], Rs[4][3][3], ts[4][3];
    prob_t P = {n, pw, us, alphas, fu, fv, uc, vc, pcs};
    {   /* N = 1: unknowns b00 b01 b02 b03 */
        double l[24], b4[4];
        for (int i = 0; i < 6; i++) { l[4 * i] = L[10 * i]; l[4 * i + 1] = L[10 * i + 1]; l[4 * i + 2] = L[10 * i + 3]; l[4 * i + 3] = L[10 * i + 6]; }
        zpo_solve_svd(l, 6, 4, rho, b4);
        double* be = Betas[1];
        if (b4[0] < 0) { be[0] = sqrt(-b4[0]); be[1] = -b4[1] / be[0]; be[2] = -b4[2] / be[0]; be[3] = -b4[3] / be[0]; }
        else { be[0] = sqrt(b4[0]); be[1] = b4[1] / be[0]; be[2] = b4[2] / be[0]; be[3] = b4[3] / be[0]; }
    }
    {   /* N = 2: b00 b01 b11 */
        double l[18], b3[3];
        for (int i = 0; i < 6; i++) { l[3 * i] = L[10 * i]; l[3 * i + 1] = L[10 * i + 1]; l[3 * i + 2] = L[10 * i + 2]; }
        zpo_solve_svd(l, 6, 3, rho, b3);
        double* be = Betas[2];
        if (b3[0] < 0) { be[0] = sqrt(-b3[0]); be[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
        else { be[0] = sqrt(b3[0]); be[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
        if (b3[1] < 0) be[0] = -be[0];
        be[2] = 0.0; be[3] = 0.0;
    }
    {   /* N = 3: b00 b01 b11 b02 b12 */
        double l[30], b5[5];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 5; j++) l[5 * i + j] = L[10 * i + j];
        zpo_solve_svd(l, 6, 5, rho, b5);
        double* be = Betas[3];
        if (b5[0] < 0) { be[0] = sqrt(-b5[0]); be[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
        else { be[0] = sqrt(b5[0]); be[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
        if (b5[1] < 0) be[0] = -be[0];
        be[2] = b5[3] / be[0];
        be[3] = 0.0;
    }
    for (int N = 1; N <= 3; N++) {
        gauss_newton(L, rho, Betas[N]);
        rep[N] = compute_R_and_t(&P, ut, Betas[N], Rs[N], ts[N]);
    }
    int N = 1;
    if (rep[2] < rep[1]) N = 2;
    if (rep[3] < rep[N]) N = 3;
    for (int i = 0; i < 3; i++) { tout[i] = ts[N][i]; for (int j = 0; j < 3; j++) Rout[3 * i + j] = Rs[N][i][j]; }
    if (dbg) {
        memcpy(dbg, ut, sizeof ut);
        memcpy(dbg + 144, d, sizeof d);
        for (int k = 1; k <= 3; k++) for (int i = 0; i < 4; i++) dbg[156 + 4 * (k - 1) + i] = Betas[k][i];
        for (int k = 1; k <= 3; k++) dbg[168 + k - 1] = rep[k];
    }
    return N;
}
