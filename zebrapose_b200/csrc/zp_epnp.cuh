// EPnP in float64 as cv2.solvePnP(SOLVEPNP_EPNP) runs it (OpenCV is the un-vendored library behind the reference's
// cv2.solvePnPRansac call, CNN_output_to_pose.py:155-157).  Device restatement of the published algorithm
// (Lepetit, Moreno-Noguer, Fua 2009) with the OpenCV behaviours that change the answer under pixel noise:
// M built in pixel units with the real camera matrix (fu != fv weights the image axes), PCA control points with the signs of OpenCV's one-sided Jacobi SVD (same pair
// order and rotation formulas), three beta initialisations x 5 Gauss-Newton steps, Horn alignment with
// "negate row 2 when det < 0", best of three by mean reprojection distance.  oracle/epnp.py is the CPU twin.
//
// The per-point work is reduced to 52 sums (ZpSums) so the same core serves the 4/5/6-point minimal solver (one
// thread per hypothesis, sums over m points in registers) and the final solve on thousands of inliers (one CTA
// per crop, block reduction of the sums).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define ZP_HD __host__ __device__
#define ZP_DBL_EPS 2.220446049250313e-16
#define ZP_DBL_MIN 2.2250738585072014e-308

// strided view of an n x n double matrix (row-major) living in shared memory; `stride` interleaves the matrices
// of the threads of a CTA so that a warp touching element (r,k) of its 32 matrices hits 32 consecutive doubles.
struct ZpMat {
    double* p;
    int stride;
    ZP_HD __forceinline__ double& operator()(int r, int k, int n) const { return p[(size_t)(r * n + k) * stride]; }
};

// One-sided (Hestenes) Jacobi on the rows of At (= A^T), OpenCV's pair order (i<j ascending), rotation formulas and
// stopping rule (JacobiSVDImpl_, modules/core/src/lapack.cpp: eps = 10*DBL_EPSILON, max(n,30) sweeps).
// On exit row i of At = sigma_i * u_i^T and W[i] = sigma_i (unsorted).  If G != nullptr it receives the accumulated
// rotations (OpenCV's Vt) as a dense row-major n x n array in registers/local memory.
template <int N>
ZP_HD void zp_jacobi_rows(ZpMat At, double* W, double* G) {
    const double eps = ZP_DBL_EPS * 10;
    for (int i = 0; i < N; i++) {
        double sd = 0;
        for (int k = 0; k < N; k++) { double t = At(i, k, N); sd = fma(t, t, sd); }
        W[i] = sd;
    }
    if (G) {
        for (int i = 0; i < N * N; i++) G[i] = 0;
        for (int i = 0; i < N; i++) G[i * N + i] = 1;
    }
    const int max_iter = N > 30 ? N : 30;
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < N - 1; i++)
            for (int j = i + 1; j < N; j++) {
                double a = W[i], b = W[j], p = 0;
#pragma unroll
                for (int k = 0; k < N; k++) p = fma(At(i, k, N), At(j, k, N), p);
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = 0; b = 0;
#pragma unroll
                for (int k = 0; k < N; k++) {
                    double x = At(i, k, N), y = At(j, k, N);
                    double t0 = c * x + s * y, t1 = -s * x + c * y;
                    At(i, k, N) = t0; At(j, k, N) = t1;
                    a = fma(t0, t0, a); b = fma(t1, t1, b);
                }
                W[i] = a; W[j] = b;
                changed = true;
                if (G) {
#pragma unroll
                    for (int k = 0; k < N; k++) {
                        double x = G[i * N + k], y = G[j * N + k];
                        G[i * N + k] = c * x + s * y; G[j * N + k] = -s * x + c * y;
                    }
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; i++) {
        double sd = 0;
        for (int k = 0; k < N; k++) { double t = At(i, k, N); sd = fma(t, t, sd); }
        W[i] = sqrt(sd);
    }
}

// 3x3 SVD-type helpers on a private (stride 1) matrix --------------------------------------------------------------

// PCA of the 3x3 scatter matrix C (symmetric): returns singular values dc[3] descending and rows uct[3][3] with the
// signs OpenCV's SVD (U_T) returns.
ZP_HD inline void zp_pca3(const double C[9], double dc[3], double uct[9]) {
    double a[9], W[3];
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) a[i * 3 + k] = C[k * 3 + i];   // At = C^T
    ZpMat At{a, 1};
    zp_jacobi_rows<3>(At, W, nullptr);
    int idx[3] = {0, 1, 2};
    // selection sort descending, as OpenCV sorts W (swap on strict <)
    for (int i = 0; i < 2; i++) {
        int j = i;
        for (int k = i + 1; k < 3; k++) if (W[idx[j]] < W[idx[k]]) j = k;
        int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
    }
    for (int i = 0; i < 3; i++) {
        double w = W[idx[i]];
        dc[i] = w;
        double s = w > ZP_DBL_MIN ? 1.0 / w : 0.0;
        for (int k = 0; k < 3; k++) uct[i * 3 + k] = a[idx[i] * 3 + k] * s;
    }
}

// Orthogonal polar factor U V^T of a 3x3 matrix H (row-major) via the same one-sided Jacobi; rank-2 inputs are
// completed with a cross product.
ZP_HD inline void zp_polar3(const double H[9], double R[9]) {
    double a[9], W[3], G[9];
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) a[i * 3 + k] = H[k * 3 + i];   // At = H^T
    ZpMat At{a, 1};
    zp_jacobi_rows<3>(At, W, G);
    // H = sum_i (row_i(At)/W_i)^T * W_i * row_i(G)  ->  U V^T = sum_i u_i g_i^T
    double wmax = fmax(W[0], fmax(W[1], W[2]));
    int bad = -1, nbad = 0;
    for (int i = 0; i < 3; i++) {
        if (W[i] > wmax * 1e-13 && W[i] > ZP_DBL_MIN) {
            double s = 1.0 / W[i];
            for (int k = 0; k < 3; k++) a[i * 3 + k] *= s;
        } else { bad = i; nbad++; }
    }
    if (nbad == 1) {
        int i1 = (bad + 1) % 3, i2 = (bad + 2) % 3;
        a[bad * 3 + 0] = a[i1 * 3 + 1] * a[i2 * 3 + 2] - a[i1 * 3 + 2] * a[i2 * 3 + 1];
        a[bad * 3 + 1] = a[i1 * 3 + 2] * a[i2 * 3 + 0] - a[i1 * 3 + 0] * a[i2 * 3 + 2];
        a[bad * 3 + 2] = a[i1 * 3 + 0] * a[i2 * 3 + 1] - a[i1 * 3 + 1] * a[i2 * 3 + 0];
    } else if (nbad > 1) {
        for (int k = 0; k < 9; k++) R[k] = nan("");
        return;
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            R[r * 3 + c] = a[0 * 3 + r] * G[0 * 3 + c] + a[1 * 3 + r] * G[1 * 3 + c] + a[2 * 3 + r] * G[2 * 3 + c];
}

// Least squares min ||A x - b|| for a 6 x NC system by Householder QR (A row-major 6 x NC, destroyed).
template <int NC>
ZP_HD inline void zp_ls6(double* A, double* b, double* x) {
    const int NR = 6;
#pragma unroll
    for (int k = 0; k < NC; k++) {
        double nrm = 0;
        for (int i = k; i < NR; i++) nrm = fma(A[i * NC + k], A[i * NC + k], nrm);
        nrm = sqrt(nrm);
        double akk = A[k * NC + k];
        double alpha = akk > 0 ? -nrm : nrm;
        double v0 = akk - alpha;
        double vnorm2 = fma(v0, v0, nrm * nrm - akk * akk);       // |v|^2
        A[k * NC + k] = alpha;
        if (vnorm2 > 0) {
            double inv = 2.0 / vnorm2;
            for (int j = k + 1; j < NC; j++) {
                double d = v0 * A[k * NC + j];
                for (int i = k + 1; i < NR; i++) d = fma(A[i * NC + k], A[i * NC + j], d);
                d *= inv;
                A[k * NC + j] -= d * v0;
                for (int i = k + 1; i < NR; i++) A[i * NC + j] -= d * A[i * NC + k];
            }
            double d = v0 * b[k];
            for (int i = k + 1; i < NR; i++) d = fma(A[i * NC + k], b[i], d);
            d *= inv;
            b[k] -= d * v0;
            for (int i = k + 1; i < NR; i++) b[i] -= d * A[i * NC + k];
        }
    }
#pragma unroll
    for (int k = NC - 1; k >= 0; k--) {
        double s = b[k];
        for (int j = k + 1; j < NC; j++) s -= A[k * NC + j] * x[j];
        x[k] = s / A[k * NC + k];
    }
}

// The 52 sums over the points that EPnP needs once the control points are fixed.
struct ZpCam { double fu, fv, uc, vc; };

struct ZpSums {
    // sum a_j a_k {1, x, y, x^2+y^2} with x = uc - u, y = vc - v (pixels), (j<=k) packed: 00 01 02 03 11 12 13 22 23 33
    double s0[10], sx[10], sy[10], sr[10];
    double w[12];                            // W_j = sum_i a_ij (pw_i - pw0), j = 0..3
    double n;
};

ZP_HD __forceinline__ int zp_pk(int j, int k) {     // packed index of the symmetric 4x4 (j<=k)
    const int base[4] = {0, 4, 7, 9};
    return base[j] + (k - j);
}

struct ZpControl {
    double cws[12];      // 4 control points (world)
    double cci[9];       // rows j: alpha_{j+1} = cci[j] . (p - cws[0])
};

// control points + barycentric basis from the centroid c0, the 3x3 scatter matrix C = sum (p-c0)(p-c0)^T and n
ZP_HD inline void zp_control_points(const double c0[3], const double C[9], double n, ZpControl& cp) {
    double dc[3], uct[9];
    zp_pca3(C, dc, uct);
    double kk[3];
    for (int k = 0; k < 3; k++) cp.cws[k] = c0[k];
    for (int i = 0; i < 3; i++) {
        kk[i] = sqrt(dc[i] / n);
        for (int k = 0; k < 3; k++) cp.cws[3 * (i + 1) + k] = c0[k] + kk[i] * uct[3 * i + k];
    }
    // CC = [k1 u1 | k2 u2 | k3 u3] -> pseudo-inverse rows u_j^T / k_j (cv::invert(DECOMP_SVD) zeroes tiny singular values)
    double thr = (kk[0] + kk[1] + kk[2]) * (2 * ZP_DBL_EPS);
    for (int j = 0; j < 3; j++) {
        double inv = kk[j] > thr ? 1.0 / kk[j] : 0.0;
        for (int k = 0; k < 3; k++) cp.cci[3 * j + k] = uct[3 * j + k] * inv;
    }
}

ZP_HD __forceinline__ void zp_alphas(const ZpControl& cp, double X, double Y, double Z, double a[4]) {
    double dx = X - cp.cws[0], dy = Y - cp.cws[1], dz = Z - cp.cws[2];
    a[1] = cp.cci[0] * dx + cp.cci[1] * dy + cp.cci[2] * dz;
    a[2] = cp.cci[3] * dx + cp.cci[4] * dy + cp.cci[5] * dz;
    a[3] = cp.cci[6] * dx + cp.cci[7] * dy + cp.cci[8] * dz;
    a[0] = 1.0 - a[1] - a[2] - a[3];
}

ZP_HD __forceinline__ void zp_accumulate(ZpSums& s, const double a[4], double x, double y, double dX, double dY,
                                              double dZ) {
    double r = x * x + y * y;
    int q = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = j; k < 4; k++) {
            double p = a[j] * a[k];
            s.s0[q] += p; s.sx[q] = fma(p, x, s.sx[q]); s.sy[q] = fma(p, y, s.sy[q]); s.sr[q] = fma(p, r, s.sr[q]);
            q++;
        }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        s.w[3 * j + 0] = fma(a[j], dX, s.w[3 * j + 0]);
        s.w[3 * j + 1] = fma(a[j], dY, s.w[3 * j + 1]);
        s.w[3 * j + 2] = fma(a[j], dZ, s.w[3 * j + 2]);
    }
}

// M^T M (12x12) from the sums, written into At (symmetric, so At = MtM).
// rows of M (epnp::fill_M): [a_j fu, 0, a_j (uc - u)] and [0, a_j fv, a_j (vc - v)]
ZP_HD inline void zp_fill_mtm(ZpMat At, const ZpSums& s, const ZpCam& cam) {
    const double fu2 = cam.fu * cam.fu, fv2 = cam.fv * cam.fv;
    for (int j = 0; j < 4; j++)
        for (int k = 0; k < 4; k++) {
            int q = j <= k ? zp_pk(j, k) : zp_pk(k, j);
            double g00 = fu2 * s.s0[q], g11 = fv2 * s.s0[q], gx = cam.fu * s.sx[q], gy = cam.fv * s.sy[q], gr = s.sr[q];
            int r = 3 * j, c = 3 * k;
            At(r + 0, c + 0, 12) = g00; At(r + 0, c + 1, 12) = 0;   At(r + 0, c + 2, 12) = gx;
            At(r + 1, c + 0, 12) = 0;   At(r + 1, c + 1, 12) = g11; At(r + 1, c + 2, 12) = gy;
            At(r + 2, c + 0, 12) = gx;  At(r + 2, c + 1, 12) = gy;  At(r + 2, c + 2, 12) = gr;
        }
}

struct ZpCandidates {
    double R[3][9], t[3][3];
    bool ok[3];
};

// Everything after the sums: null space of M^T M, L/rho, three beta initialisations + Gauss-Newton, Horn alignment.
// a_first = alphas of the first correspondence (solve_for_sign looks at its camera-frame depth), pw0 = centroid.
ZP_HD inline void zp_epnp_core(ZpMat At, const ZpSums& sums, const ZpCam& cam, const ZpControl& cp,
                                const double a_first[4], const double pw0[3], ZpCandidates& out) {
    zp_fill_mtm(At, sums, cam);
    double W[12];
    zp_jacobi_rows<12>(At, W, nullptr);
    // indices of the four smallest singular values, v[0] = smallest (OpenCV sorts descending and takes rows 11..8)
    int vi[4];
    {
        bool used[12];
        for (int i = 0; i < 12; i++) used[i] = false;
        for (int q = 0; q < 4; q++) {
            int best = -1;
            for (int i = 11; i >= 0; i--)            // ties: later rows end up last after OpenCV's selection sort
                if (!used[i] && (best < 0 || W[i] < W[best])) best = i;
            used[best] = true; vi[q] = best;
        }
    }
    double vs[4];
    for (int q = 0; q < 4; q++) vs[q] = W[vi[q]] > ZP_DBL_MIN ? 1.0 / W[vi[q]] : 0.0;
#define ZPV(q, e) (At(vi[q], (e), 12) * vs[q])
    // L (6x10) and rho
    double L[60], rho[6];
    {
        const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
        for (int r = 0; r < 6; r++) {
            double dv[4][3];
            for (int q = 0; q < 4; q++)
                for (int e = 0; e < 3; e++) dv[q][e] = ZPV(q, 3 * pa[r] + e) - ZPV(q, 3 * pb[r] + e);
#define ZPD(x, y) (dv[x][0] * dv[y][0] + dv[x][1] * dv[y][1] + dv[x][2] * dv[y][2])
            double* l = L + 10 * r;
            l[0] = ZPD(0, 0); l[1] = 2 * ZPD(0, 1); l[2] = ZPD(1, 1); l[3] = 2 * ZPD(0, 2); l[4] = 2 * ZPD(1, 2);
            l[5] = ZPD(2, 2); l[6] = 2 * ZPD(0, 3); l[7] = 2 * ZPD(1, 3); l[8] = 2 * ZPD(2, 3); l[9] = ZPD(3, 3);
#undef ZPD
            double d0 = cp.cws[3 * pa[r]] - cp.cws[3 * pb[r]], d1 = cp.cws[3 * pa[r] + 1] - cp.cws[3 * pb[r] + 1],
                   d2 = cp.cws[3 * pa[r] + 2] - cp.cws[3 * pb[r] + 2];
            rho[r] = d0 * d0 + d1 * d1 + d2 * d2;
        }
    }
    for (int cand = 0; cand < 3; cand++) {
        double be[4] = {0, 0, 0, 0};
        // ---- initial betas (find_betas_approx_1/2/3)
        if (cand == 0) {
            double A[24], b[6], x[4];
            for (int r = 0; r < 6; r++) {
                A[4 * r + 0] = L[10 * r + 0]; A[4 * r + 1] = L[10 * r + 1]; A[4 * r + 2] = L[10 * r + 3];
                A[4 * r + 3] = L[10 * r + 6]; b[r] = rho[r];
            }
            zp_ls6<4>(A, b, x);
            double sgn = x[0] < 0 ? -1.0 : 1.0;
            be[0] = sqrt(sgn * x[0]);
            be[1] = sgn * x[1] / be[0]; be[2] = sgn * x[2] / be[0]; be[3] = sgn * x[3] / be[0];
        } else if (cand == 1) {
            double A[18], b[6], x[3];
            for (int r = 0; r < 6; r++) {
                A[3 * r + 0] = L[10 * r + 0]; A[3 * r + 1] = L[10 * r + 1]; A[3 * r + 2] = L[10 * r + 2]; b[r] = rho[r];
            }
            zp_ls6<3>(A, b, x);
            if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0; }
            else { be[0] = sqrt(x[0]); be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0; }
            if (x[1] < 0) be[0] = -be[0];
        } else {
            double A[30], b[6], x[5];
            for (int r = 0; r < 6; r++) {
                for (int c = 0; c < 5; c++) A[5 * r + c] = L[10 * r + c];
                b[r] = rho[r];
            }
            zp_ls6<5>(A, b, x);
            if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0; }
            else { be[0] = sqrt(x[0]); be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0; }
            if (x[1] < 0) be[0] = -be[0];
            be[2] = x[3] / be[0];
        }
        // ---- 5 Gauss-Newton steps (gauss_newton / compute_A_and_b_gauss_newton)
        for (int it = 0; it < 5; it++) {
            double A[24], b[6], x[4];
            for (int r = 0; r < 6; r++) {
                const double* l = L + 10 * r;
                A[4 * r + 0] = 2 * l[0] * be[0] + l[1] * be[1] + l[3] * be[2] + l[6] * be[3];
                A[4 * r + 1] = l[1] * be[0] + 2 * l[2] * be[1] + l[4] * be[2] + l[7] * be[3];
                A[4 * r + 2] = l[3] * be[0] + l[4] * be[1] + 2 * l[5] * be[2] + l[8] * be[3];
                A[4 * r + 3] = l[6] * be[0] + l[7] * be[1] + l[8] * be[2] + 2 * l[9] * be[3];
                b[r] = rho[r] - (l[0] * be[0] * be[0] + l[1] * be[0] * be[1] + l[2] * be[1] * be[1] +
                                 l[3] * be[0] * be[2] + l[4] * be[1] * be[2] + l[5] * be[2] * be[2] +
                                 l[6] * be[0] * be[3] + l[7] * be[1] * be[3] + l[8] * be[2] * be[3] +
                                 l[9] * be[3] * be[3]);
            }
            zp_ls6<4>(A, b, x);
            for (int q = 0; q < 4; q++) be[q] += x[q];
        }
        // ---- camera-frame control points, sign, Horn
        double ccs[12];
        for (int e = 0; e < 12; e++)
            ccs[e] = be[0] * ZPV(0, e) + be[1] * ZPV(1, e) + be[2] * ZPV(2, e) + be[3] * ZPV(3, e);
        double z_first = a_first[0] * ccs[2] + a_first[1] * ccs[5] + a_first[2] * ccs[8] + a_first[3] * ccs[11];
        if (z_first < 0)
            for (int e = 0; e < 12; e++) ccs[e] = -ccs[e];
        // pc0 = sum_j mean(alpha_j) ccs_j ; mean alphas = (s0 row sums)/n is exact only via the sums: use them
        double am[4];
        {   // sum_i a_ij = sum_k sum_i a_ij a_ik  (because sum_k a_ik = 1)
            for (int j = 0; j < 4; j++) {
                double t = 0;
                for (int k = 0; k < 4; k++) t += sums.s0[j <= k ? zp_pk(j, k) : zp_pk(k, j)];
                am[j] = t / sums.n;
            }
        }
        double pc0[3];
        for (int e = 0; e < 3; e++) pc0[e] = am[0] * ccs[e] + am[1] * ccs[3 + e] + am[2] * ccs[6 + e] + am[3] * ccs[9 + e];
        double H[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                H[3 * r + c] = ccs[r] * sums.w[c] + ccs[3 + r] * sums.w[3 + c] + ccs[6 + r] * sums.w[6 + c] +
                               ccs[9 + r] * sums.w[9 + c];
        double* R = out.R[cand];
        zp_polar3(H, R);
        double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                     R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
        for (int r = 0; r < 3; r++)
            out.t[cand][r] = pc0[r] - (R[3 * r] * pw0[0] + R[3 * r + 1] * pw0[1] + R[3 * r + 2] * pw0[2]);
        bool ok = true;
        for (int e = 0; e < 9; e++) ok = ok && isfinite(R[e]);
        for (int e = 0; e < 3; e++) ok = ok && isfinite(out.t[cand][e]);
        out.ok[cand] = ok;
    }
#undef ZPV
}

// pixel reprojection distance of one point (epnp::reprojection_error)
ZP_HD __forceinline__ double zp_reproj_dist(const double* R, const double* t, const ZpCam& cam, double X, double Y,
                                            double Z, double u, double v) {
    double Xc = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double Yc = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double iz = 1.0 / (R[6] * X + R[7] * Y + R[8] * Z + t[2]);
    double du = u - (cam.uc + cam.fu * Xc * iz), dv = v - (cam.vc + cam.fv * Yc * iz);
    return sqrt(du * du + dv * dv);
}
