// EPnP in float64 as cv2.solvePnP(SOLVEPNP_EPNP) runs it (OpenCV is the un-vendored library behind the reference's
// cv2.solvePnPRansac call, CNN_output_to_pose.py:155-157).  Device restatement of the published algorithm
// (Lepetit, Moreno-Noguer, Fua 2009) with the OpenCV behaviours that change the answer under pixel noise:
// M built in pixel units with the real camera matrix (fu != fv weights the image axes), PCA control points with the
// signs of OpenCV's one-sided Jacobi SVD (same pair order and rotation rule), three beta initialisations x 5
// Gauss-Newton steps, Horn alignment with "negate row 2 when det < 0", best of three by mean reprojection distance.
// oracle/epnp.py is the CPU twin; tests/native/epnp_host.cu compiles this header for the host.
//
// The per-point work is reduced to 52 sums (ZpSums) so the same core serves the 4/5/6-point minimal solver (one
// thread per hypothesis) and the final solve on thousands of inliers (one CTA per crop, block reduction of the sums).
// The 12x12 null-space problem has two implementations: a serial one-sided Jacobi (row i cached in registers, matrix
// interleaved in shared memory) for the thread-per-hypothesis kernel, and a 16-lane cooperative one with a
// round-robin (tournament) pair order for the final solve.  The signs / order of the 12x12 singular vectors do not
// change EPnP's answer (the betas absorb them), so only the 3x3 PCA has to follow OpenCV's pair order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <cmath>

#define ZP_HD __host__ __device__
#if !defined(ZP_EIG_STAMP) || !defined(__CUDA_ARCH__)
#undef ZP_EIG_STAMP
#define ZP_EIG_STAMP(i)
#endif
#define ZP_DBL_EPS 2.220446049250313e-16
#define ZP_DBL_MIN 2.2250738585072014e-308

// strided view of an n x n double matrix (row-major); `stride` interleaves the matrices of the threads of a CTA so
// that a warp touching element (r,k) of its 32 matrices hits 32 consecutive doubles (conflict-free).
struct ZpMat {
    double* p;
    int stride;
    ZP_HD __forceinline__ double& operator()(int r, int k, int n) const { return p[(size_t)(r * n + k) * stride]; }
};

ZP_HD __forceinline__ double zp_rsqrt(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// Jacobi rotation that orthogonalises two rows with squared norms a, b and dot product p (Hestenes / OpenCV
// JacobiSVDImpl_ rule, written with two reciprocal square roots instead of hypot + 2 sqrt + 2 div):
//   beta = a - b, gamma = sqrt(4p^2 + beta^2);  beta < 0: s = sqrt((gamma-beta)/(2 gamma)), c = p/(gamma s)
//                                               else    : c = sqrt((gamma+beta)/(2 gamma)), s = p/(gamma c)
ZP_HD __forceinline__ void zp_rot(double a, double b, double p, double& c, double& s) {
    double p2 = 2 * p, beta = a - b;
    double g2 = fma(p2, p2, beta * beta);
    double ig = zp_rsqrt(g2);                    // 1/gamma
    double h = fma(0.5 * fabs(beta), ig, 0.5);   // (gamma + |beta|) / (2 gamma)  in [0.5, 1]
    double rh = zp_rsqrt(h);
    double big = h * rh;                         // sqrt(h)
    double small = p * ig * rh;                  // p / (gamma * sqrt(h))
    if (beta < 0) { s = big; c = small; } else { c = big; s = small; }
}

// Same orthogonalising rotation with the SMALL angle (|theta| <= pi/4, no implicit row swap): required for
// convergence under the parallel round-robin order of the cooperative Jacobi.
ZP_HD __forceinline__ void zp_rot_small(double a, double b, double p, double& c, double& s) {
    double p2 = 2 * p, beta = a - b;
    double g2 = fma(p2, p2, beta * beta);
    double ig = zp_rsqrt(g2);
    double h = fma(0.5 * fabs(beta), ig, 0.5);
    double rh = zp_rsqrt(h);
    c = h * rh;
    s = p * ig * rh;
    if (beta < 0) s = -s;
}

// ------------------------------------------------------------------------------------------------------------------
// serial one-sided Jacobi on the rows of At (N x N), OpenCV's cyclic pair order (i<j ascending) and stopping rule
// (eps = 10*DBL_EPSILON, max(N,30) sweeps).  Row i is held in registers across the j loop; squared norms follow the
// rotation recurrences and are recomputed exactly at the start of every sweep.  On exit W[i] = sigma_i (unsorted) and
// row i of At = sigma_i u_i^T.  G (nullable, dense row-major) accumulates the rotations (OpenCV's Vt).
// ------------------------------------------------------------------------------------------------------------------
template <int N>
ZP_HD inline void zp_jacobi_rows(ZpMat At, double* W, double* G) {
    const double eps = ZP_DBL_EPS * 10;
    if (G) {
        for (int i = 0; i < N * N; i++) G[i] = 0;
        for (int i = 0; i < N; i++) G[i * N + i] = 1;
    }
    const int max_iter = N > 30 ? N : 30;
    for (int iter = 0; iter < max_iter; iter++) {
        for (int i = 0; i < N; i++) {
            double s0 = 0, s1 = 0;
#pragma unroll
            for (int k = 0; k + 1 < N; k += 2) {
                double t0 = At(i, k, N), t1 = At(i, k + 1, N);
                s0 = fma(t0, t0, s0); s1 = fma(t1, t1, s1);
            }
            if (N & 1) { double t = At(i, N - 1, N); s0 = fma(t, t, s0); }
            W[i] = s0 + s1;
        }
        bool changed = false;
        for (int i = 0; i < N - 1; i++) {
            double ri[N];
#pragma unroll
            for (int k = 0; k < N; k++) ri[k] = At(i, k, N);
            double a = W[i];
            bool touched = false;
            for (int j = i + 1; j < N; j++) {
                double rj[N];
#pragma unroll
                for (int k = 0; k < N; k++) rj[k] = At(j, k, N);
                double p0 = 0, p1 = 0, p2 = 0;
#pragma unroll
                for (int k = 0; k < N; k += 3) {
                    p0 = fma(ri[k], rj[k], p0);
                    if (k + 1 < N) p1 = fma(ri[k + 1], rj[k + 1], p1);
                    if (k + 2 < N) p2 = fma(ri[k + 2], rj[k + 2], p2);
                }
                double p = p0 + p1 + p2, b = W[j];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                double c, s;
                zp_rot(a, b, p, c, s);
#pragma unroll
                for (int k = 0; k < N; k++) {
                    double x = ri[k], y = rj[k];
                    ri[k] = fma(c, x, s * y);
                    At(j, k, N) = fma(c, y, -s * x);
                }
                double cc = c * c, ss = s * s, cs2 = 2 * c * s * p;
                double na = fma(cc, a, fma(ss, b, cs2)), nb = fma(ss, a, fma(cc, b, -cs2));
                a = na > 0 ? na : 0;
                W[j] = nb > 0 ? nb : 0;
                changed = true; touched = true;
                if (G) {
#pragma unroll
                    for (int k = 0; k < N; k++) {
                        double x = G[i * N + k], y = G[j * N + k];
                        G[i * N + k] = fma(c, x, s * y); G[j * N + k] = fma(c, y, -s * x);
                    }
                }
            }
            if (touched) {
#pragma unroll
                for (int k = 0; k < N; k++) At(i, k, N) = ri[k];
                W[i] = a;
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

// 3x3 helpers on a private (stride 1) matrix -----------------------------------------------------------------------

// PCA of the 3x3 scatter matrix C (symmetric): singular values dc[3] descending and rows uct[3][3] with the signs
// OpenCV's SVD (U_T) returns.
ZP_HD inline void zp_pca3(const double C[9], double dc[3], double uct[9]) {
    double a[9], W[3];
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) a[i * 3 + k] = C[k * 3 + i];   // At = C^T
    ZpMat At{a, 1};
    zp_jacobi_rows<3>(At, W, nullptr);
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++) {          // selection sort descending, as OpenCV sorts W (swap on strict <)
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

struct ZpCam { double fu, fv, uc, vc; };

// The 52 sums over the points that EPnP needs once the control points are fixed.
struct ZpSums {
    // sum a_j a_k {1, x, y, x^2+y^2} with x = uc - u, y = vc - v (pixels), (j<=k) packed: 00 01 02 03 11 12 13 22 23 33
    double s0[10], sx[10], sy[10], sr[10];
    double w[12];                            // W_j = sum_i a_ij (pw_i - pw0), j = 0..3
    double n;
};

// what the pose-from-betas step needs from the sums: mean alphas and W_j (17 doubles instead of 53)
struct ZpHorn { double am[4]; double w[12]; };

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

// ------------------------------------------------------------------------------------------------------------------
// The same 52 sums from RAW MOMENTS (the split final solve, zp_finsplit.cu).  The barycentric coordinates are affine in the
// point, alpha = A [X Y Z 1]^T, so with T_f = sum_i f_i P_i P_i^T (P = [X Y Z 1], f in {1, x, y, x^2 + y^2}; 4 x 10 packed
// entries, a <= b: 00 01 02 03 11 12 13 22 23 33) the sums are contractions: sum a_j a_k f = (A T_f A^T)(j, k) and
// sum a_j (X_c - c0_c) = A_j . (T_1[:, c] - c0_c T_1[:, 3]).  One pass over the points, independent of the control points.
// Points are taken relative to a pivot (a point of the object) so that T_1 - n c c^T cancels at the object's own scale.
// ------------------------------------------------------------------------------------------------------------------
ZP_HD __forceinline__ int zp_pk4(int a, int b) {          // packed index of the symmetric 4x4, any order
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return 4 * lo - (lo * (lo - 1)) / 2 + (hi - lo);
}

// one point into the moments: T1[9] (entry 33 = the count, kept by the caller), Tx / Ty / Tr [10]
ZP_HD __forceinline__ void zp_moment_add(double* T1, double* Tx, double* Ty, double* Tr, double X, double Y, double Z,
                                         double x, double y) {
    const double r = fma(x, x, y * y);
    const double m[10] = {X * X, X * Y, X * Z, X, Y * Y, Y * Z, Y, Z * Z, Z, 1.0};
#pragma unroll
    for (int q = 0; q < 9; q++) T1[q] += m[q];
#pragma unroll
    for (int q = 0; q < 10; q++) {
        Tx[q] = fma(m[q], x, Tx[q]); Ty[q] = fma(m[q], y, Ty[q]); Tr[q] = fma(m[q], r, Tr[q]);
    }
}

// centroid c0 (relative to the pivot), control points and the affine map A [4][4] from T (T[9] = the point count n)
ZP_HD inline void zp_moment_frame(const double* T, double n, ZpControl& cp, double* A, double* c0) {
    c0[0] = T[3] / n; c0[1] = T[6] / n; c0[2] = T[8] / n;
    double C[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C[3 * r + c] = T[zp_pk4(r, c)] - c0[r] * T[zp_pk4(c, 3)];
    C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
    zp_control_points(c0, C, n, cp);
    for (int j = 0; j < 3; j++) {            // rows 1..3 from the control basis, row 0 = 1 - the others
        double off = 0;
        for (int k = 0; k < 3; k++) { A[4 * (j + 1) + k] = cp.cci[3 * j + k]; off += cp.cci[3 * j + k] * c0[k]; }
        A[4 * (j + 1) + 3] = -off;
    }
    for (int k = 0; k < 4; k++) A[k] = (k == 3 ? 1.0 : 0.0) - A[4 + k] - A[8 + k] - A[12 + k];
}

// sum number o of ZpSums (s0[10] | sx[10] | sy[10] | sr[10] | w[12]) from the moments T[40], A and c0
ZP_HD inline double zp_moment_sum(int o, const double* T, const double* A, const double* c0) {
    double val = 0;
    if (o < 40) {                                  // (A T_f A^T)(j, k)
        const int f = o / 10, q = o - 10 * f;
        const int j = q < 4 ? 0 : q < 7 ? 1 : q < 9 ? 2 : 3;
        const int k = q - (j == 0 ? 0 : j == 1 ? 4 : j == 2 ? 7 : 9) + j;
        const double* Tf = T + 10 * f;
        for (int aa = 0; aa < 4; aa++) {
            double row = 0;                        // (T_f A_k^T)[aa]
            for (int bb = 0; bb < 4; bb++) row = fma(Tf[zp_pk4(aa, bb)], A[4 * k + bb], row);
            val = fma(A[4 * j + aa], row, val);
        }
    } else {                                       // W_j[c]
        const int e = o - 40, j = e / 3, c = e - 3 * j;
        for (int aa = 0; aa < 4; aa++) val = fma(A[4 * j + aa], T[zp_pk4(aa, c)] - c0[c] * T[zp_pk4(aa, 3)], val);
    }
    return val;
}

// element (r, c) of M^T M; rows of M (epnp::fill_M): [a_j fu, 0, a_j (uc - u)] and [0, a_j fv, a_j (vc - v)]
ZP_HD __forceinline__ double zp_mtm(const ZpSums& s, const ZpCam& cam, int r, int c) {
    int j = r / 3, rr = r - 3 * j, k = c / 3, cc = c - 3 * k;
    int q = j <= k ? zp_pk(j, k) : zp_pk(k, j);
    if (rr == 0) return cc == 0 ? cam.fu * cam.fu * s.s0[q] : cc == 1 ? 0.0 : cam.fu * s.sx[q];
    if (rr == 1) return cc == 0 ? 0.0 : cc == 1 ? cam.fv * cam.fv * s.s0[q] : cam.fv * s.sy[q];
    return cc == 0 ? cam.fu * s.sx[q] : cc == 1 ? cam.fv * s.sy[q] : s.sr[q];
}

ZP_HD inline void zp_fill_mtm(ZpMat At, const ZpSums& s, const ZpCam& cam) {
    for (int r = 0; r < 12; r++)
        for (int c = 0; c < 12; c++) At(r, c, 12) = zp_mtm(s, cam, r, c);
}

// L (6x10) and rho (6) from the four null-space vectors V(q, e) (q = 0 smallest) and the control points
ZP_HD inline void zp_L_rho(ZpMat V, const ZpControl& cp, double* L, double* rho) {
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    for (int r = 0; r < 6; r++) {
        double dv[4][3];
        for (int q = 0; q < 4; q++)
            for (int e = 0; e < 3; e++) dv[q][e] = V(q, 3 * pa[r] + e, 12) - V(q, 3 * pb[r] + e, 12);
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

// One of EPnP's three candidates: beta initialisation `cand` (find_betas_approx_1/2/3), 5 Gauss-Newton steps,
// camera-frame control points, sign, Horn.  a_first = alphas of the first correspondence (solve_for_sign looks at its
// camera-frame depth), pw0 = centroid.  Returns false when the pose is not finite.
ZP_HD inline void zp_betas_init(int cand, const double* L, const double* rho, double be[4]) {
    be[0] = be[1] = be[2] = be[3] = 0;
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
}

// one Gauss-Newton system (compute_A_and_b_gauss_newton)
ZP_HD __forceinline__ void zp_gn_system(const double* L, const double* rho, const double be[4], double* A, double* b) {
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
}

ZP_HD inline void zp_horn_inputs(const ZpSums& sums, ZpHorn& h) {
    for (int j = 0; j < 4; j++) {       // mean alphas: sum_i a_ij = sum_k sum_i a_ij a_ik because sum_k a_ik = 1
        double s = 0;
        for (int k = 0; k < 4; k++) s += sums.s0[j <= k ? zp_pk(j, k) : zp_pk(k, j)];
        h.am[j] = s / sums.n;
    }
    for (int e = 0; e < 12; e++) h.w[e] = sums.w[e];
}

ZP_HD inline bool zp_pose_from_betas(const double be[4], ZpMat V, const ZpHorn& hs, const double a_first[4],
                                     const double pw0[3], double* R, double* t);

ZP_HD inline bool zp_candidate(int cand, const double* L, const double* rho, ZpMat V, const ZpHorn& hs,
                               const double a_first[4], const double pw0[3], double* R, double* t) {
    double be[4];
    zp_betas_init(cand, L, rho, be);
    for (int it = 0; it < 5; it++) {           // gauss_newton
        double A[24], b[6], x[4];
        zp_gn_system(L, rho, be, A, b);
        zp_ls6<4>(A, b, x);
        for (int q = 0; q < 4; q++) be[q] += x[q];
    }
    return zp_pose_from_betas(be, V, hs, a_first, pw0, R, t);
}

ZP_HD inline bool zp_pose_from_betas(const double be[4], ZpMat V, const ZpHorn& hs, const double a_first[4],
                                     const double pw0[3], double* R, double* t) {
    double ccs[12];
    for (int e = 0; e < 12; e++)
        ccs[e] = be[0] * V(0, e, 12) + be[1] * V(1, e, 12) + be[2] * V(2, e, 12) + be[3] * V(3, e, 12);
    double z_first = a_first[0] * ccs[2] + a_first[1] * ccs[5] + a_first[2] * ccs[8] + a_first[3] * ccs[11];
    if (z_first < 0)
        for (int e = 0; e < 12; e++) ccs[e] = -ccs[e];
    const double* am = hs.am;
    double pc0[3];
    for (int e = 0; e < 3; e++) pc0[e] = am[0] * ccs[e] + am[1] * ccs[3 + e] + am[2] * ccs[6 + e] + am[3] * ccs[9 + e];
    double H[9];        // sum_i (pc_i - pc0)(pw_i - pw0)^T = sum_j ccs_j W_j^T
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            H[3 * r + c] = ccs[r] * hs.w[c] + ccs[3 + r] * hs.w[3 + c] + ccs[6 + r] * hs.w[6 + c] + ccs[9 + r] * hs.w[9 + c];
    zp_polar3(H, R);
    double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                 R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    for (int r = 0; r < 3; r++) t[r] = pc0[r] - (R[3 * r] * pw0[0] + R[3 * r + 1] * pw0[1] + R[3 * r + 2] * pw0[2]);
    bool ok = true;
    for (int e = 0; e < 9; e++) ok = ok && isfinite(R[e]);
    for (int e = 0; e < 3; e++) ok = ok && isfinite(t[e]);
    return ok;
}

// ------------------------------------------------------------------------------------------------------------------
// Symmetric 12x12 eigen-decomposition for the EPnP null space: Householder tridiagonalisation (reflectors kept in the
// strict lower triangle, Q accumulated backwards in place) followed by the implicit-shift QL iteration on (d, e) with the
// plane rotations applied to the columns of Q.  About 10x fewer FP64 operations than the cyclic Jacobi it replaces
// (the Jacobi was 60 % of the minimal solver and 25 % of the final solve, profiles/r1d_phases.txt).  The signs / order
// of the 12x12 singular vectors do not change EPnP's answer, so only the 3x3 problems keep OpenCV's Jacobi.
//
// G lanes work on one matrix (G = 1: plain serial code, also what the host build runs; G = 4: a quad per hypothesis
// in the minimal solver; G = 16: half a warp in the final solve).  Lane gl owns the columns (tridiagonalisation,
// accumulation) resp. rows (QL) congruent to gl mod G; scalar recurrences are computed redundantly by every lane from
// the shared d/e arrays, and zp_gsync (a __syncwarp over the group's lanes) orders shared-memory writes against the
// other lanes' reads.  The matrix is stored with a row stride of 13 doubles: with that, and matrices 4 (mod 16)
// doubles apart, both the row- and the column-ownership access patterns of the 8 quads of a warp are conflict free.
// ------------------------------------------------------------------------------------------------------------------
#define ZP_RS 13
#define ZP_SYM_DOUBLES (12 * ZP_RS)

struct ZpSym12 {
    double* p;
    ZP_HD __forceinline__ double& operator()(int r, int c) const { return p[r * ZP_RS + c]; }
};

ZP_HD __forceinline__ void zp_gsync(unsigned mask) {
#ifdef __CUDA_ARCH__
    __syncwarp(mask);
#else
    (void)mask;
#endif
}
ZP_HD __forceinline__ bool zp_any(unsigned wmask, bool pred) {
#ifdef __CUDA_ARCH__
    return __any_sync(wmask, pred) != 0;
#else
    (void)wmask;
    return pred;
#endif
}

// z: symmetric matrix in, eigenvectors (columns) out; d[12]: eigenvalues (unsorted); e[12]: scratch.
// Householder tridiagonalisation T = Q^T A Q, Q = H_0 ... H_9: on exit d[0..11] / e[0..10] hold the diagonal / subdiagonal
// of T and column k of z, below the diagonal, the reflector u_k (|u_k|^2 = 2, H_k = I - u_k u_k^T; u_k = 0 if none).
template <int G>
ZP_HD inline void zp_tridiag12(ZpSym12 z, double* d, double* e, int gl, unsigned mask) {
    constexpr int N = 12;
    // ---- Householder tridiagonalisation: step k annihilates z(k+2.., k); u_k (|u|^2 = 2, H = I - u u^T) is left in
    //      column k below the diagonal
    for (int k = 0; k < N - 2; k++) {
        const int b0 = k + 1;
        double x0 = z(b0, k), sig = 0;
        for (int i = b0 + 1; i < N; i++) { double t = z(i, k); sig = fma(t, t, sig); }
        const double tot = fma(x0, x0, sig), dk = z(k, k);
        const bool refl = sig > 0 && tot < 1.7e308;            // false for an already tridiagonal column and for NaN/inf
        double alpha = x0, scale = 0, v0 = 0;
        if (refl) {
            double nrm = sqrt(tot);
            alpha = x0 > 0 ? -nrm : nrm;
            v0 = x0 - alpha;
            scale = zp_rsqrt(tot - alpha * x0);
        }
        zp_gsync(mask);                                        // everybody has read column k
        for (int i = b0 + gl; i < N; i += G) z(i, k) = (i == b0 ? v0 : z(i, k)) * scale;
        if (gl == 0) { d[k] = dk; e[k] = alpha; }
        zp_gsync(mask);
        if (refl) {
            for (int c = b0 + gl; c < N; c += G) {             // p = A22 u (A22 symmetric: column c = row c)
                double s0 = 0, s1 = 0;
                int r = b0;
                for (; r + 1 < N; r += 2) { s0 = fma(z(r, c), z(r, k), s0); s1 = fma(z(r + 1, c), z(r + 1, k), s1); }
                if (r < N) s0 = fma(z(r, c), z(r, k), s0);
                e[c] = s0 + s1;
            }
            zp_gsync(mask);
            double K = 0;
            for (int c = b0; c < N; c++) K = fma(z(c, k), e[c], K);
            K *= 0.5;
            for (int c = b0 + gl; c < N; c += G) {             // A22 -= u w^T + w u^T with w = p - K u
                const double uc = z(c, k), wc = fma(-K, uc, e[c]);
                for (int r = b0; r < N; r++) {
                    const double ur = z(r, k), wr = fma(-K, ur, e[r]);
                    z(r, c) -= fma(ur, wc, wr * uc);
                }
            }
        }
        zp_gsync(mask);
    }
    {
        const double d10 = z(N - 2, N - 2), d11 = z(N - 1, N - 1), e10 = z(N - 1, N - 2);
        zp_gsync(mask);
        if (gl == 0) { d[N - 2] = d10; d[N - 1] = d11; e[N - 2] = e10; e[N - 1] = 0; }
    }
}

// z: symmetric matrix in, eigenvectors (columns) out; d[12]: eigenvalues (unsorted); e[12]: scratch.  Full decomposition
// (tridiagonalisation + accumulated Q + implicit QL with vectors): reference implementation for the host tests; the
// kernels use zp_smallest4_12 below, which needs only 4 eigenpairs.
template <int G>
ZP_HD inline void zp_symeig12(ZpSym12 z, double* d, double* e, int gl, unsigned mask, unsigned wmask) {
    constexpr int N = 12;
    zp_tridiag12<G>(z, d, e, gl, mask);
    // ---- Q = H_0 H_1 ... H_9 accumulated backwards in place (column j always belongs to lane j mod G)
    if ((N - 1) % G == gl) z(N - 1, N - 1) = 1;
    for (int k = N - 3; k >= 0; k--) {
        const int b0 = k + 1;
        for (int j = b0 + ((gl - b0) % G + G) % G; j < N; j += G) {
            const double ub = z(b0, k);
            if (j == b0) {
                z(b0, b0) = fma(-ub, ub, 1.0);
                for (int i = b0 + 1; i < N; i++) z(i, b0) = -ub * z(i, k);
            } else {
                double s = 0;
                for (int i = b0 + 1; i < N; i++) s = fma(z(i, k), z(i, j), s);
                z(b0, j) = -s * ub;
                for (int i = b0 + 1; i < N; i++) z(i, j) = fma(-s, z(i, k), z(i, j));
            }
        }
        zp_gsync(mask);                                        // column k (u_k) is overwritten in the next step
    }
    for (int j = gl; j < N; j += G) {
        if (j == 0) { z(0, 0) = 1; for (int i = 1; i < N; i++) z(i, 0) = 0; }
        else z(0, j) = 0;
    }
    zp_gsync(mask);
    // ---- implicit-shift QL on (d, e).  Every lane keeps a private copy of d and e and its own rows of Q (row r belongs
    //      to lane r mod G) in REGISTERS: the l and i loops are fully unrolled so that every index is static, the scalar
    //      recurrence is computed redundantly by the lanes of a group and nothing is exchanged until the end.  (The first
    //      version kept d, e and Q in shared memory: 100 issue slots per rotation, half of them address arithmetic,
    //      LDS/STS and the group barrier that ordered the shared d/e updates.)
    //      The control flow is uniform over the whole warp (wmask = every lane that entered this call): all groups walk
    //      the same l / sweep / i loops and a group that has nothing to do at a position is predicated off, because
    //      data-dependent loop bounds would make the 8 quads of a warp diverge and run one after the other.
    constexpr int ROWS = (N + G - 1) / G;
    double dd[N], ee[N], zr[ROWS][N];
#pragma unroll
    for (int i = 0; i < N; i++) { dd[i] = d[i]; ee[i] = e[i]; }
#pragma unroll
    for (int k = 0; k < ROWS; k++) {
        const int r = gl + k * G;
#pragma unroll
        for (int c = 0; c < N; c++) zr[k][c] = r < N ? z(r, c) : 0.0;
    }
    double anorm = 0;
#pragma unroll
    for (int i = 0; i < N; i++) anorm = fmax(anorm, fabs(dd[i]) + fabs(ee[i]));
    const double small = anorm * ZP_DBL_EPS;
    const bool finite = anorm < 1.7e308;                      // a non-finite matrix is left alone
#pragma unroll
    for (int l = 0; l < N - 1; l++) {
        for (int iter = 0; iter < 40; iter++) {
            int m = N - 1;
            double dm = dd[N - 1];
#pragma unroll
            for (int j = N - 2; j >= l; j--) {                // first negligible e[j], j >= l
                const bool neg = fabs(ee[j]) <= small;
                m = neg ? j : m;
                dm = neg ? dd[j] : dm;
            }
            const bool active = m != l && finite;
            if (!zp_any(wmask, active)) break;
            const double dl = dd[l], el = ee[l];
            double g = (dd[l + 1] - dl) / (2 * (active ? el : 1.0));
            double r = sqrt(fma(g, g, 1.0));
            g = dm - dl + el / (g + (g >= 0 ? r : -r));       // d[m] - shift
            double s = 1, c = 1, p = 0;
            bool live = active;                               // false after an underflow recovery
#pragma unroll
            for (int i = N - 2; i >= l; i--) {
                const bool on = live && i < m;
                const double f = s * ee[i], b = c * ee[i];
                const double r2 = fma(f, f, g * g);
                const bool ok = r2 > 0;                        // false: underflow (or NaN) -> deflate here, end this sweep
                const double ir = zp_rsqrt(ok ? r2 : 1.0);
                const double rn = r2 * ir, sn = f * ir, cn = g * ir;
                const double gn = dd[i + 1] - p;
                const double t = fma(dd[i] - gn, sn, 2 * cn * b);
                const double pn = sn * t;
                if (on && ok) {
                    ee[i + 1] = rn; dd[i + 1] = gn + pn;
                    s = sn; c = cn; p = pn; g = fma(cn, t, -b);
#pragma unroll
                    for (int k = 0; k < ROWS; k++) {
                        const double z1 = zr[k][i + 1], z0 = zr[k][i];
                        zr[k][i + 1] = fma(sn, z0, cn * z1);
                        zr[k][i] = fma(cn, z0, -sn * z1);
                    }
                } else if (on) {
                    ee[i + 1] = 0; dd[i + 1] -= p;
#pragma unroll
                    for (int j = l; j < N - 1; j++) if (j == m) ee[j] = 0;
                    live = false;
                }
            }
            if (live) {
                dd[l] = dl - p; ee[l] = g;
#pragma unroll
                for (int j = l; j < N - 1; j++) if (j == m) ee[j] = 0;
            }
        }
    }
    zp_gsync(mask);                                           // every lane has loaded its rows before anybody stores
#pragma unroll
    for (int k = 0; k < ROWS; k++) {
        const int r = gl + k * G;
        if (r < N) {
#pragma unroll
            for (int c = 0; c < N; c++) z(r, c) = zr[k][c];
        }
    }
    if (gl == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) d[i] = dd[i];
    }
    zp_gsync(mask);
}

// ------------------------------------------------------------------------------------------------------------------
// The four smallest eigenpairs of the 12x12 matrix -- all EPnP needs -- without the QL iteration: after the
// tridiagonalisation, "eigen-lane" k (k = 0..3: the first four lanes of the group, or a loop on the host) finds the k-th
// smallest eigenvalue of T by bisection on the Sturm sequence (division-free determinant recurrence on T / |T|, 46
// halvings: fixed trip count, no divergence), then its eigenvector by inverse iteration on T - lambda I (tridiagonal LU
// with partial pivoting, 3 solves, modified Gram-Schmidt against the eigen-lanes below it after every solve, which is
// what makes a multiple eigenvalue -- the 2-dimensional null space of a 5-point M^T M -- come out as an orthonormal
// basis), and finally multiplies by the Householder reflectors (Q is never formed).  ~2.5 k issue slots against ~20 k for
// accumulating Q and running QL with vectors, and the chain of dependent operations is 10x shorter.
// ------------------------------------------------------------------------------------------------------------------
struct ZpTriLU { double dl[11], dg[12], du[11], du2[10]; unsigned piv; };

// number of eigenvalues of the (normalised) tridiagonal (td, e2 = squared off-diagonals) that are < x
ZP_HD __forceinline__ unsigned zp_signbit(double v) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(v) >> 31;
#else
    return std::signbit(v) ? 1u : 0u;
#endif
}

ZP_HD __forceinline__ int zp_sturm_count(const double* td, const double* e2, double x) {
    // p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}: one dependent DFMA per step (the product with p_{i-2} is off the chain);
    // the count is the number of sign changes along 1, p_1, .., p_12.  Signs are taken from the sign BIT: a p_i that is
    // exactly zero is followed by -e^2 p_{i-2}, so whichever sign the zero carries, p_{i-2} -> p_i -> p_{i+1} shows exactly
    // one change, as it must.  With |T| = 1 the p_i stay far from overflow; they can only become tiny next to a multiple
    // eigenvalue, so the pair is rescaled once, half way.
    double pm1 = 1.0, p0 = td[0] - x;
    unsigned signs = zp_signbit(p0);                     // bit i (from the top of the 12 collected) = sign of p_{i+1}
#pragma unroll
    for (int i = 1; i < 12; i++) {
        const double pn = fma(td[i] - x, p0, -e2[i - 1] * pm1);
        signs = (signs << 1) | zp_signbit(pn);
        pm1 = p0; p0 = pn;
        if (i == 6 && fabs(p0) < 1e-100) { p0 *= 1e150; pm1 *= 1e150; }
    }
    // signs = s_1 .. s_12 (s_12 in bit 0), s_0 = 0: changes = popcount(s ^ (s >> 1)) over the 12 adjacent pairs
    const unsigned ch = (signs ^ (signs >> 1)) & 0xFFFu;
#ifdef __CUDA_ARCH__
    return __popc(ch);
#else
    return __builtin_popcount(ch);
#endif
}

ZP_HD __forceinline__ double zp_bisect_kth(const double* td, const double* e2, int k) {
    double lo = -1.01, hi = 1.01;                         // |T| is normalised: every eigenvalue lies in [-1, 1]
    for (int it = 0; it < 46; it++) {                     // 2.02 * 2^-46 = 3e-14 |T|: ample for the inverse iteration
        const double mid = 0.5 * (lo + hi);
        const bool below = zp_sturm_count(td, e2, mid) > k;    // more than k eigenvalues < mid -> the k-th is below mid
        hi = below ? mid : hi;
        lo = below ? lo : mid;
    }
    return 0.5 * (lo + hi);
}

// LU with partial pivoting of the tridiagonal T - lam I (sub/super-diagonals te)
ZP_HD __forceinline__ void zp_tri_lu(const double* td, const double* te, double lam, ZpTriLU& f) {
    const double tiny = 1e-290;
#pragma unroll
    for (int i = 0; i < 12; i++) f.dg[i] = td[i] - lam;
#pragma unroll
    for (int i = 0; i < 11; i++) { f.dl[i] = te[i]; f.du[i] = te[i]; }
#pragma unroll
    for (int i = 0; i < 10; i++) f.du2[i] = 0;
    f.piv = 0;
#pragma unroll
    for (int i = 0; i < 11; i++) {
        if (fabs(f.dg[i]) >= fabs(f.dl[i])) {
            if (f.dg[i] == 0) f.dg[i] = tiny;
            const double m = f.dl[i] / f.dg[i];
            f.dl[i] = m;
            f.dg[i + 1] = fma(-m, f.du[i], f.dg[i + 1]);
        } else {                                         // interchange rows i and i + 1
            const double m = f.dg[i] / f.dl[i];
            f.dg[i] = f.dl[i];
            f.dl[i] = m;
            const double t = f.du[i];
            f.du[i] = f.dg[i + 1];
            f.dg[i + 1] = fma(-m, f.dg[i + 1], t);
            if (i < 10) { f.du2[i] = f.du[i + 1]; f.du[i + 1] = -m * f.du[i + 1]; }
            f.piv |= 1u << i;
        }
    }
    if (f.dg[11] == 0) f.dg[11] = tiny;
#pragma unroll
    for (int i = 0; i < 12; i++) f.dg[i] = 1.0 / f.dg[i];  // the solves multiply
}

ZP_HD __forceinline__ void zp_tri_solve(const ZpTriLU& f, double* x) {
#pragma unroll
    for (int i = 0; i < 11; i++) {
        if ((f.piv >> i) & 1u) { const double t = x[i]; x[i] = x[i + 1]; x[i + 1] = fma(-f.dl[i], x[i], t); }
        else x[i + 1] = fma(-f.dl[i], x[i], x[i + 1]);
    }
    x[11] *= f.dg[11];
    x[10] = fma(-f.du[10], x[11], x[10]) * f.dg[10];
#pragma unroll
    for (int i = 9; i >= 0; i--) x[i] = fma(-f.du2[i], x[i + 2], fma(-f.du[i], x[i + 1], x[i])) * f.dg[i];
}

// x <- x / |x|.  safe = true first brings the entries to O(1) with an exact power-of-two factor (a solve against a nearly
// singular matrix grows them by up to 1e16 or more)
ZP_HD __forceinline__ void zp_normalise12(double* x, bool safe) {
    if (safe) {
        double a = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) a = fmax(a, fabs(x[i]));
        double sc = 1.0;
        if (a > 0 && a < 1.7e308) {
#ifdef __CUDA_ARCH__
            const int ex = (__double2hiint(a) >> 20) & 0x7ff;                 // biased exponent of the largest entry
            sc = __hiloint2double((2046 - (ex < 1 ? 1 : ex > 2045 ? 2045 : ex)) << 20, 0);
#else
            int ex;
            frexp(a, &ex);
            sc = ldexp(1.0, 1 - ex);
#endif
        }
#pragma unroll
        for (int i = 0; i < 12; i++) x[i] *= sc;
    }
    double n0 = 0, n1 = 0;
#pragma unroll
    for (int i = 0; i < 12; i += 2) { n0 = fma(x[i], x[i], n0); n1 = fma(x[i + 1], x[i + 1], n1); }
    const double n2 = n0 + n1;
    const double inv = n2 > 0 ? zp_rsqrt(n2) : 0.0;
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] *= inv;
}

// x <- Q x with Q = H_0 ... H_9 (reflectors in the columns of z)
ZP_HD __forceinline__ void zp_apply_reflectors(ZpSym12 z, double* x) {
#pragma unroll
    for (int k = 9; k >= 0; k--) {
        double u[12], s0 = 0, s1 = 0;
#pragma unroll
        for (int i = k + 1; i < 12; i++) u[i] = z(i, k);
#pragma unroll
        for (int i = k + 1; i < 12; i += 2) { s0 = fma(u[i], x[i], s0); if (i + 1 < 12) s1 = fma(u[i + 1], x[i + 1], s1); }
        const double s = s0 + s1;
#pragma unroll
        for (int i = k + 1; i < 12; i++) x[i] = fma(-s, u[i], x[i]);
    }
}

ZP_HD __forceinline__ double zp_bcast(double v, int src, int width, unsigned mask) {
#ifdef __CUDA_ARCH__
    return __shfl_sync(mask, v, src, width);
#else
    (void)src; (void)width; (void)mask;
    return v;
#endif
}

// z: symmetric matrix in (destroyed); V[v * 12 + i]: eigenvector of the v-th smallest eigenvalue; lam4 (nullable): the four
// eigenvalues.  d, e: 12 doubles each of scratch shared by the group.  All G lanes of the group call it together
// (G >= 4 on the device; G = 1 = serial host code).
template <int G>
ZP_HD inline void zp_smallest4_12(ZpSym12 z, double* d, double* e, int gl, unsigned mask, double* V, double* lam4) {
    ZP_EIG_STAMP(16);
    zp_tridiag12<G>(z, d, e, gl, mask);
    zp_gsync(mask);
    ZP_EIG_STAMP(17);
    double td[12], te[11], e2[11];
    double anorm = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) anorm = fmax(anorm, fabs(d[i]) + (i < 11 ? fabs(e[i]) : 0.0) + (i > 0 ? fabs(e[i - 1]) : 0.0));
    const double inv = anorm > 0 && anorm < 1.7e308 ? 1.0 / anorm : 1.0;
#pragma unroll
    for (int i = 0; i < 12; i++) td[i] = d[i] * inv;
#pragma unroll
    // the Sturm recurrence needs an unreduced matrix: an off-diagonal of at least 1e-15 |T| (a change of the
    // eigenvalues below rounding) keeps it from decoupling into exact zeros
    for (int i = 0; i < 11; i++) { te[i] = e[i] * inv; e2[i] = fmax(te[i] * te[i], 1e-30); }
    constexpr int NL = G == 1 ? 4 : 1;                    // eigen-lanes handled by this thread
    double lam[4], x[NL][12];
    // ---- eigenvalues
#ifdef __CUDA_ARCH__
    {
        const double mine = zp_bisect_kth(td, e2, gl < 4 ? gl : 3);
#pragma unroll
        for (int k = 0; k < 4; k++) lam[k] = zp_bcast(mine, k, G, mask);
    }
#else
    for (int k = 0; k < 4; k++) lam[k] = zp_bisect_kth(td, e2, k);
#endif
    if (lam4) { if (gl == 0) for (int k = 0; k < 4; k++) lam4[k] = lam[k] * anorm; }
    ZP_EIG_STAMP(18);
    // separate coinciding eigenvalues a little so that the shifted matrices differ (LAPACK dstein does the same)
    double lp[4];
    lp[0] = lam[0];
#pragma unroll
    for (int k = 1; k < 4; k++) lp[k] = fmax(lam[k], lp[k - 1] + 1e-14);
    // ---- inverse iteration
#pragma unroll
    for (int q = 0; q < NL; q++) {
        const int k = G == 1 ? q : (gl < 4 ? gl : 3);
#pragma unroll
        for (int i = 0; i < 12; i++) x[q][i] = 1.0 + 0.25 * (double)((i * 5 + k * 3) % 7);
    }
#ifdef __CUDA_ARCH__
    ZpTriLU f;
    zp_tri_lu(td, te, lp[gl < 4 ? gl : 3], f);
    for (int it = 0; it < 3; it++) {
        zp_tri_solve(f, x[0]);
        zp_normalise12(x[0], true);
        // modified Gram-Schmidt down the eigen-lanes: the lanes above j remove their component along x_j and renormalise
        // (every lane runs the same instructions; only the lanes above j keep the result)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double xj[12], dot0 = 0, dot1 = 0, dot2 = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) xj[i] = zp_bcast(x[0][i], j, G, mask);
#pragma unroll
            for (int i = 0; i < 12; i += 3) { dot0 = fma(xj[i], x[0][i], dot0); dot1 = fma(xj[i + 1], x[0][i + 1], dot1); dot2 = fma(xj[i + 2], x[0][i + 2], dot2); }
            const double dot = dot0 + dot1 + dot2;
            const double w = gl > j && gl < 4 ? dot : 0.0;
#pragma unroll
            for (int i = 0; i < 12; i++) x[0][i] = fma(-w, xj[i], x[0][i]);
            zp_normalise12(x[0], false);
        }
    }
    zp_apply_reflectors(z, x[0]);
    if (gl < 4) {
#pragma unroll
        for (int i = 0; i < 12; i++) V[gl * 12 + i] = x[0][i];
    }
#else
    ZpTriLU f[4];
    for (int k = 0; k < 4; k++) zp_tri_lu(td, te, lp[k], f[k]);
    for (int it = 0; it < 3; it++) {
        for (int k = 0; k < 4; k++) zp_tri_solve(f[k], x[k]);
        for (int k = 0; k < 4; k++) zp_normalise12(x[k], true);
        for (int j = 0; j < 3; j++) {
            for (int k = j + 1; k < 4; k++) {
                double dot = 0;
                for (int i = 0; i < 12; i++) dot = fma(x[j][i], x[k][i], dot);
                for (int i = 0; i < 12; i++) x[k][i] = fma(-dot, x[j][i], x[k][i]);
                zp_normalise12(x[k], false);
            }
        }
    }
    for (int k = 0; k < 4; k++) {
        zp_apply_reflectors(z, x[k]);
        for (int i = 0; i < 12; i++) V[k * 12 + i] = x[k][i];
    }
#endif
    zp_gsync(mask);
    ZP_EIG_STAMP(19);
}

// Null space for EPnP: fills z with M^T M (lane gl its own columns) from the 40 sums S (memory every lane of the group
// can index dynamically: shared memory on the device) and writes the four eigenvectors of the smallest eigenvalues to
// V[v * 12 + e] (v = 0 smallest).  d, e: 12 doubles each, shared by the group.
template <int G>
ZP_HD inline void zp_nullspace4(ZpSym12 z, double* d, double* e, const double* S, const ZpCam& cam, int gl,
                                unsigned mask, double* V) {
    // column c = 3k + cc of M^T M; its 3x3 blocks are [fu^2 s0, 0, fu sx; 0, fv^2 s0, fv sy; fu sx, fv sy, sr](j,k)
    for (int c = gl; c < 12; c += G) {
        const int k = c / 3, cc = c - 3 * k;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int lo = j < k ? j : k, hi = j < k ? k : j;
            const int q = 4 * lo - (lo * (lo - 1)) / 2 + (hi - lo);         // packed index of the symmetric 4x4: 0 4 7 9
            const double s0 = S[q], sx = S[10 + q], sy = S[20 + q], sr = S[30 + q];
            z(3 * j + 0, c) = cc == 0 ? cam.fu * cam.fu * s0 : cc == 1 ? 0.0 : cam.fu * sx;
            z(3 * j + 1, c) = cc == 0 ? 0.0 : cc == 1 ? cam.fv * cam.fv * s0 : cam.fv * sy;
            z(3 * j + 2, c) = cc == 0 ? cam.fu * sx : cc == 1 ? cam.fv * sy : sr;
        }
    }
    zp_gsync(mask);
    zp_smallest4_12<G>(z, d, e, gl, mask, V, nullptr);
}

// Null space with the QL eigen-solver: fills z with M^T M (lane gl its own columns), decomposes, and writes the four
// eigenvectors of the smallest eigenvalues to V[v * 12 + e] (v = 0 smallest).  d, e: 12 doubles each, shared by the
// group.  All lanes of the group must call it together.
template <int G>
ZP_HD inline void zp_nullspace_ql(ZpSym12 z, double* d, double* e, const ZpSums& sums, const ZpCam& cam, int gl,
                                  unsigned mask, unsigned wmask, double* V) {
    for (int c = gl; c < 12; c += G)
        for (int r = 0; r < 12; r++) z(r, c) = zp_mtm(sums, cam, r, c);
    zp_gsync(mask);
    zp_symeig12<G>(z, d, e, gl, mask, wmask);
    int vi[4];
    unsigned used = 0;
    for (int q = 0; q < 4; q++) {
        int best = -1;
        double bw = 0;
        for (int i = 0; i < 12; i++) {
            double w = d[i];
            if (!((used >> i) & 1u) && (best < 0 || w < bw)) { best = i; bw = w; }
        }
        used |= 1u << best;
        vi[q] = best;
    }
    for (int q = 0; q < 4; q++)
        for (int r = gl; r < 12; r += G) V[q * 12 + r] = z(r, vi[q]);
    zp_gsync(mask);
}

// serial null space: fills At with M^T M, runs the Jacobi, and leaves the four normalised singular vectors of the
// smallest singular values in rows 0..3 of At (row 0 = smallest), i.e. At doubles as the V view afterwards.
ZP_HD inline void zp_nullspace_serial(ZpMat At, const ZpSums& sums, const ZpCam& cam) {
    zp_fill_mtm(At, sums, cam);
    double W[12];
    zp_jacobi_rows<12>(At, W, nullptr);
    int vi[4];
    bool used[12];
    for (int i = 0; i < 12; i++) used[i] = false;
    for (int q = 0; q < 4; q++) {
        int best = -1;
        for (int i = 11; i >= 0; i--)
            if (!used[i] && (best < 0 || W[i] < W[best])) best = i;
        used[best] = true; vi[q] = best;
    }
    double v[48];
    for (int q = 0; q < 4; q++) {
        double s = W[vi[q]] > ZP_DBL_MIN ? 1.0 / W[vi[q]] : 0.0;
        for (int e = 0; e < 12; e++) v[q * 12 + e] = At(vi[q], e, 12) * s;
    }
    for (int q = 0; q < 4; q++)
        for (int e = 0; e < 12; e++) At(q, e, 12) = v[q * 12 + e];
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

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------------------------
// Group-cooperative one-sided Jacobi for the 12x12 problem: G lanes per problem (G = 4: 8 problems per warp, each lane
// owns 3 columns of At; G = 16: 2 problems per warp, lanes 0..11 own one column each).  a[r][c] = At[r][col0 + c].
// Six disjoint row pairs are rotated per round in a round-robin (tournament) order, 11 rounds per sweep.  Per round:
//   - the six dot products: local FMAs + a log2(G)-step butterfly (shuffles);
//   - the rotation (small angle) and the new squared norms of pair t are computed by ONE owner lane (t mod G) -- the
//     FP64 pipe issues one warp instruction every 2 cycles, so redundant rotation math was the bottleneck -- and
//     broadcast with shuffles;
//   - every lane rotates its own columns; the slots are permuted by register moves.
// Norms are recomputed exactly every sweep.  All 32 lanes of the warp must call it together; the sweep loop runs until
// every problem of the warp has converged (extra sweeps are identity rotations).
// On exit W[r] = sigma_r (uniform in the group) and a[r][*] = this lane's columns of sigma_r u_r^T.
// ------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ double zp_groupsum(double x) {
#pragma unroll
    for (int d = 1; d < G; d <<= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}

template <int G>
__device__ inline void zp_jacobi12_group(double (*a)[(G == 4) ? 3 : 1], double* W, int gl /* lane in group */) {
    constexpr int CPL = (G == 4) ? 3 : 1;
    constexpr int PASSES = (6 + G - 1) / G;
    const double eps2 = (ZP_DBL_EPS * 10) * (ZP_DBL_EPS * 10);
    for (int sweep = 0; sweep < 30; sweep++) {
#pragma unroll
        for (int r = 0; r < 12; r++) {
            double v = 0;
#pragma unroll
            for (int k = 0; k < CPL; k++) v = fma(a[r][k], a[r][k], v);
            W[r] = zp_groupsum<G>(v);
        }
        bool changed = false;
#pragma unroll 1
        for (int round = 0; round < 11; round++) {
            double p[6];
#pragma unroll
            for (int t = 0; t < 6; t++) {
                double v = 0;
#pragma unroll
                for (int k = 0; k < CPL; k++) v = fma(a[2 * t][k], a[2 * t + 1][k], v);
                p[t] = zp_groupsum<G>(v);
            }
            double rc[PASSES], rs[PASSES], ra[PASSES], rb[PASSES];
#pragma unroll
            for (int ps = 0; ps < PASSES; ps++) {
                // this lane owns pair t = ps*G + gl (if < 6): pick its inputs with a select chain
                double pq = p[ps * G], A = W[2 * ps * G], Bn = W[2 * ps * G + 1];
#pragma unroll
                for (int t = ps * G + 1; t < 6 && t < (ps + 1) * G; t++)
                    if (gl == t - ps * G) { pq = p[t]; A = W[2 * t]; Bn = W[2 * t + 1]; }
                bool mine = ps * G + gl < 6;
                bool rot = mine && pq * pq > eps2 * A * Bn;
                double cc, ss;
                zp_rot_small(A, Bn, rot ? pq : 1.0, cc, ss);
                cc = rot ? cc : 1.0;
                ss = rot ? ss : 0.0;
                changed |= rot;
                double c2 = cc * cc, s2 = ss * ss, cs2 = 2 * cc * ss * pq;
                double na = fma(c2, A, fma(s2, Bn, cs2)), nb = fma(s2, A, fma(c2, Bn, -cs2));
                rc[ps] = cc; rs[ps] = ss; ra[ps] = na > 0 ? na : 0; rb[ps] = nb > 0 ? nb : 0;
            }
#pragma unroll
            for (int t = 0; t < 6; t++) {
                const int ps = t / G, src = t % G;
                double c = __shfl_sync(0xffffffffu, rc[ps], src, G), s = __shfl_sync(0xffffffffu, rs[ps], src, G);
                W[2 * t] = __shfl_sync(0xffffffffu, ra[ps], src, G);
                W[2 * t + 1] = __shfl_sync(0xffffffffu, rb[ps], src, G);
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    double x = a[2 * t][k], y = a[2 * t + 1][k];
                    a[2 * t][k] = fma(c, x, s * y);
                    a[2 * t + 1][k] = fma(c, y, -s * x);
                }
            }
            // tournament rotation of the slots (slot 0 fixed): bot0 -> top1 -> ... -> top5 -> bot5 -> ... -> bot1 -> bot0
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                double t1 = a[1][k];
                a[1][k] = a[3][k]; a[3][k] = a[5][k]; a[5][k] = a[7][k]; a[7][k] = a[9][k]; a[9][k] = a[11][k];
                a[11][k] = a[10][k]; a[10][k] = a[8][k]; a[8][k] = a[6][k]; a[6][k] = a[4][k]; a[4][k] = a[2][k];
                a[2][k] = t1;
            }
            {
                double w1 = W[1];
                W[1] = W[3]; W[3] = W[5]; W[5] = W[7]; W[7] = W[9]; W[9] = W[11];
                W[11] = W[10]; W[10] = W[8]; W[8] = W[6]; W[6] = W[4]; W[4] = W[2];
                W[2] = w1;
            }
        }
        if (!__any_sync(0xffffffffu, changed)) break;
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        double v = 0;
#pragma unroll
        for (int k = 0; k < CPL; k++) v = fma(a[r][k], a[r][k], v);
        W[r] = sqrt(zp_groupsum<G>(v));
    }
}

// Null space by a group of G lanes: builds this lane's columns of M^T M from the sums, runs the Jacobi and writes the
// four normalised singular vectors of the smallest singular values to V (v = 0 smallest): V[v * 12 + e], 48 doubles.
template <int G>
__device__ inline void zp_nullspace_group(const ZpSums& sums, const ZpCam& cam, int gl, double* V) {
    constexpr int CPL = (G == 4) ? 3 : 1;
    double a[12][CPL], W[12];
    const bool has_cols = gl * CPL < 12;
#pragma unroll
    for (int r = 0; r < 12; r++)
#pragma unroll
        for (int c = 0; c < CPL; c++) a[r][c] = has_cols ? zp_mtm(sums, cam, r, gl * CPL + c) : 0.0;
    zp_jacobi12_group<G>(a, W, gl);
    bool used[12];
#pragma unroll
    for (int r = 0; r < 12; r++) used[r] = false;
    for (int v = 0; v < 4; v++) {
        int bi = -1;
        double bw = 0, bv[CPL];
#pragma unroll
        for (int r = 11; r >= 0; r--)
            if (!used[r] && (bi < 0 || W[r] < bw)) {
                bi = r; bw = W[r];
#pragma unroll
                for (int c = 0; c < CPL; c++) bv[c] = a[r][c];
            }
#pragma unroll
        for (int r = 0; r < 12; r++) used[r] = used[r] || r == bi;
        double inv = bw > ZP_DBL_MIN ? 1.0 / bw : 0.0;
        if (has_cols)
#pragma unroll
            for (int c = 0; c < CPL; c++) V[v * 12 + gl * CPL + c] = bv[c] * inv;
    }
}
#endif
