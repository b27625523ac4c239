// Host build of the device EPnP core (zebrapose_b200/csrc/zp_epnp.cuh) so the CPU test-suite can check the very code
// the kernels run against cv2 / the oracle.  stdin: n, mode, K(9), then n rows "X Y Z u v"; stdout: 12 doubles.
// mode 1: the 52 EPnP sums accumulated point by point (zp_final_cl_kernel); mode 2: from 40 raw moments relative to the
// first point and their contractions (zp_finsplit.cu: zp_moment_add / zp_moment_frame / zp_moment_sum).
#include <cstdio>
#include <vector>
#include "../../zebrapose_b200/csrc/zp_epnp.cuh"

int main() {
    int n, f32;
    double K[9];
    while (scanf("%d %d", &n, &f32) == 2) {
        for (int i = 0; i < 9; i++) if (scanf("%lf", &K[i]) != 1) return 1;
        std::vector<double> X(n), Y(n), Z(n), x(n), y(n);
        ZpCam cam{K[0], K[4], K[2], K[5]};
        double c0[3] = {0, 0, 0};
        for (int i = 0; i < n; i++) {
            double u, v;
            if (scanf("%lf %lf %lf %lf %lf", &X[i], &Y[i], &Z[i], &u, &v) != 5) return 1;
            x[i] = u; y[i] = v;
            c0[0] += X[i]; c0[1] += Y[i]; c0[2] += Z[i];
        }
        for (int e = 0; e < 3; e++) c0[e] /= n;
        double C[9] = {0};
        for (int i = 0; i < n; i++) {
            double d[3] = {X[i] - c0[0], Y[i] - c0[1], Z[i] - c0[2]};
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) C[3 * r + c] += d[r] * d[c];
        }
        ZpControl cp;
        ZpSums s;
        double af[4];
        if (f32 == 2) {
            const double g[3] = {X[0], Y[0], Z[0]};
            double T[40], A[16], c0p[3];
            for (int q = 0; q < 40; q++) T[q] = 0;
            for (int i = 0; i < n; i++)
                zp_moment_add(T, T + 10, T + 20, T + 30, X[i] - g[0], Y[i] - g[1], Z[i] - g[2], cam.uc - x[i], cam.vc - y[i]);
            T[9] = n;
            zp_moment_frame(T, (double)n, cp, A, c0p);
            double* S = (double*)&s;
            for (int o = 0; o < 52; o++) S[o] = zp_moment_sum(o, T, A, c0p);
            s.n = n;
            zp_alphas(cp, 0.0, 0.0, 0.0, af);                 // the first point is the pivot
            for (int e = 0; e < 3; e++) c0[e] = c0p[e] + g[e];   // t = pc0 - R c0 wants the world centroid
        } else {
            zp_control_points(c0, C, (double)n, cp);
            for (int q = 0; q < 10; q++) { s.s0[q] = s.sx[q] = s.sy[q] = s.sr[q] = 0; }
            for (int q = 0; q < 12; q++) s.w[q] = 0;
            s.n = n;
            for (int i = 0; i < n; i++) {
                double a[4];
                zp_alphas(cp, X[i], Y[i], Z[i], a);
                if (i == 0) for (int e = 0; e < 4; e++) af[e] = a[e];
                zp_accumulate(s, a, cam.uc - x[i], cam.vc - y[i], X[i] - c0[0], Y[i] - c0[1], Z[i] - c0[2]);
            }
        }
        double zbuf[ZP_SYM_DOUBLES], dd[12], ee[12], at[48];
        zp_nullspace4<1>(ZpSym12{zbuf}, dd, ee, s.s0, cam, 0, 0u, at);
        ZpMat At{at, 1};
        double L[60], rho[6];
        zp_L_rho(At, cp, L, rho);
        ZpHorn hs;
        zp_horn_inputs(s, hs);
        int best = -1; double be = 0, Rb[9], tb[3];
        for (int c = 0; c < 3; c++) {
            double R[9], t[3];
            if (!zp_candidate(c, L, rho, At, hs, af, c0, R, t)) continue;
            double e = 0;
            for (int i = 0; i < n; i++) e += zp_reproj_dist(R, t, cam, X[i], Y[i], Z[i], x[i], y[i]);
            e /= n;
            if (!(e == e)) continue;
            if (best < 0 || e < be) { best = c; be = e; for (int q = 0; q < 9; q++) Rb[q] = R[q]; for (int q = 0; q < 3; q++) tb[q] = t[q]; }
        }
        if (best < 0) { for (int e = 0; e < 12; e++) printf("nan "); printf("\n"); continue; }
        for (int e = 0; e < 9; e++) printf("%.17g ", Rb[e]);
        for (int e = 0; e < 3; e++) printf("%.17g ", tb[e]);
        printf("\n");
    }
    return 0;
}
