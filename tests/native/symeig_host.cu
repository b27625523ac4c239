// Host build of the 12x12 eigen-solvers of zebrapose_b200/csrc/zp_epnp.cuh.  stdin = 144 doubles per matrix (row-major,
// symmetric).  Mode "full" (default): stdout = 12 eigenvalues + the 144 entries of the eigenvector matrix (row-major,
// eigenvectors in columns) from zp_symeig12.  Mode "small4" (argv[1]): 4 eigenvalues + 4 x 12 eigenvectors (rows) from
// zp_smallest4_12, the routine the kernels run.
#include <cstdio>
#include <cstring>
#include "../../zebrapose_b200/csrc/zp_epnp.cuh"

int main(int argc, char** argv) {
    const bool small4 = argc > 1 && !strcmp(argv[1], "small4");
    double a[144];
    for (;;) {
        for (int i = 0; i < 144; i++) if (scanf("%lf", &a[i]) != 1) return 0;
        double zb[ZP_SYM_DOUBLES], d[12], e[12];
        ZpSym12 z{zb};
        for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) z(r, c) = a[r * 12 + c];
        if (small4) {
            double V[48], lam[4];
            zp_smallest4_12<1>(z, d, e, 0, 0u, V, lam);
            for (int i = 0; i < 4; i++) printf("%.17g ", lam[i]);
            for (int i = 0; i < 48; i++) printf("%.17g ", V[i]);
        } else {
            zp_symeig12<1>(z, d, e, 0, 0u, 0u);
            for (int i = 0; i < 12; i++) printf("%.17g ", d[i]);
            for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) printf("%.17g ", z(r, c));
        }
        printf("\n");
    }
}
