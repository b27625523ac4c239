// Host build of zp_symeig12 (zebrapose_b200/csrc/zp_epnp.cuh): stdin = 144 doubles per matrix (row-major, symmetric),
// stdout = 12 eigenvalues followed by the 144 entries of the eigenvector matrix (row-major, eigenvectors in columns).
#include <cstdio>
#include "../../zebrapose_b200/csrc/zp_epnp.cuh"

int main() {
    double a[144];
    for (;;) {
        for (int i = 0; i < 144; i++) if (scanf("%lf", &a[i]) != 1) return 0;
        double zb[ZP_SYM_DOUBLES], d[12], e[12];
        ZpSym12 z{zb};
        for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) z(r, c) = a[r * 12 + c];
        zp_symeig12<1>(z, d, e, 0, 0u, 0u);
        for (int i = 0; i < 12; i++) printf("%.17g ", d[i]);
        for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) printf("%.17g ", z(r, c));
        printf("\n");
    }
}
