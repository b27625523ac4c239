// Projection matrices and the inlier predicate shared by the minimal solvers (write P), scoring and the final solve.
// Every operation is spelled out (fma / __fmul_rn), so the values do not depend on the -fmad setting of the translation
// unit that includes this header (zp_cvsolve.cu is compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// P = diag(1/thr, 1/thr, 1) K [R|t] evaluated in double, rounded once to float32.  Folding the threshold into rows 0,1
// (and into u, v: see zp_inlier_d) turns the test into (x - u z)^2 + (y - v z)^2 <= z^2.  A non-finite pose gives P = 0,
// for which the predicate is false for every point.
__device__ __forceinline__ void zp_make_P(const double* pose, const double* K, double inv_thr, float P[12]) {
    const double fx = K[0], sk = K[1], cx = K[2], fy = K[4], cy = K[5];
    bool fin = true;
#pragma unroll
    for (int e = 0; e < 12; e++) fin = fin && isfinite(pose[e]);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        double r0 = c < 3 ? pose[c] : pose[9], r1 = c < 3 ? pose[3 + c] : pose[10], r2 = c < 3 ? pose[6 + c] : pose[11];
        P[c] = fin ? (float)__dmul_rn(fma(fx, r0, fma(sk, r1, __dmul_rn(cx, r2))), inv_thr) : 0.f;
        P[4 + c] = fin ? (float)__dmul_rn(fma(fy, r1, __dmul_rn(cy, r2)), inv_thr) : 0.f;
        P[8 + c] = fin ? (float)r2 : 0.f;
    }
}
