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

// d = (x - u z)^2 + (y - v z)^2 - z^2 with [x y z] = P [X Y Z 1] and u, v already divided by thr; the point is an inlier
// iff d < 0, i.e. iff the SIGN BIT of d is set (14 FP32-pipe instructions, no compare; explicit fmaf so every kernel
// rounds identically).
__device__ __forceinline__ float zp_inlier_d(const float4& p0, const float4& p1, const float4& p2, float u, float v,
                                             float X, float Y, float Z) {
    float x = fmaf(p0.x, X, fmaf(p0.y, Y, fmaf(p0.z, Z, p0.w)));
    float y = fmaf(p1.x, X, fmaf(p1.y, Y, fmaf(p1.z, Z, p1.w)));
    float z = fmaf(p2.x, X, fmaf(p2.y, Y, fmaf(p2.z, Z, p2.w)));
    float dx = fmaf(-u, z, x);
    float dy = fmaf(-v, z, y);
    float e = fmaf(dx, dx, __fmul_rn(dy, dy));
    return fmaf(-z, z, e);                         // same roundings as the packed (FFMA2) form in zp_score_kernel
}

// The inlier decision exactly as cv2's PnPRansacCallback::computeError makes it (projectPoints in double without
// distortion -> float32 image point -> float32 squared distance <= float32(thr^2)); every operation spelled out, no
// contraction.  Used where ONE hypothesis per crop is evaluated (the winner's final inlier set) for the points the
// float32 division-free predicate puts within 1e-3 px of the threshold -- the slack north_star allows the scoring
// kernel, removed where it is free to remove.  (cv2 rebuilds R from its Rodrigues vector: 1e-16 relative, 1e-13 px.)
__device__ __forceinline__ bool zp_inlier_exact(const double* pose, double fx, double fy, double cx, double cy, float u,
                                                float v, float Xf, float Yf, float Zf, float thr2) {
    const double X = Xf, Y = Yf, Z = Zf;
    double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pose[0], X), __dmul_rn(pose[1], Y)), __dmul_rn(pose[2], Z)), pose[9]);
    double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pose[3], X), __dmul_rn(pose[4], Y)), __dmul_rn(pose[5], Z)), pose[10]);
    double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pose[6], X), __dmul_rn(pose[7], Y)), __dmul_rn(pose[8], Z)), pose[11]);
    z = z != 0.0 ? __ddiv_rn(1.0, z) : 1.0;
    x = __dmul_rn(x, z); y = __dmul_rn(y, z);
    const float px = (float)__dadd_rn(__dmul_rn(x, fx), cx), py = (float)__dadd_rn(__dmul_rn(y, fy), cy);
    const float dx = __fsub_rn(u, px), dy = __fsub_rn(v, py);
    const float err = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    return err <= thr2;
}
