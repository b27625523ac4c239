// Microbenchmark of the scoring inner loop: packed FFMA2 (as zp_score_kernel) vs scalar FFMA, P broadcast from shared memory,
// 8 correspondences per thread in registers, 128 threads per CTA, 5 CTAs per SM.  Reports algorithmic TFLOP/s (27 flop / eval).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 pk(float lo, float hi) { return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32); }
constexpr int H = 160, PPT = 8;

template <int MODE>
__global__ void __launch_bounds__(128) k(int* out, int reps, const float* gP) {
    __shared__ __align__(16) float sP[H * 24];
    for (int i = threadIdx.x; i < H * 24; i += 128) sP[i] = gP[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float u[PPT], v[PPT], X[PPT], Y[PPT], Z[PPT];
    for (int j = 0; j < PPT; j++) { u[j] = -(0.1f * threadIdx.x + j); v[j] = -(0.2f * threadIdx.x - j); X[j] = 0.01f * (threadIdx.x + 3 * j); Y[j] = 0.02f * (threadIdx.x - j); Z[j] = 0.03f * j + 0.5f; }
    f32x2 nu2[PPT / 2], nv2[PPT / 2], X2[PPT / 2], Y2[PPT / 2], Z2[PPT / 2];
    for (int k2 = 0; k2 < PPT / 2; k2++) { nu2[k2] = pk(u[2 * k2], u[2 * k2 + 1]); nv2[k2] = pk(v[2 * k2], v[2 * k2 + 1]); X2[k2] = pk(X[2 * k2], X[2 * k2 + 1]); Y2[k2] = pk(Y[2 * k2], Y[2 * k2 + 1]); Z2[k2] = pk(Z[2 * k2], Z[2 * k2 + 1]); }
    int total = 0;
    for (int r = 0; r < reps; r++) {
        for (int iq = 0; iq < H; iq += 32) {
            int acc = 0;
#pragma unroll 2
            for (int il = 0; il < 32; il++) {
                uint32_t bits = 0;
                if (MODE == 0) {
                    const ulonglong2* pp = (const ulonglong2*)(sP + 24 * (iq + il));
                    const ulonglong2 q0 = pp[0], q1 = pp[1], q2 = pp[2], q3 = pp[3], q4 = pp[4], q5 = pp[5];
                    constexpr int NP = PPT / 2;
                    f32x2 x[NP], y[NP], z[NP];
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) z[k2] = fma2(q5.x, Z2[k2], q5.y);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) z[k2] = fma2(q4.y, Y2[k2], z[k2]);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) z[k2] = fma2(q4.x, X2[k2], z[k2]);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) x[k2] = fma2(q1.x, Z2[k2], q1.y);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) x[k2] = fma2(q0.y, Y2[k2], x[k2]);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) x[k2] = fma2(q0.x, X2[k2], x[k2]);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) y[k2] = fma2(q3.x, Z2[k2], q3.y);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) y[k2] = fma2(q2.y, Y2[k2], y[k2]);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) y[k2] = fma2(q2.x, X2[k2], y[k2]);
#pragma unroll
                    for (int k2 = 0; k2 < NP; k2++) {
                        f32x2 dx = fma2(nu2[k2], z[k2], x[k2]), dy = fma2(nv2[k2], z[k2], y[k2]);
                        f32x2 e = fma2(dx, dx, mul2(dy, dy));
                        f32x2 d = fma2(z[k2] ^ 0x8000000080000000ull, z[k2], e);
                        bits = __funnelshift_l((uint32_t)d, bits, 1);
                        bits = __funnelshift_l((uint32_t)(d >> 32), bits, 1);
                    }
                } else {
                    const float4* pp = (const float4*)(sP + 12 * (iq + il));      // scalar layout: 12 floats per hypothesis
                    const float4 p0 = pp[0], p1 = pp[1], p2 = pp[2];
                    float x[PPT], y[PPT], z[PPT];
#pragma unroll
                    for (int j = 0; j < PPT; j++) z[j] = fmaf(p2.z, Z[j], p2.w);
#pragma unroll
                    for (int j = 0; j < PPT; j++) z[j] = fmaf(p2.y, Y[j], z[j]);
#pragma unroll
                    for (int j = 0; j < PPT; j++) z[j] = fmaf(p2.x, X[j], z[j]);
#pragma unroll
                    for (int j = 0; j < PPT; j++) x[j] = fmaf(p0.z, Z[j], p0.w);
#pragma unroll
                    for (int j = 0; j < PPT; j++) x[j] = fmaf(p0.y, Y[j], x[j]);
#pragma unroll
                    for (int j = 0; j < PPT; j++) x[j] = fmaf(p0.x, X[j], x[j]);
#pragma unroll
                    for (int j = 0; j < PPT; j++) y[j] = fmaf(p1.z, Z[j], p1.w);
#pragma unroll
                    for (int j = 0; j < PPT; j++) y[j] = fmaf(p1.y, Y[j], y[j]);
#pragma unroll
                    for (int j = 0; j < PPT; j++) y[j] = fmaf(p1.x, X[j], y[j]);
#pragma unroll
                    for (int j = 0; j < PPT; j++) {
                        float dx = fmaf(u[j], z[j], x[j]), dy = fmaf(v[j], z[j], y[j]);
                        float e = fmaf(dx, dx, dy * dy);
                        float d = fmaf(-z[j], z[j], e);
                        bits = __funnelshift_l(__float_as_uint(d), bits, 1);
                    }
                }
                int c = MODE == 2 ? __popc(bits) : __reduce_add_sync(0xffffffffu, __popc(bits));
                if (MODE == 2) acc += c; else if (lane == il) acc += c;
            }
            total += acc;
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = total;
}
template <int MODE> void run(const char* name, int ctas_per_sm) {
    int blocks = 148 * ctas_per_sm, reps = 40;
    int* d; cudaMalloc(&d, blocks * 128 * 4);
    float* gP; cudaMalloc(&gP, H * 24 * 4);
    float hP[H * 24]; for (int i = 0; i < H * 24; i++) hP[i] = 0.001f * (i % 97) - 0.04f;
    cudaMemcpy(gP, hP, sizeof(hP), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; r++) { cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(d, reps, gP); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double evals = (double)blocks * 128 * PPT * H * reps;
    printf("%-34s %d CTA/SM  %7.3f ms  %6.1f TFLOP/s (27 flop/eval)\n", name, ctas_per_sm, best, 27.0 * evals / best / 1e9);
    cudaFree(d); cudaFree(gP);
}
int main() {
    for (int c : {2, 4, 5, 8}) { run<0>("packed FFMA2 + REDUX", c); run<1>("scalar FFMA + REDUX", c); run<2>("scalar FFMA, no REDUX", c); }
    return 0;
}
