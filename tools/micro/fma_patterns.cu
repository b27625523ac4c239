// Microbenchmark: issue cost of packed FFMA2 vs scalar FFMA for different numbers of distinct register operands.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_patterns fma_patterns.cu && ./fma_patterns
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 pk(float lo, float hi) { return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32); }

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    f32x2 pa = pk(a, a), pb = pk(b, b);
    f32x2 x[8], y[8], w[8];
    for (int i = 0; i < 8; i++) { x[i] = pk(threadIdx.x + i, 1.f + i); y[i] = pk(0.999f - 0.001f * i, 0.998f); w[i] = pk(0.001f * i, 0.002f); }
    float xs[16], ys[16], ws[16];
    for (int i = 0; i < 16; i++) { xs[i] = threadIdx.x + i; ys[i] = 0.999f - 0.001f * i; ws[i] = 0.001f * i; }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { for (int i = 0; i < 8; i++) x[i] = fma2(x[i], pa, pb); }                 // 1 distinct
        if (MODE == 1) { for (int i = 0; i < 8; i++) x[i] = fma2(x[i], y[i], pb); }               // 2 distinct
        if (MODE == 2) { for (int i = 0; i < 8; i++) x[i] = fma2(x[i], y[i], w[i]); }             // 3 distinct
        if (MODE == 3) { for (int i = 0; i < 8; i++) x[i] = fma2(pa, y[i], x[i]); }               // shared multiplier, 2 distinct (acc form)
        if (MODE == 4) { for (int i = 0; i < 16; i++) xs[i] = fmaf(xs[i], a, b); }                // scalar 1 distinct
        if (MODE == 5) { for (int i = 0; i < 16; i++) xs[i] = fmaf(xs[i], ys[i], b); }            // scalar 2 distinct
        if (MODE == 6) { for (int i = 0; i < 16; i++) xs[i] = fmaf(xs[i], ys[i], ws[i]); }        // scalar 3 distinct
        if (MODE == 7) { for (int i = 0; i < 16; i++) xs[i] = fmaf(a, ys[i], xs[i]); }            // scalar shared multiplier
    }
    f32x2 s = 0; for (int i = 0; i < 8; i++) s ^= x[i];
    float t = 0; for (int i = 0; i < 16; i++) t += xs[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s ^ (unsigned)(s >> 32)) + t;
}
template <int MODE> void run(const char* name) {
    int blocks = 148 * 8, threads = 256, iters = 20000;
    float* d; cudaMalloc(&d, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; r++) { cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double flops = 2.0 * 16 * (double)iters * blocks * threads;   // 16 scalar-equivalent FMAs per iteration in every mode
    printf("%-44s %7.3f ms  %6.1f TFLOP/s\n", name, best, flops / best / 1e9);
    cudaFree(d);
}
int main() {
    run<0>("FFMA2 x=fma2(x,A,B)      1 distinct");
    run<1>("FFMA2 x=fma2(x,y,B)      2 distinct");
    run<2>("FFMA2 x=fma2(x,y,w)      3 distinct");
    run<3>("FFMA2 x=fma2(A,y,x)      2 distinct (acc)");
    run<4>("FFMA  x=fma(x,a,b)       1 distinct");
    run<5>("FFMA  x=fma(x,y,b)       2 distinct");
    run<6>("FFMA  x=fma(x,y,w)       3 distinct");
    run<7>("FFMA  x=fma(a,y,x)       2 distinct (acc)");
    return 0;
}
