// Dependent-chain latencies of the FP64 operations the exact EPnP replay is made of (one warp, one thread measured):
// cycles per DADD, DMUL, DFMA, division, square root, shared-memory load.  nvcc -arch=sm_100a -fmad=false -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double a, double b, long long* out, double* sink) {
    __shared__ double sm[64];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sm[i] = (double)((i * 7 + 1) % 64);
    __syncthreads();
    const int N = 2048;
    double x = a + threadIdx.x;
    long long t0, t1;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dadd_rn(x, b);
    t1 = clock64(); if (threadIdx.x == 0) out[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dmul_rn(x, b);
    t1 = clock64(); if (threadIdx.x == 0) out[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, b, a);
    t1 = clock64(); if (threadIdx.x == 0) out[2] = t1 - t0;
    x = a + 3.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = b / x + 1.5;
    t1 = clock64(); if (threadIdx.x == 0) out[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = sqrt(x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) out[4] = t1 - t0;
    int idx = threadIdx.x & 63;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = (int)sm[idx];
    t1 = clock64(); if (threadIdx.x == 0) out[5] = t1 - t0;
    // 3 interleaved 12-term sequential sums (the pair step's p, a, b)
    double ri[12], rj[12];
    for (int k = 0; k < 12; k++) { ri[k] = sm[k] + x; rj[k] = sm[k + 12] + b; }
    t0 = clock64();
    double acc = 0;
    for (int rep = 0; rep < 64; rep++) {
        double p = 0, aa = 0, bb = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) { p += ri[k] * rj[k]; aa += ri[k] * ri[k]; bb += rj[k] * rj[k]; }
        acc += p + aa + bb; ri[rep % 12] = acc;
    }
    t1 = clock64(); if (threadIdx.x == 0) out[6] = t1 - t0;
    sink[threadIdx.x] = x + idx + acc;
}
int main() {
    long long* d; double* s; cudaMalloc(&d, 64); cudaMalloc(&s, 8 * 1024);
    for (int threads : {32, 128, 512}) {
        for (int rep = 0; rep < 2; rep++) lat<<<1, threads>>>(1.0000001, 0.99999, d, s);
        long long h[8]; cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
        printf("threads %d  cycles/op: dadd %.1f dmul %.1f dfma %.1f div(+add) %.1f sqrt(+add) %.1f lds %.1f  3x12 sums %.1f per triple\n", threads,
               h[0] / 2048.0, h[1] / 2048.0, h[2] / 2048.0, h[3] / 2048.0, h[4] / 2048.0, h[5] / 2048.0, h[6] / 64.0);
    }
    return 0;
}
