// Single-warp issue-rate probes (cycles per warp-instruction with 8 independent streams):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/issue_bench tools/issue_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(double* out, long long* cyc, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, s = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, s);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    int v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], (threadIdx.x + 1 + i) & 31);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 4 independent 64-bit shuffles each feeding a DFMA, interleaved with 4 independent DFMAs
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double x = __shfl_sync(0xffffffffu, a[i], (threadIdx.x + 1 + i) & 31);
            a[i] = fma(x, m, s);
            a[4 + i] = fma(a[4 + i], m, s);
        }
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // DMUL throughput
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] * m;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    double r = 0; int q = 0;
    for (int i = 0; i < 8; ++i) { r += a[i]; q += v[i]; }
    out[threadIdx.x] = r + q;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 4 * 8);
    const int iters = 4096;
    k<<<1, 32>>>(out, cyc, iters);
    long long h[4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("cycles per warp-instruction, 8 independent streams, one warp: DFMA %.2f  SHFL32 %.2f  (4x shfl64->DFMA + 4 DFMA per iter: %.1f cycles/iter)  DMUL %.2f\n",
           h[0] / (8.0 * iters), h[1] / (8.0 * iters), h[2] / (double)iters, h[3] / (8.0 * iters));
    return 0;
}
