// Dependent-chain latencies of the fp64 operations the homography solver is made of (one warp, sm_100a).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o scripts/_build/fp64_probe scripts/fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed, int n) {
    __shared__ double sm[64 * 32];
    double x = seed + threadIdx.x;
    long long t[8];
    t[0] = clock64();
    for (int i = 0; i < n; ++i) x = x * 1.0000001 + 0.5;           // DMUL + DADD (no contraction)
    t[1] = clock64();
    for (int i = 0; i < n; ++i) x = __fma_rn(x, 1.0000001, 0.5);   // DFMA
    t[2] = clock64();
    for (int i = 0; i < n; ++i) x = 3.0 / x + 1.0;                 // DDIV (+DADD)
    t[3] = clock64();
    for (int i = 0; i < n; ++i) x = sqrt(x) + 2.0;                 // DSQRT (+DADD)
    t[4] = clock64();
    sm[threadIdx.x] = x;
    int idx = threadIdx.x;
    for (int i = 0; i < n; ++i) { double v = sm[idx]; idx = ((int)v & 31) ^ (idx & 31); sm[idx] = v + 1.0; }   // LDS -> cvt -> STS chain
    t[5] = clock64();
    double mv = 0; int best = 0;
    for (int i = 0; i < n; ++i) { double v = fabs(x + i); if (mv < v) { mv = v; best = i; } x = mv * 0.999; }  // compare-select chain
    t[6] = clock64();
    out[threadIdx.x] = x + idx + best;
    if (threadIdx.x == 0) for (int j = 0; j < 6; ++j) cyc[j] = t[j + 1] - t[j];
}
int main() {
    double *out; long long *cyc, h[6];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 6 * 8);
    const int n = 4096;
    for (int threads = 1; threads <= 32; threads *= 32) {
        k<<<1, threads>>>(out, cyc, 1.5, n);
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %2d: dmul+dadd %.1f  dfma %.1f  ddiv+dadd %.1f  dsqrt+dadd %.1f  lds-sts chain %.1f  cmp-select %.1f cycles/iter\n", threads,
               (double)h[0] / n, (double)h[1] / n, (double)h[2] / n, (double)h[3] / n, (double)h[4] / n, (double)h[5] / n);
    }
    return 0;
}
