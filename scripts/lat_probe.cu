// Dependent-issue latencies (cycles) of the instructions on the critical path of an MCMC step, one warp, clock64():
// DFMA, DADD, DSETP->SEL, FFMA, 64-bit SHFL.BFLY (two SHFL), SHFL+DADD round, VOTE, LDS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/lat_probe.cu -o scripts/lat_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double a, double b, float fa, float fb) {
    __shared__ double sm[64];
    double v = threadIdx.x * 1e-3 + 1.0;
    float f = threadIdx.x * 1e-3f + 1.0f;
    long long t0, t1;
    sm[threadIdx.x] = v; sm[threadIdx.x + 32] = v;
    __syncwarp();
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = fma(v, a, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = v + a;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = (v > b) ? a : b + v * 0.0;   // DSETP -> select (plus a DFMA to keep the dependence)
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, fa, fb);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = __shfl_xor_sync(0xffffffffu, v, 4, 8);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v += __shfl_xor_sync(0xffffffffu, v, 4, 8);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    t0 = clock64();
    int p = threadIdx.x & 1;
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = __any_sync(0xffffffffu, p ^ (i & 1)) ? 1 : 0;
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx & 63] & 31;
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = v + f + p + idx;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
    for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 1.0000001f, 1e-9f);
    long long h[8];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char* n[8] = {"DFMA", "DADD", "DSETP+sel(+DFMA)", "FFMA", "SHFL64", "SHFL64+DADD", "VOTE.ANY", "LDS64+cvt"};
    printf("{");
    for (int i = 0; i < 8; ++i) printf("\"%s\": %.1f%s", n[i], (double)h[i] / N, i < 7 ? ", " : "}\n");
    return 0;
}
