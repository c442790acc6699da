// FP64 tensor-core (DMMA, mma.sync .f64) throughput and latency of the GPU this runs on, by instruction shape:
// the denominator for the block form of the correlated-Gaussian likelihood (models.cuh GaussCorr::log_likelihood_block).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/dmma_probe.cu -o scripts/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// SHAPE 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16; NACC independent accumulators per warp
template <int SHAPE, int NACC>
__global__ void k_dmma(double* out, int iters, double seed) {
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = seed - 1e-9 * (threadIdx.x + i);
    double c[NACC][4];
#pragma unroll
    for (int q = 0; q < NACC; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[q][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            if (SHAPE == 0) { double (&cc)[2] = reinterpret_cast<double (&)[2]>(c[q]); mma884(cc, a[0], b[0]); }
            if (SHAPE == 1) { mma1684(c[q], reinterpret_cast<const double (&)[2]>(a), b[0]); }
            if (SHAPE == 2) { mma1688(c[q], reinterpret_cast<const double (&)[4]>(a), reinterpret_cast<const double (&)[2]>(b)); }
            if (SHAPE == 3) { mma16816(c[q], a, b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < NACC; ++q) s += c[q][0] + c[q][1] + c[q][2] + c[q][3];
    if (s == 123456789.0) out[0] = s;
}

template <int SHAPE, int NACC>
void run(const char* name, int sms, double fma_per_inst, int warps_per_sm) {
    double* out;
    cudaMalloc(&out, sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 13;
    const int threads = 128, blocks = sms * warps_per_sm / 4;
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_dmma<SHAPE, NACC><<<blocks, threads>>>(out, iters, 0.5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inst = (double)iters * NACC * blocks * 4;            // warp instructions
    const double tf = 2.0 * fma_per_inst * inst / (best * 1e-3) / 1e12;
    const double cyc_per_inst_sm = best * 1e-3 * clk * 1e3 / ((double)iters * NACC * warps_per_sm);   // SM cycles per warp instruction
    const double cyc_per_iter_warp = best * 1e-3 * clk * 1e3 / (double)iters / NACC;                  // cycles per instruction as one warp sees it
    printf("{\"shape\": \"%s\", \"independent_accumulators\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f, \"sm_cycles_per_instruction\": %.2f, "
           "\"warp_cycles_per_instruction\": %.1f}\n", name, NACC, warps_per_sm, tf, cyc_per_inst_sm, cyc_per_iter_warp);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    // latency: one warp per SM sub-partition... one warp per SM, one dependent chain
    run<0, 1>("m8n8k4", sms, 256, 4);
    run<1, 1>("m16n8k4", sms, 512, 4);
    run<2, 1>("m16n8k8", sms, 1024, 4);
    run<3, 1>("m16n8k16", sms, 2048, 4);
    // throughput: 16 warps per SM, 4 independent accumulators each; 32 warps, 8
    run<0, 4>("m8n8k4", sms, 256, 16);
    run<1, 4>("m16n8k4", sms, 512, 16);
    run<2, 4>("m16n8k8", sms, 1024, 16);
    run<3, 4>("m16n8k16", sms, 2048, 16);
    run<0, 8>("m8n8k4", sms, 256, 32);
    run<3, 8>("m16n8k16", sms, 2048, 32);
    return 0;
}
