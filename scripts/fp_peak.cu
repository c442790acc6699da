// Measured FP64 (DFMA) and FP32 (FFMA) pipe peaks of the GPU this runs on: the denominators for
// the "FP64 pipe utilisation" figure of the chain kernels (MEASURED_PEAKS.json has only HBM and
// bf16 numbers).  Eight independent FMA chains per thread, 1024 threads x 8 blocks per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/fp_peak.cu -o scripts/fp_peak
#include <cstdio>
#include <cuda_runtime.h>

template <class T>
__global__ void k_fma(T* out, int iters, T a, T b) {
    T v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == (T)123456789) out[0] = s;   // never true: keeps the loop alive
}

template <class T>
double run(const char* name, int sms, int iters) {
    T* out;
    cudaMalloc(&out, sizeof(T));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 2, threads = 1024;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_fma<T><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(out);
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double f64 = run<double>("fp64", p.multiProcessorCount, 1 << 14);
    const double f32 = run<float>("fp32", p.multiProcessorCount, 1 << 17);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz\": %d, \"l2_bytes\": %d, \"fp64_fma_tflops\": %.3f, "
           "\"fp32_fma_tflops\": %.3f, \"how\": \"8 independent FMA chains per thread, 2x1024 threads per SM, best of 5, CUDA events\"}\n",
           p.name, p.multiProcessorCount, clk, p.l2CacheSize, f64, f32);
    return 0;
}
