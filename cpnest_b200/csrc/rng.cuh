// Counter-based RNG for the chain kernels: Philox4x32-10 (Salmon et al. 2011).
// Replaces the reference's per-process, unseeded MT19937 (`random` in cpnest/proposal.py:6-7,
// cpnest/sampler.py:9).  counter = (chain id lo, chain id hi, step, block), key = run seed:
// a chain's random stream does not depend on the grid shape or on which GPU evolves it.
#pragma once
#include "sysinc.cuh"

namespace cpn {

struct Philox4 { uint32_t v[4]; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += W0;
        k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// uniform on (0,1): never 0 or 1, so log(u) is finite and < 0
__host__ __device__ __forceinline__ double u32_to_unit_open(uint32_t r) {
    return ((double)r + 0.5) * (1.0 / 4294967296.0);
}
// uniform index in [0,n)
__host__ __device__ __forceinline__ uint32_t u32_to_index(uint32_t r, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __umulhi(r, n);
#else
    return (uint32_t)(((uint64_t)r * n) >> 32);
#endif
}

#ifdef __CUDACC__
// two N(0,1) from two u32 (Box-Muller, fp32 special-function units).  The proposals only need
// symmetric unit-variance draws; 24-bit resolution is ample.
__device__ __forceinline__ void box_muller_f(uint32_t r0, uint32_t r1, float& g0, float& g1) {
    float u1 = ((float)(r0 >> 8) + 0.5f) * (1.0f / 16777216.0f);           // (0,1)
    float th = ((float)(r1 >> 8) + 0.5f) * (2.0f / 16777216.0f) - 1.0f;    // (-1,1) -> angle pi*th
    float rad = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(3.14159265358979f * th, &s, &c);
    g0 = rad * c;
    g1 = rad * s;
}
__device__ __forceinline__ void box_muller(uint32_t r0, uint32_t r1, double& g0, double& g1) {
    float a, b;
    box_muller_f(r0, r1, a, b);
    g0 = (double)a;
    g1 = (double)b;
}
#endif

}  // namespace cpn
