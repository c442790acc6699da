// Lane groups: a chain is evolved by G lanes of a warp (G in {1,2,4,8,16,32}); each lane holds
// E coordinates, coordinate d lives in lane d % G, slot d / G (so a point row in HBM is read
// with consecutive lanes on consecutive doubles).  All shuffles are issued by the full warp.
#pragma once
#include "sysinc.cuh"

namespace cpn {

constexpr unsigned FULL = 0xffffffffu;

template <int G>
struct Group {
    static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "lane group size");
    int lane;          // lane within the group
    unsigned gmask;    // bits of this group's lanes in the warp
    __device__ __forceinline__ Group() {
        int wl = threadIdx.x & 31;
        lane = wl & (G - 1);
        gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (wl & ~(G - 1)));
    }
    // butterfly sum: every lane ends with the bit-identical total
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int m = G >> 1; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m, G);
        return v;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int m = G >> 1; m > 0; m >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, m, G));
        return v;
    }
    __device__ __forceinline__ double min(double v) const {
#pragma unroll
        for (int m = G >> 1; m > 0; m >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, m, G));
        return v;
    }
    __device__ __forceinline__ double prod(double v) const {
#pragma unroll
        for (int m = G >> 1; m > 0; m >>= 1) v *= __shfl_xor_sync(FULL, v, m, G);
        return v;
    }
    __device__ __forceinline__ bool all(bool p) const {
        if (G == 1) return p;
        unsigned b = __ballot_sync(FULL, p);
        return (b & gmask) == gmask;
    }
    __device__ __forceinline__ double bcast(double v, int src_lane) const {
        if (G == 1) return v;
        return __shfl_sync(FULL, v, src_lane, G);
    }
};

// coordinate d of a point distributed as x[e] <-> d = lane + e*G, fetched by every lane.
template <int G, int E>
__device__ __forceinline__ double coord(const Group<G>& g, const double (&x)[E], int d) {
    int e = d / G, src = d % G;
    double v = x[0];
#pragma unroll
    for (int k = 1; k < E; ++k) v = (e == k) ? x[k] : v;
    return g.bcast(v, src);
}

}  // namespace cpn
