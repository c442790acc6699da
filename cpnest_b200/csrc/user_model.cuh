// Functor around a user-written likelihood, compiled at run time (csrc/rtc.cu) into the same chain kernels as the
// built-in functors: the device form of subclassing cpnest.model.Model with one's own log_likelihood / log_prior
// (cpnest/model.py:55-82).  Included AFTER the user's source, which must define, at global scope,
//
//     __device__ double cpn_user_log_likelihood(const double* x, int D, const double* params);
//
// and may define (announcing each with the macro)
//
//     #define CPN_USER_HAS_PRIOR      __device__ double cpn_user_log_prior(const double* x, int D, const double* params);
//     #define CPN_USER_HAS_GRADIENT   __device__ void cpn_user_grad_log_likelihood(const double* x, int D, const double* params, double* grad);
//
// x is the dense point (D doubles), params the blob handed to cpn_create.  log_prior is called for points strictly
// inside the bounds only (Model.log_prior returns -inf outside, model.py:79-80); without CPN_USER_HAS_PRIOR the prior
// is flat.  The point of a chain is distributed over G lanes; it is gathered into the chain's shared-memory scratch
// and every lane of the group evaluates the user function on it (same instructions, same result on every lane).
#pragma once
#include "group.cuh"
#include "models.cuh"

namespace cpn {

struct UserModel : ModelBase {
#ifdef CPN_USER_HAS_PRIOR
    static constexpr bool kFlatPrior = false;
#else
    static constexpr bool kFlatPrior = true;
#endif
    static constexpr bool kScratch = true;
    static constexpr bool kCheap = false;

    template <int G, int E>
    static __device__ __forceinline__ void scatter(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            if (d < m.D) m.scratch[d] = x[e];
        }
        __syncwarp();
    }
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        scatter<G, E>(g, x, m);
        return cpn_user_log_likelihood(m.scratch, m.D, m.blob);
    }
    template <int G, int E>
    static __device__ __forceinline__ double log_prior(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
#ifdef CPN_USER_HAS_PRIOR
        scatter<G, E>(g, x, m);
        return cpn_user_log_prior(m.scratch, m.D, m.blob);
#else
        return 0.0;
#endif
    }
    // central differences of a user scalar function along every coordinate (all lanes walk the same loop)
    template <int G, int E, class F>
    static __device__ __forceinline__ void finite_differences(const Group<G>& g, const double (&x)[E], const ModelCtx& m, F f,
                                                              double sign, double (&out)[E]) {
        scatter<G, E>(g, x, m);
        double* xs = m.scratch + ((m.D + 3) & ~3);         // private second row: the perturbed point
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            if (d < m.D) xs[d] = x[e];
            out[e] = 0.0;
        }
        __syncwarp();
        for (int k = 0; k < m.D; ++k) {
            const double xk = m.scratch[k];
            const double h = 1e-6 * fmax(1.0, fabs(xk));
            __syncwarp();
            if (g.lane == 0) xs[k] = xk + h;
            __syncwarp();
            const double fp = f(xs, m.D, m.blob);
            __syncwarp();
            if (g.lane == 0) xs[k] = xk - h;
            __syncwarp();
            const double fm = f(xs, m.D, m.blob);
            __syncwarp();
            if (g.lane == 0) xs[k] = xk;
            const double dk = sign * (fp - fm) / (2.0 * h);
#pragma unroll
            for (int e = 0; e < E; ++e) out[e] = (g.lane + e * G == k) ? dk : out[e];
        }
        __syncwarp();
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
#ifdef CPN_USER_HAS_GRADIENT
        scatter<G, E>(g, x, m);
        double* gs = m.scratch + ((m.D + 3) & ~3);
        cpn_user_grad_log_likelihood(m.scratch, m.D, m.blob, gs);   // every lane writes the same values
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            gr[e] = (d < m.D) ? gs[d] : 0.0;
        }
        __syncwarp();
#else
        finite_differences<G, E>(g, x, m, cpn_user_log_likelihood, 1.0, gr);
#endif
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_potential(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gv)[E]) {
#ifdef CPN_USER_HAS_PRIOR
        finite_differences<G, E>(g, x, m, cpn_user_log_prior, -1.0, gv);
#else
#pragma unroll
        for (int e = 0; e < E; ++e) gv[e] = 0.0;
#endif
    }
};

}  // namespace cpn
