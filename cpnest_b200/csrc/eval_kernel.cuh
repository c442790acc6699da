// Model.new_point / log_prior / log_likelihood on a set of points (cpnest/model.py:35-82),
// used by the live-set initialisation (Sampler.reset, cpnest/sampler.py:116-126) and by the
// functor parity entry point.
#pragma once
#include "sysinc.cuh"
#include "group.cuh"
#include "models.cuh"
#include "rng.cuh"

namespace cpn {

struct EvalParams {
    double* x;         // [n][Dp]  in (draw == 0) / out (draw == 1)
    double* logL;
    double* logP;
    int n, D, Dp;
    const double* lo;
    const double* hi;
    const double* blob;
    int draw;          // 1: draw uniformly in the box until log_prior is finite (model.py:44-52)
    double* grad_logL; // [n][Dp] or null: gradient of log_likelihood (reflection normal of the HMC kernel)
    double* grad_V;    // [n][Dp] or null: Model.force, gradient of the potential -log_prior
    uint32_t seed_lo, seed_hi;
    unsigned long long id_base;
};

constexpr int kEvalThreads = 128;

template <int G, int E, class Model>
__global__ void __launch_bounds__(kEvalThreads) k_eval_points(const EvalParams p) {
    const Group<G> g;
    const int cpb = kEvalThreads / G;
    int c = blockIdx.x * cpb + (int)threadIdx.x / G;
    if ((c & ~(32 / G - 1)) >= p.n) return;
    const bool valid = c < p.n;
    if (!valid) c = p.n - 1;
    extern __shared__ double smem[];
    ModelCtx m;
    m.blob = p.blob;
    m.D = p.D;
    m.scratch = Model::kScratch ? smem + ((int)threadIdx.x / G) * kScratchRows * p.Dp : nullptr;
    Model::prepare(m);
    set_valid_mask<G, E>(g, m);
    const int D = p.D;
    double x[E];
    double logP = CPN_NEG_INF;
    if (p.draw) {
        const unsigned long long id = p.id_base + (unsigned long long)c;
        bool need = true;
        for (int attempt = 0; attempt < 100000; ++attempt) {
            if (!__any_sync(FULL, need)) break;
            double xt[E];
            bool inb = true;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                xt[e] = 0.0;
                if (d < D) {
                    // 53-bit uniform from two words of block (attempt, d)
                    const Philox4 r = philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)attempt,
                                                    0x80000000u | (uint32_t)d, p.seed_lo, p.seed_hi);
                    const unsigned long long bits = (((unsigned long long)r.v[0] << 32) | r.v[1]) >> 11;
                    const double u = ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
                    const double lo = p.lo[d], hi = p.hi[d];
                    xt[e] = lo + (hi - lo) * u;
                    inb = inb && (lo < xt[e]) && (xt[e] < hi);
                }
            }
            inb = g.all(inb);
            double lp = Model::template log_prior<G, E>(g, xt, m);
            lp = inb ? lp : CPN_NEG_INF;
            if (need && lp > CPN_NEG_INF) {
#pragma unroll
                for (int e = 0; e < E; ++e) x[e] = xt[e];
                logP = lp;
                need = false;
            }
        }
    } else {
        bool inb = true;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            x[e] = (d < D) ? p.x[(size_t)c * p.Dp + d] : 0.0;
            if (d < D) inb = inb && (p.lo[d] < x[e]) && (x[e] < p.hi[d]);
        }
        inb = g.all(inb);
        const double lp = Model::template log_prior<G, E>(g, x, m);
        logP = inb ? lp : CPN_NEG_INF;
    }
    const double logL = Model::template log_likelihood<G, E>(g, x, m);
    if (p.grad_logL != nullptr) {
        double gl[E], gv[E];
        Model::template grad_log_likelihood<G, E>(g, x, m, gl);
        Model::template grad_potential<G, E>(g, x, m, gv);
        if (valid) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                if (d < D) { p.grad_logL[(size_t)c * p.Dp + d] = gl[e]; p.grad_V[(size_t)c * p.Dp + d] = gv[e]; }
            }
        }
    }
    if (valid) {
        if (p.draw) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                if (d < p.Dp) p.x[(size_t)c * p.Dp + d] = (d < D) ? x[e] : 0.0;
            }
        }
        if (g.lane == 0) {
            p.logL[c] = logL;
            p.logP[c] = logP;
        }
    }
}

}  // namespace cpn
