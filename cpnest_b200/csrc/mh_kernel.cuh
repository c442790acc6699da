// Constrained Metropolis-Hastings chains: the device form of
//   MetropolisHastingsSampler.yield_sample   cpnest/sampler.py:322-358
//   ProposalCycle.get_sample                 cpnest/proposal.py:91-97
//   EnsembleWalk / EnsembleStretch / DifferentialEvolution / EnsembleEigenVector
//                                            cpnest/proposal.py:209-342
//   Model.in_bounds / log_prior / log_likelihood   cpnest/model.py:20-33, 55-82
// One launch evolves all chains of a nested-sampling batch.  A chain is owned by a group of G
// lanes; every lane keeps E coordinates of the current and of the proposed point in registers.
// Proposal draws (Philox4x32-10), ensemble gathers from the live set in HBM/L2, the box test,
// the prior MH test, the likelihood functor and the hard constraint logL > logLmin all run in
// this kernel; nothing goes back to the host between steps.
#pragma once
#include <stdint.h>
#include "group.cuh"
#include "models.cuh"
#include "rng.cuh"
#include "state.cuh"
#include "cpnest_b200.h"

namespace cpn {

constexpr int kMHThreads = 128;
constexpr int kMaxPeers = 8;

enum StartMode { START_SURVIVOR = 0, START_SELF = 1, START_GIVEN = 2 };

struct MHParams {
    // ensemble = the live set, read-only during the launch
    const double* ens_x;      // [ens_n][Dp]
    const double* ens_logL;   // [ens_n]
    const double* ens_logP;   // [ens_n]
    int ens_n;
    int Dp;
    int D;
    // chain start
    int start_mode;
    const int* order;         // live slots sorted by ascending logL; survivors are order[n_skip..ens_n)
    int n_skip;
    const double* start_x;    // START_GIVEN: [chains][Dp]
    const double* start_logL;
    const double* start_logP;
    // box
    const double* lo;
    const double* hi;
    // proposal cycle + eigen directions
    const uint8_t* cycle;
    int cycle_len;
    int cycle_pos0;
    const double* eig_scale;  // [D]     sqrt|lambda_i|
    const double* eig_vec;    // [D][Dp] row i = eigenvector i
    // chain control
    const NSState* state;     // logLmin / nmcmc / done live on the device
    int use_override;
    double logLmin_override;
    int nmcmc_override;
    int maxmcmc;
    int j0, j1;               // chains [j0, j1) of the batch are evolved by this launch
    unsigned long long chain_base;
    uint32_t seed_lo, seed_hi;
    const double* blob;
    // outputs, replicated to every peer's staging area (n_out == 1 on a single GPU)
    int n_out;
    double* out_x[kMaxPeers];     // [K][Dp]
    double* out_logL[kMaxPeers];
    double* out_logP[kMaxPeers];
    int* out_steps[kMaxPeers];
    int* out_acc[kMaxPeers];
    int* out_start[kMaxPeers];    // live slot the chain started from (-1: given start)
    int* out_evals[kMaxPeers];    // likelihood evaluations of the chain
    // replay (parity) mode
    const double* tape;       // [chains][maxmcmc][8]
    int compat;               // fp32 rounding of LivePoint scalars (cpnest/parameter.pyx:96-120)
    double* states;           // [chains][maxmcmc][D] or null
};

template <bool EXACT>
struct Arith {
    // EXACT: one IEEE operation per reference operation (no FMA contraction), so a replay
    // reproduces the reference's fp64 numbers bit for bit.
    static __device__ __forceinline__ double add(double a, double b) { return EXACT ? __dadd_rn(a, b) : a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return EXACT ? __dsub_rn(a, b) : a - b; }
    static __device__ __forceinline__ double mul(double a, double b) { return EXACT ? __dmul_rn(a, b) : a * b; }
};

template <int G, int E>
__device__ __forceinline__ void load_row(const Group<G>& g, const double* row, int D, double (&v)[E]) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
        int d = g.lane + e * G;
        v[e] = (d < D) ? __ldg(row + d) : 0.0;
    }
}

// What one MCMC step consumes: the proposal kind, the ensemble members it draws and its random
// numbers.  Everything here is independent of the chain's current point, so the draw (and the
// gather it implies) of step t+1 is issued while step t is still being evaluated.
struct Draw {
    int kind;
    uint32_t i0, i1, i2;
    double g0, g1, g2, us, umh;
};

// Philox blocks of G consecutive steps are computed once, one step per lane, and handed out by
// shuffle: the stream is the same as if every lane computed block (chain, step, k) itself.
template <int G>
struct RngFeed {
    uint32_t a[4], b[4];
    __device__ __forceinline__ void refill(int step0, int lane, uint32_t cid_lo, uint32_t cid_hi, uint32_t k0, uint32_t k1) {
        const Philox4 p = philox4x32_10(cid_lo, cid_hi, (uint32_t)(step0 + lane), 0u, k0, k1);
        const Philox4 q = philox4x32_10(cid_lo, cid_hi, (uint32_t)(step0 + lane), 1u, k0, k1);
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = p.v[k]; b[k] = q.v[k]; }
    }
    __device__ __forceinline__ void get(int src_lane, uint32_t (&ra)[4], uint32_t (&rb)[4]) const {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ra[k] = (G == 1) ? a[k] : __shfl_sync(FULL, a[k], src_lane, G);
            rb[k] = (G == 1) ? b[k] : __shfl_sync(FULL, b[k], src_lane, G);
        }
    }
};

__device__ __forceinline__ Draw make_draw(int kind, const uint32_t (&r)[4], const uint32_t (&q)[4], uint32_t ens_n, uint32_t D) {
    Draw d;
    d.kind = kind;
    d.umh = u32_to_unit_open(r[3]);
    d.i0 = d.i1 = d.i2 = 0;
    d.g0 = d.g1 = d.g2 = d.us = 0.0;
    if (kind == CPN_WALK) {
        // three distinct members, `sample(list(ensemble), 3)` proposal.py:230
        d.i0 = u32_to_index(r[0], ens_n);
        d.i1 = u32_to_index(r[1], ens_n - 1);
        d.i1 += (d.i1 >= d.i0);
        d.i2 = u32_to_index(r[2], ens_n - 2);
        const uint32_t lo2 = min(d.i0, d.i1), hi2 = max(d.i0, d.i1);
        d.i2 += (d.i2 >= lo2);
        d.i2 += (d.i2 >= hi2);
        double g3;
        box_muller(q[0], q[1], d.g0, d.g1);
        box_muller(q[2], q[3], d.g2, g3);
    } else if (kind == CPN_STRETCH) {
        d.i0 = u32_to_index(r[0], ens_n);                    // random.choice, proposal.py:253
        d.us = 2.0 * u32_to_unit_open(r[1]) - 1.0;           // uniform(-1,1), proposal.py:255
    } else if (kind == CPN_DE) {
        d.i0 = u32_to_index(r[0], ens_n);                    // sample(...,2), proposal.py:284
        d.i1 = u32_to_index(r[1], ens_n - 1);
        d.i1 += (d.i1 >= d.i0);
        double gd, unused;
        box_muller(r[2] << 16, r[2] & 0xffff0000u, gd, unused);
        d.g0 = 1.0 + 1e-4 * gd;                              // gauss(1.0, sigma), proposal.py:285-286
    } else {
        d.i0 = u32_to_index(r[0], D);                        // randrange(D), proposal.py:338
        double unused;
        box_muller(r[1], r[2], d.g0, unused);
    }
    return d;
}

__device__ __forceinline__ Draw tape_draw(int kind, const double* t, int compat) {
    Draw d;
    d.kind = kind;
    d.i0 = (uint32_t)t[0]; d.i1 = (uint32_t)t[1]; d.i2 = (uint32_t)t[2];
    d.g0 = t[3]; d.g1 = t[4]; d.g2 = t[5]; d.us = t[6]; d.umh = t[7];
    if (compat && kind != CPN_EIGEN) {   // EnsembleEigenVector is full fp64 (proposal.py:339-341)
        d.g0 = (double)(float)d.g0; d.g1 = (double)(float)d.g1; d.g2 = (double)(float)d.g2;
    }
    return d;
}

// the (up to three) rows a draw gathers: ensemble members, or one eigenvector
template <int G, int E>
__device__ __forceinline__ void gather_rows(const Group<G>& g, const MHParams& p, const Draw& d, double (&r0)[E],
                                            double (&r1)[E], double (&r2)[E]) {
    const double* base0 = (d.kind == CPN_EIGEN) ? p.eig_vec : p.ens_x;
    load_row<G, E>(g, base0 + (size_t)d.i0 * p.Dp, p.D, r0);
    if (d.kind == CPN_WALK || d.kind == CPN_DE) load_row<G, E>(g, p.ens_x + (size_t)d.i1 * p.Dp, p.D, r1);
    if (d.kind == CPN_WALK) load_row<G, E>(g, p.ens_x + (size_t)d.i2 * p.Dp, p.D, r2);
}

template <int G, int E, class Model, bool REPLAY>
__global__ void __launch_bounds__(kMHThreads) k_mh_chains(const MHParams p) {
    using A = Arith<REPLAY>;
    if (!REPLAY && p.state != nullptr && p.state->done) return;
    const Group<G> g;
    const int cpb = kMHThreads / G;
    const int nchains = p.j1 - p.j0;
    int c = blockIdx.x * cpb + (int)threadIdx.x / G;      // chain of this launch
    if ((c & ~(32 / G - 1)) >= nchains) return;            // whole warp beyond the batch
    const bool valid = c < nchains;
    if (!valid) c = nchains - 1;                           // shadow a real chain, discard results
    const int j = p.j0 + c;                                // chain of the batch

    extern __shared__ double smem[];
    ModelCtx m;
    m.blob = p.blob;
    m.D = p.D;
    m.scratch = Model::kScratch ? smem + ((int)threadIdx.x / G) * p.Dp : nullptr;

    const int D = p.D;
    const double logLmin = p.use_override ? p.logLmin_override : p.state->logLmin;
    const int nmcmc = p.use_override ? p.nmcmc_override : p.state->nmcmc;
    const int maxmcmc = p.maxmcmc;
    const unsigned long long chain_id = p.chain_base + (unsigned long long)j;
    const uint32_t cid_lo = (uint32_t)chain_id, cid_hi = (uint32_t)(chain_id >> 32);

    // ---- start point -------------------------------------------------------------------------
    double x[E], logL, logP;
    int start_slot = -1;
    {
        const double* row;
        if (p.start_mode == START_GIVEN) {
            row = p.start_x + (size_t)c * p.Dp;
            logL = p.start_logL[c];
            logP = p.start_logP[c];
        } else {
            int slot;
            if (p.start_mode == START_SELF) {
                slot = j;
            } else {
                // a uniformly drawn survivor: every survivor satisfies logL > logLmin
                Philox4 r = philox4x32_10(cid_lo, cid_hi, 0xffffffffu, 0u, p.seed_lo, p.seed_hi);
                slot = p.order[p.n_skip + (int)u32_to_index(r.v[0], (uint32_t)(p.ens_n - p.n_skip))];
            }
            start_slot = slot;
            row = p.ens_x + (size_t)slot * p.Dp;
            logL = p.ens_logL[slot];
            logP = p.ens_logP[slot];
        }
        load_row<G, E>(g, row, D, x);
    }
    // the box, held in registers; coordinates beyond D get (-inf, +inf) so the test needs no branch
    double blo[E], bhi[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int d = g.lane + e * G;
        blo[e] = (d < D) ? __ldg(p.lo + d) : CPN_NEG_INF;
        bhi[e] = (d < D) ? __ldg(p.hi + d) : -CPN_NEG_INF;
    }

    int steps = 0, accepted = 0, evals = 0;
    bool active = true;
    int cpos = p.cycle_pos0;
    const double* tape = REPLAY ? p.tape + (size_t)c * maxmcmc * 8 : nullptr;
    const uint32_t ens_n = (uint32_t)p.ens_n;

    // ---- software pipeline: draw + gather of step `it+1` overlap the evaluation of step `it` -----------
    RngFeed<G> feed;
    Draw cur;
    double r0[E], r1[E], r2[E];
    auto next_draw = [&](int it) -> Draw {
        cpos = (cpos + 1 == p.cycle_len) ? 0 : cpos + 1;          // pre-increment, proposal.py:93
        const int kind = p.cycle[cpos];
        if (REPLAY) return tape_draw(kind, tape + (size_t)min(it, maxmcmc - 1) * 8, p.compat);
        if ((it & (G - 1)) == 0) feed.refill(it, g.lane, cid_lo, cid_hi, p.seed_lo, p.seed_hi);
        uint32_t ra[4], rb[4];
        feed.get(it & (G - 1), ra, rb);
        return make_draw(kind, ra, rb, ens_n, (uint32_t)D);
    };
    cur = next_draw(0);
#pragma unroll
    for (int e = 0; e < E; ++e) r1[e] = r2[e] = 0.0;
    gather_rows<G, E>(g, p, cur, r0, r1, r2);

    for (int it = 0; it < maxmcmc; ++it) {
        if (!__any_sync(FULL, active)) break;
        const Draw nxt = next_draw(it + 1);
        double n0[E], n1[E], n2[E];
#pragma unroll
        for (int e = 0; e < E; ++e) n1[e] = n2[e] = 0.0;
        gather_rows<G, E>(g, p, nxt, n0, n1, n2);

        // ---- proposal ------------------------------------------------------------------------------
        const int kind = cur.kind;
        double xn[E];
        double log_J = 0.0;
        if (kind == CPN_WALK) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (REPLAY) {
                    const double com = __ddiv_rn(__dadd_rn(__dadd_rn(r0[e], r1[e]), r2[e]), 3.0);
                    double o = x[e];
                    o = __dadd_rn(o, __dmul_rn(cur.g0, __dsub_rn(r0[e], com)));
                    o = __dadd_rn(o, __dmul_rn(cur.g1, __dsub_rn(r1[e], com)));
                    o = __dadd_rn(o, __dmul_rn(cur.g2, __dsub_rn(r2[e], com)));
                    xn[e] = o;
                } else {
                    const double com = (r0[e] + r1[e] + r2[e]) * (1.0 / 3.0);
                    xn[e] = x[e] + cur.g0 * (r0[e] - com) + cur.g1 * (r1[e] - com) + cur.g2 * (r2[e] - com);
                }
            }
        } else if (kind == CPN_STRETCH) {
            const double xs = A::mul(cur.us, 0.69314718055994530942);     // uniform(-1,1)*log(scale)
            double Z = exp(xs);
            if (REPLAY && p.compat) Z = (double)(float)Z;
            log_J = A::mul((double)D, xs);                                // proposal.py:260 (un-rounded x)
#pragma unroll
            for (int e = 0; e < E; ++e) xn[e] = A::add(r0[e], A::mul(Z, A::sub(x[e], r0[e])));
        } else if (kind == CPN_DE) {
#pragma unroll
            for (int e = 0; e < E; ++e) xn[e] = A::add(x[e], A::mul(cur.g0, A::sub(r1[e], r0[e])));
        } else {
            const double jump = A::mul(__ldg(p.eig_scale + cur.i0), cur.g0);   // sqrt|lambda_i| * gauss(0,1)
#pragma unroll
            for (int e = 0; e < E; ++e) xn[e] = A::add(x[e], A::mul(jump, r0[e]));
        }

        // ---- box test, prior MH test, likelihood, hard constraint -------------------------------------
        bool inb = true;
#pragma unroll
        for (int e = 0; e < E; ++e) inb = inb & (blo[e] < xn[e]) & (xn[e] < bhi[e]);    // strict, model.py:33
        inb = g.all(inb);
        double logP_new = Model::template log_prior<G, E>(g, xn, m);
        logP_new = inb ? logP_new : CPN_NEG_INF;
        const double delta = A::add(A::sub(logP_new, logP), log_J);   // sampler.py:337
        bool pass;
        if (REPLAY) {
            pass = delta > ((cur.umh > 0.0) ? log(cur.umh) : CPN_NEG_INF);
        } else {
            pass = (delta >= 0.0) || (delta > log(cur.umh));            // log(u) < 0 always
        }
        pass = pass && active;
        if (__any_sync(FULL, pass)) {
            const double logL_new = Model::template log_likelihood<G, E>(g, xn, m);   // sampler.py:338
            evals += pass ? 1 : 0;
            if (pass && logL_new > logLmin) {                           // sampler.py:339
#pragma unroll
                for (int e = 0; e < E; ++e) x[e] = xn[e];
                logL = logL_new;
                logP = logP_new;
                ++accepted;
            }
        }
        if (active) {
            ++steps;
            if (REPLAY && p.states != nullptr && valid) {
                double* srow = p.states + ((size_t)c * maxmcmc + it) * D;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int d = g.lane + e * G;
                    if (d < D) srow[d] = x[e];
                }
            }
            if ((steps >= nmcmc && accepted > 0) || steps >= maxmcmc) active = false;   // sampler.py:347
        }
        cur = nxt;
#pragma unroll
        for (int e = 0; e < E; ++e) { r0[e] = n0[e]; r1[e] = n1[e]; r2[e] = n2[e]; }
    }

    // ---- hand the evolved point back (`producer_pipe.send`, sampler.py:246) ------------------
    if (valid) {
        for (int o = 0; o < p.n_out; ++o) {
            double* row = p.out_x[o] + (size_t)j * p.Dp;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                if (d < p.Dp) row[d] = (d < D) ? x[e] : 0.0;
            }
            if (g.lane == 0) {
                p.out_logL[o][j] = logL;
                p.out_logP[o][j] = logP;
                p.out_steps[o][j] = steps;
                p.out_acc[o][j] = accepted;
                if (p.out_start[o] != nullptr) p.out_start[o][j] = start_slot;
                if (p.out_evals[o] != nullptr) p.out_evals[o][j] = evals;
            }
        }
    }
}

}  // namespace cpn
