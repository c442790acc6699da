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
    unsigned long long* counters; // [0] steps [1] accepted [2] evals [3] failed chains
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

    int steps = 0, accepted = 0, evals = 0;
    bool active = true;
    int cpos = p.cycle_pos0;
    const double* tape = REPLAY ? p.tape + (size_t)c * maxmcmc * 8 : nullptr;
    const uint32_t ens_n = (uint32_t)p.ens_n;

    for (int it = 0; it < maxmcmc; ++it) {
        if (!__any_sync(FULL, active)) break;
        cpos = (cpos + 1 == p.cycle_len) ? 0 : cpos + 1;          // pre-increment, proposal.py:93
        const int kind = p.cycle[cpos];

        // ---- random inputs of this step --------------------------------------------------------
        uint32_t i0, i1, i2;
        double g0, g1, g2, us, umh;
        if (REPLAY) {
            const double* t = tape + (size_t)it * 8;
            i0 = (uint32_t)t[0]; i1 = (uint32_t)t[1]; i2 = (uint32_t)t[2];
            g0 = t[3]; g1 = t[4]; g2 = t[5]; us = t[6]; umh = t[7];
            if (p.compat && kind != CPN_EIGEN) {   // EnsembleEigenVector is full fp64 (proposal.py:339-341)
                g0 = (double)(float)g0; g1 = (double)(float)g1; g2 = (double)(float)g2;
            }
        } else {
            const Philox4 r = philox4x32_10(cid_lo, cid_hi, (uint32_t)it, 0u, p.seed_lo, p.seed_hi);
            umh = u32_to_unit_open(r.v[3]);
            i0 = i1 = i2 = 0;
            g0 = g1 = g2 = us = 0.0;
            if (kind == CPN_WALK) {
                // three distinct members, `sample(list(ensemble), 3)` proposal.py:230
                i0 = u32_to_index(r.v[0], ens_n);
                i1 = u32_to_index(r.v[1], ens_n - 1);
                i1 += (i1 >= i0);
                i2 = u32_to_index(r.v[2], ens_n - 2);
                const uint32_t lo2 = min(i0, i1), hi2 = max(i0, i1);
                i2 += (i2 >= lo2);
                i2 += (i2 >= hi2);
                const Philox4 q = philox4x32_10(cid_lo, cid_hi, (uint32_t)it, 1u, p.seed_lo, p.seed_hi);
                double g3;
                box_muller(q.v[0], q.v[1], g0, g1);
                box_muller(q.v[2], q.v[3], g2, g3);
            } else if (kind == CPN_STRETCH) {
                i0 = u32_to_index(r.v[0], ens_n);                    // random.choice, proposal.py:253
                us = 2.0 * u32_to_unit_open(r.v[1]) - 1.0;           // uniform(-1,1), proposal.py:255
            } else if (kind == CPN_DE) {
                i0 = u32_to_index(r.v[0], ens_n);                    // sample(...,2), proposal.py:284
                i1 = u32_to_index(r.v[1], ens_n - 1);
                i1 += (i1 >= i0);
                double gd, unused;
                box_muller(r.v[2] << 16, r.v[2] & 0xffff0000u, gd, unused);
                g0 = 1.0 + 1e-4 * gd;                                // gauss(1.0, sigma), proposal.py:285-286
            } else {
                i0 = u32_to_index(r.v[0], (uint32_t)D);              // randrange(D), proposal.py:338
                double unused;
                box_muller(r.v[1], r.v[2], g0, unused);
            }
        }

        // ---- proposal ------------------------------------------------------------------------------
        double xn[E];
        double log_J = 0.0;
        if (kind == CPN_WALK) {
            double s0[E], s1[E], s2[E];
            load_row<G, E>(g, p.ens_x + (size_t)i0 * p.Dp, D, s0);
            load_row<G, E>(g, p.ens_x + (size_t)i1 * p.Dp, D, s1);
            load_row<G, E>(g, p.ens_x + (size_t)i2 * p.Dp, D, s2);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (REPLAY) {
                    const double com = __ddiv_rn(__dadd_rn(__dadd_rn(s0[e], s1[e]), s2[e]), 3.0);
                    double o = x[e];
                    o = __dadd_rn(o, __dmul_rn(g0, __dsub_rn(s0[e], com)));
                    o = __dadd_rn(o, __dmul_rn(g1, __dsub_rn(s1[e], com)));
                    o = __dadd_rn(o, __dmul_rn(g2, __dsub_rn(s2[e], com)));
                    xn[e] = o;
                } else {
                    const double com = (s0[e] + s1[e] + s2[e]) * (1.0 / 3.0);
                    xn[e] = x[e] + g0 * (s0[e] - com) + g1 * (s1[e] - com) + g2 * (s2[e] - com);
                }
            }
        } else if (kind == CPN_STRETCH) {
            double a[E];
            load_row<G, E>(g, p.ens_x + (size_t)i0 * p.Dp, D, a);
            const double xs = A::mul(us, 0.69314718055994530942);     // uniform(-1,1)*log(scale)
            double Z = exp(xs);
            if (REPLAY && p.compat) Z = (double)(float)Z;
            log_J = A::mul((double)D, xs);                            // proposal.py:260 (un-rounded x)
#pragma unroll
            for (int e = 0; e < E; ++e) xn[e] = A::add(a[e], A::mul(Z, A::sub(x[e], a[e])));
        } else if (kind == CPN_DE) {
            double a[E], b[E];
            load_row<G, E>(g, p.ens_x + (size_t)i0 * p.Dp, D, a);
            load_row<G, E>(g, p.ens_x + (size_t)i1 * p.Dp, D, b);
#pragma unroll
            for (int e = 0; e < E; ++e) xn[e] = A::add(x[e], A::mul(g0, A::sub(b[e], a[e])));
        } else {
            double v[E];
            load_row<G, E>(g, p.eig_vec + (size_t)i0 * p.Dp, D, v);
            const double jump = A::mul(__ldg(p.eig_scale + i0), g0);   // sqrt|lambda_i| * gauss(0,1)
#pragma unroll
            for (int e = 0; e < E; ++e) xn[e] = A::add(x[e], A::mul(jump, v[e]));
        }

        // ---- box test, prior MH test, likelihood, hard constraint -------------------------------------
        bool inb = true;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            if (d < D) inb = inb && (__ldg(p.lo + d) < xn[e]) && (xn[e] < __ldg(p.hi + d));   // strict, model.py:33
        }
        inb = g.all(inb);
        double logP_new = Model::template log_prior<G, E>(g, xn, m);
        logP_new = inb ? logP_new : CPN_NEG_INF;
        const double delta = A::add(A::sub(logP_new, logP), log_J);   // sampler.py:337
        bool pass;
        if (REPLAY) {
            pass = delta > ((umh > 0.0) ? log(umh) : CPN_NEG_INF);
        } else {
            pass = (delta >= 0.0) || (delta > log(umh));                // log(u) < 0 always
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
            }
        }
        if (g.lane == 0 && p.counters != nullptr) {
            atomicAdd(p.counters + 0, (unsigned long long)steps);
            atomicAdd(p.counters + 1, (unsigned long long)accepted);
            atomicAdd(p.counters + 2, (unsigned long long)evals);
            if (accepted == 0) atomicAdd(p.counters + 3, 1ull);
        }
    }
}

}  // namespace cpn
