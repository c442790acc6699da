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
#include "chain_common.cuh"

namespace cpn {

struct MHParams : ChainParams {
    // proposal cycle + eigen directions
    const uint8_t* cycle;
    int cycle_len;
    int cycle_pos0;
    const double* eig_scale;  // [D]     sqrt|lambda_i|
    const double* eig_vec;    // [D][Dp] row i = eigenvector i
    // replay (parity) mode
    const double* tape;       // [chains][maxmcmc][8]
    int compat;               // fp32 rounding of LivePoint scalars (cpnest/parameter.pyx:96-120)
    double* states;           // [chains][maxmcmc][D] or null
};

// ---- what one MCMC step consumes besides the chain's current point ---------------------------------
// Nothing of it depends on the chain state, so the inputs (and the gather they imply) of step t+1 are
// produced while step t is still being evaluated.
//
// Production form.  All four proposals of the DefaultProposalCycle are instances of
//     new = cx * old + c0 * r0 + c1 * r1 + c2 * r2          (r_j: gathered ensemble rows / eigen-direction)
//   EnsembleWalk          old + sum_j g_j (s_j - com), com = (s0+s1+s2)/3  ->  cx = 1, c_j = g_j - mean(g)   proposal.py:230-234
//   EnsembleStretch       a + Z (old - a)                                  ->  cx = Z, c0 = 1 - Z            proposal.py:253-260
//   DifferentialEvolution old + g (b - a), g ~ N(1, 1e-4)                  ->  cx = 1, c0 = -g, c1 = g       proposal.py:284-286
//   EnsembleEigenVector   old + sqrt|lambda_i| N(0,1) v_i                  ->  cx = 1, c0 = jump, r0 = v_i   proposal.py:336-342
// so a step is straight-line code whatever the proposal kind: the compiler can interleave the draw and the
// gather of the following step with the dependent chain proposal -> box test -> likelihood reduction ->
// accept of this one.  `thr` = log(u) - log_J: the prior MH test (sampler.py:337) reads dlogP > thr.
struct StepIn {
    int kind;
    uint32_t i0, i1, i2;
    double cx, c0, c1, c2;
    double thr;
};

// The inputs of G consecutive steps are computed once, one step per lane (Philox blocks (chain, step, 0/1),
// Box-Muller, log(u), exp), packed into 8 words and handed out step after step through the lane's row of a
// shared-memory table (two 128-bit broadcast loads per step; 8 shuffles before v8) instead of a redundant RNG
// evaluation on every lane; the stream is a pure function of (seed, chain, step) - independent of G and of the grid.
// FLAT: the prior is flat, so the prior MH test `dlogP > log(u) - log_J` (sampler.py:337) can only fail through the
// box (dlogP = -inf) unless log_J != 0, i.e. only EnsembleStretch needs log(u); the other kinds get thr = -1 (any
// negative number gives the same decisions as log(u) < 0).
template <int G, bool FLAT>
struct DrawFeed {
    uint32_t w[8];   // i0, i1, i2, c0, c1, c2 (fp32 bits), thr (fp64 bits)
    uint4* row;      // this lane's 32-byte row of the block's table (G > 1)
    __device__ __forceinline__ void refill(int kind, uint32_t step, uint32_t cid_lo, uint32_t cid_hi, uint32_t k0, uint32_t k1,
                                           uint32_t ens_n, uint32_t D, const double* eig_scale) {
        const Philox4 r = philox4x32_10(cid_lo, cid_hi, step, 0u, k0, k1);
        double thr = (FLAT && kind != CPN_STRETCH) ? -1.0 : log(u32_to_unit_open(r.v[3]));
        uint32_t i0 = u32_to_index(r.v[0], ens_n), i1, i2;
        float c0, c1 = 0.0f, c2 = 0.0f;
        if (kind == CPN_WALK) {
            // three distinct members, `sample(list(ensemble), 3)` proposal.py:230
            i1 = u32_to_index(r.v[1], ens_n - 1);
            i1 += (i1 >= i0);
            i2 = u32_to_index(r.v[2], ens_n - 2);
            const uint32_t lo2 = min(i0, i1), hi2 = max(i0, i1);
            i2 += (i2 >= lo2);
            i2 += (i2 >= hi2);
            const Philox4 q = philox4x32_10(cid_lo, cid_hi, step, 1u, k0, k1);
            float g0, g1, g2, g3;
            box_muller_f(q.v[0], q.v[1], g0, g1);
            box_muller_f(q.v[2], q.v[3], g2, g3);
            const float gbar = (g0 + g1 + g2) * (1.0f / 3.0f);
            c0 = g0 - gbar; c1 = g1 - gbar; c2 = g2 - gbar;
        } else if (kind == CPN_STRETCH) {
            i1 = i2 = i0;                                                          // random.choice, proposal.py:253
            const double xs = (2.0 * u32_to_unit_open(r.v[1]) - 1.0) * 0.69314718055994530942;   // uniform(-1,1)*log(2)
            c0 = (float)(1.0 - exp(xs));                                           // Z = exp(x), proposal.py:256-258
            thr -= (double)D * log(1.0 - (double)c0);                              // log_J = D log Z of the Z applied, :260
        } else if (kind == CPN_DE) {
            i1 = u32_to_index(r.v[1], ens_n - 1);                                  // sample(...,2), proposal.py:284
            i1 += (i1 >= i0);
            i2 = i1;
            float gd, unused;
            box_muller_f(r.v[2] << 16, r.v[2] & 0xffff0000u, gd, unused);
            c1 = fmaf(1e-4f, gd, 1.0f);                                            // gauss(1.0, 1e-4), proposal.py:285-286
            c0 = -c1;
        } else {
            i0 = u32_to_index(r.v[0], D);                                          // randrange(D), proposal.py:338
            i1 = i2 = i0;
            float g, unused;
            box_muller_f(r.v[1], r.v[2], g, unused);
            c0 = (float)(__ldg(eig_scale + i0) * (double)g);                       // sqrt|lambda_i| * gauss(0,1), :339
        }
        w[0] = i0; w[1] = i1; w[2] = i2;
        w[3] = __float_as_uint(c0); w[4] = __float_as_uint(c1); w[5] = __float_as_uint(c2);
        w[6] = (uint32_t)__double2loint(thr);
        w[7] = (uint32_t)__double2hiint(thr);
        if (G > 1) {
            __syncwarp();                       // every lane has read the previous G steps
            row[0] = make_uint4(w[0], w[1], w[2], w[3]);
            row[1] = make_uint4(w[4], w[5], w[6], w[7]);
            __syncwarp();
        }
    }
    // src_row: row of the lane of this group that prepared the step
    __device__ __forceinline__ StepIn get(int kind, const uint4* src_row) const {
        uint32_t v[8];
        if (G == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = w[k];
        } else {
            const uint4 a = src_row[0], b = src_row[1];
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        StepIn s;
        s.kind = kind;
        s.i0 = v[0]; s.i1 = v[1]; s.i2 = v[2];
        s.c0 = (double)__uint_as_float(v[3]); s.c1 = (double)__uint_as_float(v[4]); s.c2 = (double)__uint_as_float(v[5]);
        s.cx = (kind == CPN_STRETCH) ? 1.0 - s.c0 : 1.0;
        s.thr = __hiloint2double((int)v[7], (int)v[6]);
        return s;
    }
};

// the three rows a production step gathers (kinds that need fewer re-read row i0: an L1 hit).  Straight-line:
// every slot is loaded, also the slots beyond D of the last lanes, which then read the head of the following row
// (the arrays carry kRowSlack doubles of slack at their end).  Those values are finite and never used: the box
// test (+-inf bounds), the functors (valid_dim masks) and the final store ignore slots beyond D.
template <int G, int E>
__device__ __forceinline__ void gather_rows(const Group<G>& g, const MHParams& p, const StepIn& s, double (&r0)[E],
                                            double (&r1)[E], double (&r2)[E]) {
    const bool eig = s.kind == CPN_EIGEN;
    const unsigned Dp = (unsigned)p.Dp, lane = (unsigned)g.lane;
    const double* b0 = (eig ? p.eig_vec : p.ens_x) + (s.i0 * Dp + lane);
    const double* b1 = p.ens_x + (s.i1 * Dp + lane);
    const double* b2 = p.ens_x + (s.i2 * Dp + lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        r0[e] = __ldg(b0 + e * G);
        r1[e] = __ldg(b1 + e * G);
        r2[e] = __ldg(b2 + e * G);
    }
}

// ---- replay (parity) form: the reference's own random numbers from the tape, one IEEE operation per
// reference operation ---------------------------------------------------------------------------------
struct TapeIn {
    int kind;
    uint32_t i0, i1, i2;
    double a0, a1, a2;     // WALK: the 3 Gaussians; STRETCH: a0 = Z; DE: a0 = factor; EIGEN: a0 = N(0,1)
    double log_J;          // proposal.py:260 (STRETCH), else 0
    double logu;           // log of the uniform of the prior MH test, sampler.py:337
};

__device__ __forceinline__ TapeIn tape_step(int kind, const double* t, int compat, int D) {
    TapeIn s;
    s.kind = kind;
    s.i0 = (uint32_t)t[0]; s.i1 = (uint32_t)t[1]; s.i2 = (uint32_t)t[2];
    s.a0 = t[3]; s.a1 = t[4]; s.a2 = t[5];
    s.log_J = 0.0;
    if (compat && kind != CPN_EIGEN) {   // EnsembleEigenVector is full fp64 (proposal.py:339-341)
        s.a0 = (double)(float)s.a0; s.a1 = (double)(float)s.a1; s.a2 = (double)(float)s.a2;
    }
    if (kind == CPN_STRETCH) {
        const double xs = __dmul_rn(t[6], 0.69314718055994530942);     // uniform(-1,1)*log(scale), proposal.py:255-256
        double Z = exp(xs);
        if (compat) Z = (double)(float)Z;
        s.a0 = Z;
        s.log_J = __dmul_rn((double)D, xs);                              // proposal.py:260 (un-rounded x)
    }
    s.logu = (t[7] > 0.0) ? log(t[7]) : CPN_NEG_INF;
    return s;
}

template <int G, int E, class Model, bool REPLAY>
// E <= 4: 128 registers, 4 blocks = 16 warps per SM, one wave for up to 4736 chains of 16 lanes (one latency-bound
// wave: time per launch is flat in the number of chains up to there)
__global__ void __launch_bounds__(kMHThreads, (E <= 4) ? 4 : 1) k_mh_chains(const MHParams p) {
    extern __shared__ double smem[];
    __shared__ uint4 feed_tab[REPLAY ? 1 : kMHThreads][2];     // DrawFeed: one 32-byte row per thread
    // team mode: the block's warps evolve one chain together and split the functor's data sum (models.cuh kTeam)
    constexpr int TEAM = (!REPLAY && G == 32 && Model::kTeam > 1) ? kMHThreads / 32 : 1;
    __shared__ double team_buf[TEAM];
    Chain<G, E, Model> ch;
    if (!ch.begin(p, smem, REPLAY, TEAM, team_buf)) return;
    const Group<G>& g = ch.g;
    const ModelCtx& m = ch.m;
    const int c = ch.c, D = p.D, nmcmc = ch.nmcmc, maxmcmc = p.maxmcmc;
    const bool valid = ch.valid;
    const double logLmin = ch.logLmin;
    const uint32_t cid_lo = ch.cid_lo, cid_hi = ch.cid_hi;
    double (&x)[E] = ch.x;
    double (&blo)[E] = ch.blo;
    double (&bhi)[E] = ch.bhi;
    double& logL = ch.logL;
    double& logP = ch.logP;

    int steps = 0, accepted = 0, evals = 0;
    bool active = true;
    const int clen = p.cycle_len;
    // position in the proposal cycle of the next step to be drawn (pre-increment, proposal.py:93)
    int npos = (p.cycle_pos0 + 1) % clen;

    if constexpr (REPLAY) {
        // ================================ parity mode =================================================
        const double* tape = p.tape + (size_t)c * maxmcmc * 8;
        for (int it = 0; it < maxmcmc; ++it) {
            if (!__any_sync(FULL, active)) break;
            const int kind = p.cycle[npos];
            npos = (npos + 1 == clen) ? 0 : npos + 1;
            const TapeIn s = tape_step(kind, tape + (size_t)it * 8, p.compat, D);
            double r0[E], r1[E], r2[E], xn[E];
#pragma unroll
            for (int e = 0; e < E; ++e) r1[e] = r2[e] = 0.0;
            load_row<G, E>(g, (kind == CPN_EIGEN ? p.eig_vec : p.ens_x) + (size_t)s.i0 * p.Dp, D, r0);
            if (kind == CPN_WALK || kind == CPN_DE) load_row<G, E>(g, p.ens_x + (size_t)s.i1 * p.Dp, D, r1);
            if (kind == CPN_WALK) load_row<G, E>(g, p.ens_x + (size_t)s.i2 * p.Dp, D, r2);
            if (kind == CPN_WALK) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double com = __ddiv_rn(__dadd_rn(__dadd_rn(r0[e], r1[e]), r2[e]), 3.0);
                    double o = x[e];
                    o = __dadd_rn(o, __dmul_rn(s.a0, __dsub_rn(r0[e], com)));
                    o = __dadd_rn(o, __dmul_rn(s.a1, __dsub_rn(r1[e], com)));
                    o = __dadd_rn(o, __dmul_rn(s.a2, __dsub_rn(r2[e], com)));
                    xn[e] = o;
                }
            } else if (kind == CPN_STRETCH) {
#pragma unroll
                for (int e = 0; e < E; ++e) xn[e] = __dadd_rn(r0[e], __dmul_rn(s.a0, __dsub_rn(x[e], r0[e])));
            } else if (kind == CPN_DE) {
#pragma unroll
                for (int e = 0; e < E; ++e) xn[e] = __dadd_rn(x[e], __dmul_rn(s.a0, __dsub_rn(r1[e], r0[e])));
            } else {
                const double jump = __dmul_rn(__ldg(p.eig_scale + s.i0), s.a0);   // sqrt|lambda_i| * gauss(0,1)
#pragma unroll
                for (int e = 0; e < E; ++e) xn[e] = __dadd_rn(x[e], __dmul_rn(jump, r0[e]));
            }
            bool inb = true;
#pragma unroll
            for (int e = 0; e < E; ++e) inb = inb & (blo[e] < xn[e]) & (xn[e] < bhi[e]);    // strict, model.py:33
            inb = g.all(inb);
            double logP_new = Model::template log_prior<G, E>(g, xn, m);
            logP_new = inb ? logP_new : CPN_NEG_INF;
            const double delta = __dadd_rn(__dsub_rn(logP_new, logP), s.log_J);   // sampler.py:337
            const bool pass = (delta > s.logu) && active;
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
                if (p.states != nullptr && valid) {
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
    } else {
        // ================================ production ==================================================
        // The loop exists twice: `mm` is the chain's functor context, or - when the gathered rows are zero-padded up to
        // G*E doubles and the functor tolerates zeros in the slots beyond D (Model::pad_ok) - a copy whose slot masks
        // are the constant all-ones, which lets the compiler drop the masking of every slot (17 of 155 instructions per
        // step for the iid Gaussian at D=50): those slots then hold exact zeros in x, in every gathered row and hence in
        // every proposal.
        auto production = [&](const ModelCtx& mm) {
        const uint32_t ens_n = (uint32_t)p.ens_n;
        DrawFeed<G, Model::kFlatPrior> feed;
        feed.row = &feed_tab[threadIdx.x][0];
        const uint4* grow = &feed_tab[threadIdx.x & ~(G - 1)][0];        // row of this group's lane 0
        auto next_in = [&](int it) -> StepIn {
            const int kind = p.cycle[npos];
            if ((it & (G - 1)) == 0) {                                   // this lane prepares step it + lane
                int lpos = npos + g.lane;                                // (npos + lane) % clen without the division
                while (lpos >= clen) lpos -= clen;
                feed.refill(p.cycle[lpos], (uint32_t)(it + g.lane), cid_lo, cid_hi, p.seed_lo, p.seed_hi, ens_n, (uint32_t)D,
                            p.eig_scale);
            }
            npos = (npos + 1 == clen) ? 0 : npos + 1;
            return feed.get(kind, grow + 2 * (it & (G - 1)));
        };
        // One MCMC step on (s, r0..r2); meanwhile the inputs of the following step are drawn and gathered into
        // (ns, n0..n2).  The caller alternates two such sets, so nothing is copied between steps.
        auto mcmc_step = [&](int it, const StepIn& s, const double (&r0)[E], const double (&r1)[E], const double (&r2)[E],
                             StepIn& ns, double (&n0)[E], double (&n1)[E], double (&n2)[E]) {
            ns = next_in(it + 1);
            gather_rows<G, E>(g, p, ns, n0, n1, n2);
            double xn[E];
            bool inb = true;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                xn[e] = fma(s.c2, r2[e], fma(s.c1, r1[e], fma(s.c0, r0[e], s.cx * x[e])));
                inb = inb & (blo[e] < xn[e]) & (xn[e] < bhi[e]);            // strict, model.py:33
            }
            inb = g.all(inb);
            double logP_new = Model::template log_prior<G, E>(g, xn, mm);
            logP_new = inb ? logP_new : CPN_NEG_INF;
            const bool pass = (logP_new - logP > s.thr) && active;          // sampler.py:337
            if (Model::kCheap || __any_sync(FULL, pass)) {
                const double logL_new = Model::template log_likelihood<G, E>(g, xn, mm);   // sampler.py:338
                evals += pass ? 1 : 0;
                const bool acc = pass && logL_new > logLmin;                // sampler.py:339
#pragma unroll
                for (int e = 0; e < E; ++e) x[e] = acc ? xn[e] : x[e];
                logL = acc ? logL_new : logL;
                logP = acc ? logP_new : logP;
                accepted += acc ? 1 : 0;
            }
            steps += active ? 1 : 0;
            if ((steps >= nmcmc && accepted > 0) || steps >= maxmcmc) active = false;   // sampler.py:347
        };

        StepIn sa, sb;
        double a0[E], a1[E], a2[E], b0[E], b1[E], b2[E];
#pragma unroll
        for (int e = 0; e < E; ++e) a0[e] = a1[e] = a2[e] = b0[e] = b1[e] = b2[e] = 0.0;
        sa = next_in(0);
        gather_rows<G, E>(g, p, sa, a0, a1, a2);
        for (int it = 0; it < maxmcmc; it += 2) {
            mcmc_step(it, sa, a0, a1, a2, sb, b0, b1, b2);
            if (!__any_sync(FULL, active)) break;
            mcmc_step(it + 1, sb, b0, b1, b2, sa, a0, a1, a2);
            if (!__any_sync(FULL, active)) break;
        }
        };
        if (p.padded && Model::pad_ok(m)) {
            ModelCtx mp = m;
#pragma unroll
            for (int e = 0; e < 8; ++e) mp.vmsk[e] = 0xffffffffu;
            production(mp);
        } else {
            production(m);
        }
    }

    ch.finish(p, steps, accepted, evals, 0, 0);
}

}  // namespace cpn
