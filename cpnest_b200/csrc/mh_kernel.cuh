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

constexpr int kCycInline = 256;   // proposal cycles up to this length travel in the launch parameters (constant bank)
constexpr int kKindWin = 16;     // steps whose kinds fit one word of MHParams::cycw

struct MHParams : ChainParams {
    // proposal cycle + eigen directions
    const uint8_t* cycle;     // [cycle_len] in device memory (replay kernel; production when cycle_len > kCycInline)
    int cycle_len;
    int cycle_pos0;
    const double* eig_scale;  // [D]     sqrt|lambda_i|
    const double* eig_vec;    // [D][Dp] row i = eigenvector i
    int eig_row0;             // production: eig_vec == ens_x + eig_row0 * Dp (the eigen-directions live behind the live rows in one
                              // allocation, so a step gathers every row it needs from one base with a 32-bit row number)
    // replay (parity) mode
    const double* tape;       // [chains][maxmcmc][8]
    int compat;               // fp32 rounding of LivePoint scalars (cpnest/parameter.pyx:96-120)
    double* states;           // [chains][maxmcmc][D] or null
    // production: the cycle inline (ProposalCycle's default length is 100, cpnest/proposal.py:52): the kind of a step is the
    // same for every chain of the launch, read through the uniform datapath, and the proposal code branches on it uniformly
    uint8_t cyc[kCycInline];
    // cycw[q] = {the kinds of the kKindWin steps that start at cycle position q, two bits each (step j of the window in bits
    // 2j, 2j+1); the cycle position of the window after it, (q + kKindWin) % cycle_len}: one uniform load per window, a shift
    // and a mask per step, no arithmetic on the position
    uint2 cycw[kCycInline];
    // production, test entry (cpn_trace_mh): what every step of every chain consumed besides the chain state,
    // [chains][maxmcmc][8] doubles {kind, i0, i1, i2, c0, c1, c2, thr} (see DrawFeed), or null
    double* draws;
    // production, optional: the history of the first chain_log_n chains of the batch, the device form of the sampler's
    // `samples` deque (cpnest/sampler.py:345: every step appends the current point; written to mcmc_chain_*.dat at
    // verbose >= 3, :261-267).  A chain's state only changes when a step is accepted, so the log holds events, not steps:
    // rows [chain_log_n][chain_log_cap][Dp + 4] = {tag, step, logL, logP, x[Dp]}, tag 0: this is the state from `step` on
    // (step -1: the start point), tag 1: the chain ended after `step` steps; chain_log_count[slot] = events written so far
    // (the ring index is count % cap).  Null: off.
    double* chain_log;
    unsigned long long* chain_log_count;
    int chain_log_n, chain_log_cap;
};

// ---- what one MCMC step consumes besides the chain's current point ---------------------------------
// Nothing of it depends on the chain state, so the inputs (and the gather they imply) of step t+1 are
// produced while step t is still being evaluated.
//
// Production form, per proposal kind of the DefaultProposalCycle (r_j: gathered ensemble rows / eigen-direction):
//   EnsembleWalk          old + sum_j g_j (s_j - com), com = (s0+s1+s2)/3  =  old + sum_j c_j s_j, c_j = g_j - mean(g)   proposal.py:230-234
//   EnsembleStretch       a + (old - a) Z,  log_J = D x, Z = exp(x), x = uniform(-1,1) log 2                            proposal.py:253-260
//   DifferentialEvolution old + (b - a) g, g ~ N(1, 1e-4)                                                               proposal.py:284-286
//   EnsembleEigenVector   old + jump v_i, jump = sqrt|lambda_i| N(0,1)                                                   proposal.py:336-342
// The kind of step t is the same for all chains of a launch (same cycle, same position): it is read through the uniform
// datapath (MHParams::cycw), the step branches on it uniformly and gathers only the rows that kind needs (1.67 rows per
// step on average instead of 3).
// Scalars c_j, Z, g are fp32 values (the reference rounds the scalar of every LivePoint product to fp32,
// parameter.pyx:96-120); the EigenVector jump is fp64 (proposal.py:339-341 does not go through LivePoint.__mul__).
// `thr` = log(u) - log_J: the prior MH test (sampler.py:337) reads dlogP > thr.
template <bool B>
struct BoolC { static constexpr bool value = B; };

constexpr int kKindShift = 30;                         // ensemble sizes stay below 2^30
constexpr uint32_t kIndexMask = (1u << kKindShift) - 1u;

struct StepIn {
    uint32_t i0, i1, i2;      // rows to gather (EIGEN: i0 = eig_row0 + direction); chain-history instantiation: i0 carries the
                              // proposal kind in its two top bits (kKindShift)
    uint32_t w3, w4, w5;      // WALK: c0, c1, c2 (fp32 bits); STRETCH: w3 = Z; DE: w3 = g; EIGEN: (w4, w5) = jump (fp64 bits)
    double thr;
};

// The inputs of G consecutive steps are computed once, one step per lane (Philox blocks (chain, step, 0/1),
// Box-Muller, log(u)), packed into 8 words and handed out step after step through the lane's row of a
// shared-memory table (two 128-bit broadcast loads per step) instead of a redundant RNG evaluation on every lane;
// the stream is a pure function of (seed, chain, step) - independent of G and of the grid.
// FLAT: the prior is flat, so the prior MH test `dlogP > log(u) - log_J` (sampler.py:337) can only fail through the
// box (dlogP = -inf) unless log_J != 0, i.e. only EnsembleStretch needs log(u); the other kinds get thr = -1 (any
// negative number gives the same decisions as log(u) < 0).
template <int G, bool FLAT>
struct DrawFeed {
    uint4* row;      // this lane's 32-byte row of the block's table
    StepIn own;      // G == 1 only
    // dump: where to record the 8 values of this step (test entry), or null
    // tag: the kind rides in the top bits of i0 (kernels that cannot read it through the uniform datapath)
    // self_slot: the live slot the chain started from (-1: the start point was given).  The reference pops a chain's point
    // from the pool before it evolves it (sampler.py:325: `old = self.evolution_points.popleft()`), so a proposal never
    // meets the chain's own start among the ensemble members; here the members are drawn from the other ens_n - 1 slots
    // (a stretch move towards oneself would be a null move that counts as accepted).
    __device__ __forceinline__ void refill(int kind, uint32_t step, uint32_t cid_lo, uint32_t cid_hi, uint32_t k0, uint32_t k1,
                                           uint32_t ens_n, int self_slot, uint32_t D, const double* eig_scale, uint32_t eig_row0,
                                           bool tag, double* dump) {
        const uint32_t self = (uint32_t)self_slot;            // compact index c of the other slots -> slot c + (c >= self)
        if (self_slot >= 0) ens_n -= 1u;
        const Philox4 r = philox4x32_10(cid_lo, cid_hi, step, 0u, k0, k1);
        double thr = (FLAT && kind != CPN_STRETCH) ? -1.0 : log(u32_to_unit_open(r.v[3]));
        uint32_t i0 = u32_to_index(r.v[0], kind == CPN_EIGEN ? D : ens_n);       // EIGEN: randrange(D), proposal.py:338
        uint32_t i1 = i0, i2 = i0;
        uint32_t w3 = 0u, w4 = 0u, w5 = 0u;
        // inputs of the Box-Muller pair every kind but STRETCH needs (one shared evaluation: the lanes of a warp
        // prepare steps of different kinds, the code after the kind-specific part runs converged)
        uint32_t ba = r.v[1], bb = r.v[2], bc = 0u, bd = 0u;
        if (kind == CPN_WALK) {
            // three distinct members, `sample(list(ensemble), 3)` proposal.py:230
            i1 = u32_to_index(r.v[1], ens_n - 1);
            i1 += (i1 >= i0);
            i2 = u32_to_index(r.v[2], ens_n - 2);
            const uint32_t lo2 = min(i0, i1), hi2 = max(i0, i1);
            i2 += (i2 >= lo2);
            i2 += (i2 >= hi2);
            if (self_slot >= 0) { i0 += (i0 >= self); i1 += (i1 >= self); i2 += (i2 >= self); }
            const Philox4 q = philox4x32_10(cid_lo, cid_hi, step, 1u, k0, k1);
            ba = q.v[0]; bb = q.v[1]; bc = q.v[2]; bd = q.v[3];
        } else if (kind == CPN_DE) {
            i1 = u32_to_index(r.v[1], ens_n - 1);                                  // sample(...,2), proposal.py:284
            i1 += (i1 >= i0);
            if (self_slot >= 0) { i0 += (i0 >= self); i1 += (i1 >= self); }
            ba = r.v[2] << 16; bb = r.v[2] & 0xffff0000u;
        } else if (kind == CPN_STRETCH) {
            if (self_slot >= 0) i0 += (i0 >= self);
        }
        float ga, gb;
        box_muller_f(ba, bb, ga, gb);
        if (kind == CPN_WALK) {
            float gc, unused;
            box_muller_f(bc, bd, gc, unused);
            const float gbar = (ga + gb + gc) * (1.0f / 3.0f);
            w3 = __float_as_uint(ga - gbar); w4 = __float_as_uint(gb - gbar); w5 = __float_as_uint(gc - gbar);
        } else if (kind == CPN_STRETCH) {
            // x = uniform(-1,1) log 2, Z = exp(x) = 2^t rounded to fp32 (the LivePoint product rounds it, parameter.pyx:96),
            // log_J = D x from the un-rounded x (proposal.py:255-260)
            const float t = (float)(2.0 * u32_to_unit_open(r.v[1]) - 1.0);
            w3 = __float_as_uint(exp2f(t));
            thr -= (double)D * ((double)t * 0.69314718055994530942);
        } else if (kind == CPN_DE) {
            w3 = __float_as_uint(fmaf(1e-4f, ga, 1.0f));                            // gauss(1.0, 1e-4), proposal.py:285-286
        } else {
            const double jump = __ldg(eig_scale + i0) * (double)ga;               // sqrt|lambda_i| * gauss(0,1), :339
            w4 = (uint32_t)__double2loint(jump); w5 = (uint32_t)__double2hiint(jump);
        }
        if (dump != nullptr) {
            dump[0] = (double)kind; dump[1] = (double)i0; dump[2] = (double)i1; dump[3] = (double)i2;
            if (kind == CPN_EIGEN) {
                dump[4] = __hiloint2double((int)w5, (int)w4); dump[5] = 0.0; dump[6] = 0.0;
            } else {
                dump[4] = (double)__uint_as_float(w3); dump[5] = (double)__uint_as_float(w4); dump[6] = (double)__uint_as_float(w5);
            }
            dump[7] = thr;
        }
        if (kind == CPN_EIGEN) i0 += eig_row0;                                    // row of the direction in the joint array
        if (tag) i0 |= (uint32_t)kind << kKindShift;
        if (G == 1) {                       // a chain per lane: the step is this lane's own
            own.i0 = i0; own.i1 = i1; own.i2 = i2; own.w3 = w3; own.w4 = w4; own.w5 = w5; own.thr = thr;
            return;
        }
        __syncwarp();                       // every lane has read the previous G steps
        row[0] = make_uint4(i0, i1, i2, w3);
        row[1] = make_uint4(w4, w5, (uint32_t)__double2loint(thr), (uint32_t)__double2hiint(thr));
        __syncwarp();
    }
    // src_row: row of the lane of this group that prepared the step
    __device__ __forceinline__ StepIn get(const uint4* src_row) const {
        if (G == 1) return own;
        const uint4 a = src_row[0], b = src_row[1];
        StepIn s;
        s.i0 = a.x; s.i1 = a.y; s.i2 = a.z; s.w3 = a.w; s.w4 = b.x; s.w5 = b.y;
        s.thr = __hiloint2double((int)b.w, (int)b.z);
        return s;
    }
};

// The rows a step gathers: WALK three members, DE two, STRETCH one, EIGEN one eigen-direction, all from one array (live rows,
// then the eigen-directions) through one per-lane base pointer: address = base + row * (8 Dp), one IMAD.WIDE per row.  Every
// slot of a lane is loaded, also the slots beyond D of the last lanes, which then read the head of the following row (the
// array carries kRowSlack doubles of slack at its end).  Those values are finite and never used: the box test (+-inf bounds),
// the functors (valid_dim masks) and the final store ignore slots beyond D.
template <int G, int E>
__device__ __forceinline__ void gather_row(const double* lane_base, uint32_t row_bytes, uint32_t row, double (&r)[E]) {
    const double* b = reinterpret_cast<const double*>(reinterpret_cast<const char*>(lane_base) + (unsigned long long)row * row_bytes);
#pragma unroll
    for (int e = 0; e < E; ++e) r[e] = __ldg(b + e * G);
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

template <int G, int E, class Model, bool REPLAY, bool LOG = false>
// LOG: the production kernel with the chain history switched on (a separate instantiation: the event log costs registers
// and instruction-cache space in the hot loop, 8% of the step when merely compiled in)
// E <= 4: 128 registers, 4 blocks = 16 warps per SM, one wave for up to 4736 chains of 16 lanes (one latency-bound
// wave: time per launch is flat in the number of chains up to there)
__global__ void __launch_bounds__(kMHThreads, (E <= 4) ? 4 : 1) k_mh_chains(const MHParams p) {
    extern __shared__ double smem[];
    __shared__ uint4 feed_tab[REPLAY ? 1 : kMHThreads][2];     // DrawFeed: one 32-byte row per thread
    // team mode: the block's warps evolve one chain together and split the functor's data sum (models.cuh kTeam)
    constexpr int TEAM = (!REPLAY && G == 32 && Model::kTeam > 1) ? kMHThreads / 32 : 1;
    __shared__ double team_buf[TEAM];
    // block mode: the chains of the block evaluate the likelihood together (models.cuh block_coop): no warp leaves early,
    // the warps stay in lock step through the functor's block barriers
    constexpr bool BLK = !REPLAY && TEAM == 1 && Model::template block_coop<G>;
    Chain<G, E, Model> ch;
    if (!ch.begin(p, smem, REPLAY, TEAM, team_buf, BLK)) return;
    if constexpr (BLK) {
        Model::block_prepare(ch.m, Model::kScratch ? (kMHThreads / G) * kScratchRows * p.Dp : 0);
    }
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
        // are the constant all-ones, which lets the compiler drop the masking of every slot: those slots then hold exact
        // zeros in x, in every gathered row and hence in every proposal.
        auto production = [&](const ModelCtx& mm) {
        const uint32_t ens_n = (uint32_t)p.ens_n;
        DrawFeed<G, Model::kFlatPrior> feed;
        feed.row = &feed_tab[threadIdx.x][0];
        // row of this group's lane 0 in the feed table as a shared-window address held in a register (through a generic
        // pointer the compiler rebuilds the window base from SR_CgaCtaId in front of every read)
        uint32_t feed_base = (uint32_t)__cvta_generic_to_shared(&feed_tab[threadIdx.x & ~(G - 1)][0]);
        asm volatile("" : "+r"(feed_base));
        // KU: the kind of a step is read from the inline cycle through the uniform datapath (uniform branches, no
        // reconvergence bookkeeping); else (chain-history instantiation, which also takes the cycles too long to travel
        // inline) it rides in the top bits of the step's first index
        constexpr bool KU = !LOG;
        double* dump = (p.draws != nullptr && valid) ? p.draws + (size_t)c * maxmcmc * 8 : nullptr;
        // chain history (off unless the caller asked for it): events of this chain go to its slot's ring
        double* clog = nullptr;
        unsigned long long clog_n = 0;
        if (LOG && p.chain_log != nullptr && valid && ch.j < p.chain_log_n) {
            clog = p.chain_log + (size_t)ch.j * p.chain_log_cap * (p.Dp + 4);
            clog_n = p.chain_log_count[ch.j];
        }
        auto log_event = [&](double tag, double at) {          // executed by all lanes of a logging chain's group
            double* row = clog + (size_t)(clog_n % (unsigned long long)p.chain_log_cap) * (p.Dp + 4);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                if (d < p.Dp) row[4 + d] = (d < D) ? x[e] : 0.0;
            }
            if (g.lane == 0) { row[0] = tag; row[1] = at; row[2] = logL; row[3] = logP; }
            ++clog_n;
        };
        if (LOG && clog != nullptr) log_event(0.0, -1.0);
        const int max_resend = ch.can_restart(p) ? kMaxResend : 0;
        int attempts = 0, attempt_end = maxmcmc;
        // per-lane base of the row array (live rows, then the eigen-directions) in registers: opaque to the compiler, which
        // would otherwise re-read the pointer from the constant bank in front of every gather
        const double* lane_base = p.ens_x + g.lane;
        asm volatile("" : "+l"(lane_base));
        const uint32_t row_bytes = (uint32_t)p.Dp * 8u;
        const uint32_t eig_row0 = (uint32_t)p.eig_row0;
        // (the group's ballot mask too: rebuilt from the thread index in every step otherwise)
        Group<G> gp = g;
        asm volatile("" : "+r"(gp.gmask));
        // The inputs of step `it`.  They are drawn G steps at a time, one step per lane, when the first step of a window is
        // asked for; by then every lane holds the previous window in registers.
        auto load_in = [&](int it) -> StepIn {
            if ((it & (G - 1)) == 0) {                                   // this lane prepares step it + lane
                const int st = it + g.lane;
                const int lpos = (p.cycle_pos0 + 1 + st) % clen;           // pre-increment, proposal.py:93
                const int kind = KU ? (int)p.cyc[lpos] : (int)p.cycle[lpos];
                feed.refill(kind, (uint32_t)st, cid_lo, cid_hi, p.seed_lo, p.seed_hi, ens_n, ch.start_slot, (uint32_t)D, p.eig_scale, eig_row0, !KU,
                            (dump != nullptr && st < maxmcmc) ? dump + (size_t)st * 8 : nullptr);
            }
            if (G == 1) return feed.own;
            StepIn s;
            uint32_t t0, t1;
            const uint32_t addr = feed_base + 32u * (uint32_t)(it & (G - 1));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s.i0), "=r"(s.i1), "=r"(s.i2), "=r"(s.w3) : "r"(addr) : "memory");
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(s.w4), "=r"(s.w5), "=r"(t0), "=r"(t1) : "r"(addr) : "memory");
            s.thr = __hiloint2double((int)t1, (int)t0);
            return s;
        };
        auto gather = [&](int kind, const StepIn& s, double (&r0)[E], double (&r1)[E], double (&r2)[E]) {
            gather_row<G, E>(lane_base, row_bytes, KU ? s.i0 : (s.i0 & kIndexMask), r0);
            if ((kind & 1) == 0) {                                          // WALK, DE
                gather_row<G, E>(lane_base, row_bytes, s.i1, r1);
                if (kind == CPN_WALK) gather_row<G, E>(lane_base, row_bytes, s.i2, r2);
            }
        };
        // position in the cycle -> kind (KU: a uniform load from the launch parameters)
        // AHEAD: how many steps before its proposal the rows of a step are gathered.  2: two register sets alternate, a gather
        // has two steps to arrive (the L2 round trip is about as long as one step of a cheap functor).  1 (block mode: the
        // likelihood makes a step several L2 round trips long, and its own state needs the registers): one set.
        constexpr int AHEAD = BLK ? 1 : 2;
        // kinds of the window of kKindWin steps that holds the step whose inputs are read next (KU)
        // (the word of the following window is loaded one window ahead: its latency never meets a step)
        uint2 kw = p.cycw[KU ? npos : 0];
        uint2 kwn = p.cycw[KU ? kw.y : 0];
        auto kind_of = [&](int st) -> int {                                 // st in the current window, or the first of the next
            if ((st & (kKindWin - 1)) == 0 && st > 0) { kw = kwn; kwn = p.cycw[kw.y]; }
            return (int)((kw.x >> (2 * (st & (kKindWin - 1)))) & 3u);
        };
        // One MCMC step on its register set (s, kind, r0..r2).  As soon as the proposal has consumed the rows, the inputs
        // of the step AHEAD steps later are read and ITS rows are gathered into the same registers.
        // TAIL = false: no chain of the launch can stop at this step (fewer than min(Nmcmc, maxmcmc) steps done), so the
        // step carries no per-chain `active` state.
        auto mcmc_step = [&](auto tail, int it, StepIn& s, int& kind, double (&r0)[E], double (&r1)[E], double (&r2)[E]) {
            constexpr bool TAIL = decltype(tail)::value;
            double xn[E];
            // (two uniform bit tests: a four-way comparison chain becomes a jump table, an indirect branch per step)
            static_assert(CPN_WALK == 0 && CPN_STRETCH == 1 && CPN_DE == 2 && CPN_EIGEN == 3, "kind bits");
            if (kind & 2) {
                if (kind & 1) {                                             // EIGEN: old + jump v_i, proposal.py:340-341
                    const double jump = __hiloint2double((int)s.w5, (int)s.w4);
#pragma unroll
                    for (int e = 0; e < E; ++e) xn[e] = fma(jump, r0[e], x[e]);
                } else {                                                    // DE: old + (b - a) g, proposal.py:286
                    const double gd = (double)__uint_as_float(s.w3);
#pragma unroll
                    for (int e = 0; e < E; ++e) xn[e] = fma(gd, r1[e] - r0[e], x[e]);
                }
            } else {
                if (kind & 1) {                                             // STRETCH: a + (old - a) Z, proposal.py:258
                    const double Z = (double)__uint_as_float(s.w3);
#pragma unroll
                    for (int e = 0; e < E; ++e) xn[e] = fma(Z, x[e] - r0[e], r0[e]);
                } else {                                                    // WALK: old + sum_j c_j s_j, proposal.py:230-234
                    const double c0 = (double)__uint_as_float(s.w3), c1 = (double)__uint_as_float(s.w4), c2 = (double)__uint_as_float(s.w5);
#pragma unroll
                    for (int e = 0; e < E; ++e) xn[e] = fma(c2, r2[e], fma(c1, r1[e], fma(c0, r0[e], x[e])));
                }
            }
            const double thr = s.thr;
            s = load_in(it + AHEAD);
            kind = KU ? kind_of(it + AHEAD) : (int)(s.i0 >> kKindShift);
            gather(kind, s, r0, r1, r2);
            // strict box test, model.py:33 (independent comparison chains: the FP64 compare has a long latency)
            bool in4[4] = {true, true, true, true};
#pragma unroll
            for (int e = 0; e < E; ++e) {
                in4[(2 * e) & 3] = in4[(2 * e) & 3] & (blo[e] < xn[e]);
                in4[(2 * e + 1) & 3] = in4[(2 * e + 1) & 3] & (xn[e] < bhi[e]);
            }
            const bool inb = gp.all((in4[0] & in4[1]) & (in4[2] & in4[3]));
            double logP_new = Model::template log_prior<G, E>(g, xn, mm);
            bool pass;
            if (Model::kFlatPrior) {
                // the prior is a constant inside the box: `logP_new - logP > thr` (sampler.py:337) does not wait for the box test
                pass = (logP_new - logP > thr) && inb;
            } else {
                logP_new = inb ? logP_new : CPN_NEG_INF;
                pass = logP_new - logP > thr;                               // sampler.py:337
            }
            if (TAIL) pass = pass && active;
            if (BLK || Model::kCheap || __any_sync(FULL, pass)) {
                double logL_new;                                            // sampler.py:338
                if constexpr (BLK) logL_new = Model::template log_likelihood_block<G, E>(g, xn, mm);
                else logL_new = Model::template log_likelihood<G, E>(g, xn, mm);
                evals += pass ? 1 : 0;
                const bool acc = pass && logL_new > logLmin;                // sampler.py:339
                if (__any_sync(FULL, acc)) {
#pragma unroll
                    for (int e = 0; e < E; ++e) x[e] = acc ? xn[e] : x[e];
                    logL = acc ? logL_new : logL;
                    logP = acc ? logP_new : logP;
                    accepted += acc ? 1 : 0;
                    if (LOG && clog != nullptr && acc) log_event(0.0, (double)it);
                }
            }
            if (TAIL) {
                steps += active ? 1 : 0;
                if (steps >= nmcmc && accepted > 0) active = false;         // sampler.py:347
                else if (active && steps >= attempt_end) {                  // maxmcmc steps and no move
                    if (attempts < max_resend) {                            // NestedSampling.py:354-357: another sample, please
                        ++attempts;
                        attempt_end += maxmcmc;
                        ch.restart(p, attempts);
                        if (LOG && clog != nullptr) log_event(0.0, (double)it);
                    } else {
                        active = false;
                    }
                }
            }
        };

        StepIn sa, sb;
        double a0[E], a1[E], a2[E], b0[E], b1[E], b2[E];
#pragma unroll
        for (int e = 0; e < E; ++e) a0[e] = a1[e] = a2[e] = b0[e] = b1[e] = b2[e] = 0.0;
        sa = load_in(0);
        int ka = KU ? kind_of(0) : (int)(sa.i0 >> kKindShift);
        int kb = 0;
        gather(ka, sa, a0, a1, a2);
        if (AHEAD == 2) {
            sb = load_in(1);
            kb = KU ? kind_of(1) : (int)(sb.i0 >> kKindShift);
            gather(kb, sb, b0, b1, b2);
        }
        auto step_a = [&](auto tail, int it_) { mcmc_step(tail, it_, sa, ka, a0, a1, a2); };
        auto step_b = [&](auto tail, int it_) {
            if constexpr (AHEAD == 2) mcmc_step(tail, it_, sb, kb, b0, b1, b2);
            else mcmc_step(tail, it_, sa, ka, a0, a1, a2);
        };
        // a chain stops after step number n (counting from 1) if (n >= Nmcmc and it has moved) or n >= maxmcmc
        // (sampler.py:347): none does during the first min(Nmcmc, maxmcmc) - 1 steps
        const int n_main = max(0, min(nmcmc, maxmcmc) - 1) & ~1;
        // (the main phase in two halves: between them the chain may leave a snapshot for the two-lag control)
        const bool twolag = p.state != nullptr && !p.use_override && p.state->ctl[p.kind].twolag != 0;
        const int n_half = (n_main / 2) & ~1;
        int it = 0;
        for (int half = 0; half < 2; ++half) {
            const int it1 = half == 0 ? n_half : n_main;
            for (; it < it1; it += 2) {
                step_a(BoolC<false>{}, it);
                step_b(BoolC<false>{}, it + 1);
            }
            if (half == 0 && twolag) ch.snapshot_mid(p);
        }
        steps = n_main;
        const int it_end = maxmcmc * (1 + max_resend);
        for (; it < it_end; it += 2) {
            step_a(BoolC<true>{}, it);
            if (!(BLK ? __syncthreads_or(active) != 0 : __any_sync(FULL, active) != 0)) break;
            step_b(BoolC<true>{}, it + 1);
            if (!(BLK ? __syncthreads_or(active) != 0 : __any_sync(FULL, active) != 0)) break;
        }
        if (LOG && clog != nullptr) {
            log_event(1.0, (double)steps);
            if (g.lane == 0) p.chain_log_count[ch.j] = clog_n;
        }
        };
        if (p.padded && Model::pad_ok(m)) {
            ModelCtx mp = m;
#pragma unroll
            for (int e = 0; e < 8; ++e) mp.vmsk[e] = 0xffffffffu;
            Model::pad_specialise(mp);
            production(mp);
        } else {
            production(m);
        }
    }

    ch.finish(p, steps, accepted, evals, 0, 0);
}

}  // namespace cpn
