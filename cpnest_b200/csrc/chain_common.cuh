// What the three chain kernels (mh_kernel.cuh, slice_kernel.cuh, hmc_kernel.cuh) share: the launch parameters that
// describe the ensemble, the chain start, the box, the chain control and the outputs; the per-thread chain context
// (which chain, its start point, the box in registers, the functor context); and the store of the evolved point
// (`producer_pipe.send((acceptance, sub_acceptance, Nmcmc, point))`, cpnest/sampler.py:246).
#pragma once
#include "sysinc.cuh"
#include "group.cuh"
#include "models.cuh"
#include "rng.cuh"
#include "state.cuh"
#include "cpnest_b200.h"

namespace cpn {

constexpr int kMHThreads = 128;   // threads per block of every chain kernel
constexpr int kMaxPeers = 8;
constexpr int kMaxResend = 15;     // attempts of a chain that cannot move before it gives up (NestedSampling.py:354-357 loops for ever)
constexpr int kRowSlack = 256;    // doubles of slack behind row arrays the chain kernels gather from (mh_kernel.cuh gather_rows)

enum StartMode { START_SURVIVOR = 0, START_SELF = 1, START_GIVEN = 2 };

struct ChainParams {
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
    // chain control
    const NSState* state;     // logLmin / nmcmc / done live on the device
    int kind;                 // cpn_sampler whose controller (state->ctl[kind]) gives the chain length
    int use_override;
    double logLmin_override;
    int nmcmc_override;
    int maxmcmc;
    int j0, j1;               // chains [j0, j1) of the batch are evolved by this launch
    int padded;               // rows of ens_x / eig_vec are zero-padded to at least G*E doubles (Dp >= G*E)
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
    int* out_ne[kMaxPeers];       // slice sampler: expansions summed over the chain's slices (else null)
    int* out_nc[kMaxPeers];       // slice sampler: contractions
    double* out_mid_x[kMaxPeers]; // MH, two-lag control: the chain's point after half of its steps (KindCtl::twolag), [K][Dp]
    double* out_mid_logL[kMaxPeers];
};

template <int G, int E>
__device__ __forceinline__ void load_row(const Group<G>& g, const double* row, int D, double (&v)[E]) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
        int d = g.lane + e * G;
        v[e] = (d < D) ? __ldg(row + d) : 0.0;
    }
}

// Per-thread context of a chain: G lanes own it, each holds E coordinates (coordinate d <-> lane d % G, slot d / G).
template <int G, int E, class Model>
struct Chain {
    Group<G> g;
    int c;                 // chain of this launch
    int j;                 // chain of the batch
    bool valid;            // false: this group shadows the last real chain and discards its results
    ModelCtx m;
    uint32_t cid_lo, cid_hi;   // Philox counter words identifying the chain
    double logLmin;
    int nmcmc;
    double x[E], logL, logP;
    int start_slot;
    double blo[E], bhi[E];     // the box; coordinates beyond D get (-inf, +inf) so the test needs no branch

    // false: the whole warp lies beyond the batch (or the run is over) and must return.
    // team > 1 (G == 32 only): the block's `team` warps all own chain blockIdx.x; team_buf is block-shared scratch
    // keep_all: warps beyond the batch stay (they shadow the last chain): the kernel runs block barriers
    __device__ __forceinline__ bool begin(const ChainParams& p, double* smem, bool replay, int team = 1, double* team_buf = nullptr,
                                          bool keep_all = false) {
        if (!replay && p.state != nullptr && p.state->done) return false;
        const int cpb = kMHThreads / G;
        const int nchains = p.j1 - p.j0;
        if (team > 1) {
            c = blockIdx.x;
            if (c >= nchains) return false;
            m.team = team;
            m.team_rank = (int)threadIdx.x / G;
            m.team_buf = team_buf;
            valid = m.team_rank == 0;              // one warp hands the chain back
        } else {
            c = blockIdx.x * cpb + (int)threadIdx.x / G;
            if (!keep_all && (c & ~(32 / G - 1)) >= nchains) return false;
            valid = c < nchains;
        }
        if (!valid && team == 1) c = nchains - 1;
        j = p.j0 + c;
        m.blob = p.blob;
        m.D = p.D;
        m.scratch = Model::kScratch ? smem + ((int)threadIdx.x / G) * kScratchRows * p.Dp : nullptr;
        Model::prepare(m);
        set_valid_mask<G, E>(g, m);
        logLmin = p.use_override ? p.logLmin_override : p.state->logLmin;
        nmcmc = p.use_override ? p.nmcmc_override : p.state->ctl[p.kind].nmcmc;
        const unsigned long long chain_id = p.chain_base + (unsigned long long)j;
        cid_lo = (uint32_t)chain_id;
        cid_hi = (uint32_t)(chain_id >> 32);
        // start point: given, the chain's own slot, or a uniformly drawn survivor (every survivor satisfies
        // logL > logLmin; the reference's slice sampler looks for such a member, sampler.py:471-478)
        start_slot = -1;
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
                const Philox4 r = philox4x32_10(cid_lo, cid_hi, 0xffffffffu, 0u, p.seed_lo, p.seed_hi);
                slot = p.order[p.n_skip + (int)u32_to_index(r.v[0], (uint32_t)(p.ens_n - p.n_skip))];
            }
            start_slot = slot;
            row = p.ens_x + (size_t)slot * p.Dp;
            logL = p.ens_logL[slot];
            logP = p.ens_logP[slot];
        }
        load_row<G, E>(g, row, p.D, x);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            blo[e] = (d < p.D) ? __ldg(p.lo + d) : CPN_NEG_INF;
            bhi[e] = (d < p.D) ? __ldg(p.hi + d) : -CPN_NEG_INF;
        }
        return true;
    }

    // `NestedSampler.consume_sample` sends a point down the pipe again when the sampler's reply does not beat the threshold
    // (cpnest/NestedSampling.py:354-357: `while not accepted: ... send(self.params[k])`), i.e. it asks for another sample
    // until one arrives that does.  A device chain that made no accepted step hands back its start point, a copy of a live
    // survivor; instead of returning it the chain starts again from another survivor (attempt a draws it from Philox block
    // (chain, 2^32 - 1 - a, 0); the chain's step counter - and with it its random stream - keeps running).
    // Only chains that draw their own start (START_SURVIVOR): a caller that supplies the start points (host protocol) sees
    // `accepted == 0` in the reply and resends itself, as the reference does.
    __device__ __forceinline__ bool can_restart(const ChainParams& p) const { return p.start_mode == START_SURVIVOR; }
    __device__ __forceinline__ void restart(const ChainParams& p, int attempt) {
        const Philox4 r = philox4x32_10(cid_lo, cid_hi, 0xffffffffu - (uint32_t)attempt, 0u, p.seed_lo, p.seed_hi);
        const int slot = p.order[p.n_skip + (int)u32_to_index(r.v[0], (uint32_t)(p.ens_n - p.n_skip))];
        start_slot = slot;
        logL = p.ens_logL[slot];
        logP = p.ens_logP[slot];
        load_row<G, E>(g, p.ens_x + (size_t)slot * p.Dp, p.D, x);
    }

    // the chain's point half-way, for the two-lag chain-length control (into every peer's staging block, like finish())
    __device__ __forceinline__ void snapshot_mid(const ChainParams& p) const {
        if (!valid) return;
        for (int o = 0; o < p.n_out; ++o) {
            if (p.out_mid_x[o] == nullptr) continue;
            double* row = p.out_mid_x[o] + (size_t)j * p.Dp;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                if (d < p.Dp) row[d] = (d < p.D) ? x[e] : 0.0;
            }
            if (g.lane == 0) p.out_mid_logL[o][j] = logL;
        }
    }

    // strict box test, model.py:33
    __device__ __forceinline__ bool in_box(const double (&v)[E]) const {
        bool inb = true;
#pragma unroll
        for (int e = 0; e < E; ++e) inb = inb & (blo[e] < v[e]) & (v[e] < bhi[e]);
        return g.all(inb);
    }

    // hand the evolved point back (`producer_pipe.send`, sampler.py:246), into every peer's staging block
    __device__ __forceinline__ void finish(const ChainParams& p, int steps, int accepted, int evals, int ne, int nc) const {
        if (!valid) return;
        for (int o = 0; o < p.n_out; ++o) {
            double* row = p.out_x[o] + (size_t)j * p.Dp;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                if (d < p.Dp) row[d] = (d < p.D) ? x[e] : 0.0;
            }
            if (g.lane == 0) {
                p.out_logL[o][j] = logL;
                p.out_logP[o][j] = logP;
                p.out_steps[o][j] = steps;
                p.out_acc[o][j] = accepted;
                if (p.out_start[o] != nullptr) p.out_start[o][j] = start_slot;
                if (p.out_evals[o] != nullptr) p.out_evals[o][j] = evals;
                if (p.out_ne[o] != nullptr) p.out_ne[o][j] = ne;
                if (p.out_nc[o] != nullptr) p.out_nc[o][j] = nc;
            }
        }
    }
};

}  // namespace cpn
