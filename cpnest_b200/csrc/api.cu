// C-ABI of cpnest_b200 (include/cpnest_b200.h): host-side driver of the device kernels.
// The host only enqueues work; the nested-sampling state (threshold, evidence, chain length,
// termination flag) lives on the device, so a run proceeds without per-batch host round trips.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <string>
#include <vector>

#include "cpnest_b200.h"
#include "ensemble.cuh"
#include "eval_kernel.cuh"
#include "launch.h"
#include "mh_kernel.cuh"
#include "ns_kernels.cuh"
#include "rtc.h"
#include "hmc_kernel.cuh"
#include "slice_kernel.cuh"

using namespace cpn;

static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define CKL() CK(cudaGetLastError())

typedef int (*mh_fn)(int, int, int, const MHParams&, cudaStream_t);
typedef int (*eval_fn)(int, int, const EvalParams&, cudaStream_t);
typedef int (*slice_fn)(int, int, int, const SliceParams&, cudaStream_t);
typedef int (*hmc_fn)(int, int, int, const HMCParams&, cudaStream_t);
static const hmc_fn kHMC[CPN_NUM_FUNCTORS] = {launch_hmc_gauss_iid, launch_hmc_gauss_corr, launch_hmc_eggbox,
                                              launch_hmc_rosenbrock, launch_hmc_ackley, launch_hmc_ring,
                                              launch_hmc_gauss_mean_sigma, launch_hmc_gauss_mixture, nullptr, launch_hmc_gauss_mix_nd};
static const slice_fn kSlice[CPN_NUM_FUNCTORS] = {launch_slice_gauss_iid, launch_slice_gauss_corr, launch_slice_eggbox,
                                                  launch_slice_rosenbrock, launch_slice_ackley, launch_slice_ring,
                                                  launch_slice_gauss_mean_sigma, launch_slice_gauss_mixture, nullptr, launch_slice_gauss_mix_nd};
static const mh_fn kMH[CPN_NUM_FUNCTORS] = {launch_mh_gauss_iid, launch_mh_gauss_corr, launch_mh_eggbox,
                                            launch_mh_rosenbrock, launch_mh_ackley, launch_mh_ring,
                                            launch_mh_gauss_mean_sigma, launch_mh_gauss_mixture, nullptr, launch_mh_gauss_mix_nd};
static const eval_fn kEval[CPN_NUM_FUNCTORS] = {launch_eval_gauss_iid, launch_eval_gauss_corr, launch_eval_eggbox,
                                                launch_eval_rosenbrock, launch_eval_ackley, launch_eval_ring,
                                                launch_eval_gauss_mean_sigma, launch_eval_gauss_mixture, nullptr, launch_eval_gauss_mix_nd};

struct cpn_ctx {
    cpn_config cfg;
    int D, Dp, N, G, E;
    cudaStream_t stream, own_stream;
    long long launches;
    double *d_lo, *d_hi, *d_blob;
    double *d_x, *d_logL, *d_logP;
    int* d_order[2];
    double* d_skeys[2];
    int cur;
    int Npad;
    double* d_sort_keys;
    int* d_sort_idx;
    // replacement staging: one block [x N*Dp | logL N | logP N | steps N | acc N | start N | evals N],
    // private by default, or a caller-provided (symmetric, peer-mapped) block for multi-GPU runs
    void* d_stage_own;
    void* stage_peer[kMaxPeers];
    int n_peers;                // 0: private staging; else == world_size, stage_peer[rank] is the local block
    int stage_half;             // which half of the peer-mapped blocks the current batch uses
    int rank, world;
    double *d_new_x, *d_new_logL, *d_new_logP;
    int *d_new_steps, *d_new_acc, *d_new_start, *d_new_evals, *d_new_ne, *d_new_nc, *d_rank;
    double *d_mid_x, *d_mid_logL;   // mid-chain snapshots of the batch (two-lag chain-length control), in the staging block
    double *d_rho, *d_rho2, *d_rho_mid, *d_lag_ema;
    double* d_corr_part;            // partial moments of k_chain_corr_partial, [D+2][segments][5]
    size_t corr_part_cap;
    double* d_snew;
    int* d_bucket;                  // ranking of the replacements: bucket histogram / starts [N+1], cursors [N+1], members [N]
    int* d_nle;                     // [N] #(survivors <= key) of every replacement
    double *d_start_x, *d_start_logL, *d_start_logP;
    double *d_mean, *d_cov, *d_covwork, *d_mean_partial, *d_cov_partial;
    // eigen-directions are double buffered: chains read buffer eig_cur while a refresh, running on the
    // side stream concurrently with the chain kernel, writes the other one
    double *d_eigval[2], *d_eigvec[2], *d_eigscale[2], *d_eigmean[2], *d_covbuf[2];
    int hmc_base_L;             // HamiltonianProposal.set_integration_parameters (proposal.py:524-536)
    double hmc_base_dt;
    int eig_cur;
    bool eig_pending, wait_cov;
    int eig_pending_age;        // batches started since the pending refresh was launched
    cudaStream_t side;
    cudaEvent_t ev_main, ev_cov, ev_eig;
    cudaStream_t acc_stream;    // evidence accumulation of a batch, beside its chain kernel
    cudaEvent_t ev_thr, ev_acc, ev_ins, ev_corr;
    bool acc_pending;           // an accumulation is in flight that the main stream has not waited for yet
    uint8_t* d_cycle;
    uint8_t h_cycle[4096];      // host copy: cycles up to kCycInline slots travel in the launch parameters of the MH kernel
    int cycle_len;
    uint8_t* d_slice_cycle;     // EnsembleSliceProposalCycle
    int slice_cycle_len;
    NSState* d_state;
    long long log_cap;          // capacity (rows) of the dead / history / insertion logs
    double *d_dead_x, *d_dead_logL, *d_dead_logP, *d_hist_logL, *d_hist_logvol, *d_ins;
    long long dead_upper;       // host upper bound of rows used in the logs
    long long batches_enqueued;
    unsigned long long chain_counter;
    long long chains_since_stats;
    long long eig_calls;
    LSE* d_trap_partial;
    double* d_scalar;
    double* d_rec;              // [N][D+2] record scratch of the host protocol
    int* d_sweeps;
    NSState* h_state;           // pinned
    int* h_idx;                 // pinned [2 * nlive]: index arguments of cpn_evolve_host_table on their way to the device
    cudaEvent_t ev;
    double* d_chain_log;        // chain history of the first chain_log_n chains of every MH batch (mh_kernel.cuh), or null
    unsigned long long* d_chain_log_count;
    int chain_log_n, chain_log_cap;
    RtcModule* rtc;             // kernels compiled at run time around a user-written functor (CPN_USER)
    bool user_flat_prior;
    bool live_ready;
    bool staged_external;       // replacements were injected by cpn_stage_replacements
    // sampler population: relative numbers of MH / slice / HMC chains in a batch (cpn_config.mix; a pure population
    // has one non-zero entry).  `primary` is the kind whose controller cpn_get_state reports.
    int mix[kNumKinds];
    int primary;
    bool mixed;
    cudaStream_t kind_stream[kNumKinds];   // the kinds' chain kernels of a batch run side by side
    cudaEvent_t ev_fork, ev_kind[kNumKinds];
    cudaStream_t chain_stream;  // stream of the chain launch in progress (null: c->stream)
    int last_kb[kNumKinds + 1]; // chain ranges of the kinds in the last evolved batch
};
static inline cudaStream_t chain_stream(cpn_ctx* c) { return c->chain_stream ? c->chain_stream : c->stream; }

// The K chains of a batch by sampler kind: kind s owns [kb[s], kb[s+1]).  Largest-remainder split of K in the
// proportions c->mix, at least one chain for every kind present (the reference gives every sampler process one
// of the nthreads worst points per iteration, cpnest/NestedSampling.py:324-342).
static int split_by_kind(const int* mix, int K, int* kb) {
    long long W = 0;
    int active = 0;
    for (int s = 0; s < kNumKinds; ++s) {
        if (mix[s] < 0) return fail("mix[%d] = %d must be >= 0", s, mix[s]);
        W += mix[s]; active += mix[s] > 0;
    }
    if (active == 0) return fail("mix must not be all zero");
    if (K < active) return fail("K=%d chains cannot hold %d sampler kinds", K, active);
    int n[kNumKinds];
    long long rem[kNumKinds];
    int used = 0;
    for (int s = 0; s < kNumKinds; ++s) {
        const long long q = (long long)K * mix[s];
        n[s] = (int)(q / W);
        rem[s] = q % W;
        if (mix[s] > 0 && n[s] == 0) { n[s] = 1; rem[s] = 0; }
        used += n[s];
    }
    while (used < K) {          // hand out what the floors left, largest remainder first (ties: lower kind)
        int best = -1;
        for (int s = 0; s < kNumKinds; ++s)
            if (mix[s] > 0 && (best < 0 || rem[s] > rem[best])) best = s;
        n[best] += 1; rem[best] -= W; used += 1;
    }
    while (used > K) {          // the minimum of one chain per kind overshot: take from the largest
        int best = 0;
        for (int s = 1; s < kNumKinds; ++s) if (n[s] > n[best]) best = s;
        n[best] -= 1; used -= 1;
    }
    kb[0] = 0;
    for (int s = 0; s < kNumKinds; ++s) kb[s + 1] = kb[s] + n[s];
    return 0;
}
static int kind_ranges(cpn_ctx* c, int K, int* kb) { return split_by_kind(c->mix, K, kb); }
extern "C" int cpn_kind_ranges(const int32_t* mix, int32_t K, int32_t* kb) {
    if (!mix || !kb) return fail("null argument");
    if (K < 1) return fail("K=%d must be >= 1", K);
    return split_by_kind(mix, K, kb);
}

// Host -> device copy that has LANDED when it returns.  cudaMemcpy from pageable memory may return once the data is
// staged, before the DMA into device memory has completed; that DMA is ordered in the legacy default stream, but the
// kernels of this library run on non-blocking streams, which the default stream does not order with.
static cudaError_t h2d_sync(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(0);
}

template <class T>
static int dalloc(T** p, size_t n) {
    CK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    CK(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    // the memset runs on the legacy default stream, the kernels on non-blocking streams: without this
    // the zeroing could land after a kernel has written the buffer
    CK(cudaStreamSynchronize(0));
    return 0;
}

// Slices of the ensemble the moments are summed over: a function of nlive alone (the shape of the partial sums, and with it
// every bit of the result, is the same on every rank and in every run); 32 up to nlive = 1e4, ~320 points per slice beyond
// (at nlive = 8e4 the 32-slice covariance took longer than the chain kernel it runs beside, and the merge had to wait for it:
// 0.25 ms on one batch in six)
static int stat_blocks(int N) { return std::min(256, std::max(32, (N + 319) / 320)); }

static const int kCfgs[][2] = {{1, 1}, {1, 2}, {1, 4}, {1, 8}, {4, 4},
#ifdef CPN_EXP_G4E16
                                {4, 16},
#endif
                                {8, 2}, {8, 8}, {16, 4}, {32, 1}, {32, 2}, {32, 4}, {32, 8}};
static const int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);

// smallest E instantiated for G lanes that covers D coordinates; 0 if none
static int min_E(int G, int D) {
    int best = 0;
    for (int i = 0; i < kNumCfgs; ++i)
        if (kCfgs[i][0] == G && G * kCfgs[i][1] >= D && (best == 0 || kCfgs[i][1] < best)) best = kCfgs[i][1];
    return best;
}

// (G, E) = lanes per chain, coordinates per lane.  With `lanes` given it is honoured; otherwise the
// narrowest group that still puts >= 8 warps on each of the 148 SMs for `chains` chains is chosen
// (narrow groups waste fewer issue slots on redundant scalar work, wide groups shorten the chain's
// critical path when there are too few chains to fill the machine).
static int pick_config(int D, int lanes, int functor, long long chains, int* G, int* E) {
    if (lanes > 0) {
        const int e = min_E(lanes, D);
        if (!e) return fail("no kernel configuration with %d lanes per chain covers dim=%d", lanes, D);
        *G = lanes; *E = e;
        return 0;
    }
    if (functor == CPN_GAUSS_MIXTURE && D <= 32) { *G = 32; *E = 1; return 0; }   // lanes split the data sum
    if (functor == CPN_GAUSS_CORR && D > 16) {
        // the D x D mat-vec of the likelihood is the step's cost.  Up to 64 dimensions: 16 lanes per chain, the eight chains of
        // a block are the columns of FP64 tensor-core products (models.cuh log_likelihood_block).  Beyond: the widest group
        // walks the mat-vec in the fewest dependent multiply-adds per lane.
        if (D <= 64 && min_E(16, D)) { *G = 16; *E = min_E(16, D); return 0; }
        const int e = min_E(32, D);
        if (e) { *G = 32; *E = e; return 0; }
    }
    // 33..64 dimensions: 16 lanes x 4 coordinates is the fastest MH configuration from half a wave of chains upwards
    // (D = 50: 0.092 ns per chain-step at K = 4704, 0.083 at K = 37632; 8 lanes: 0.099 / 0.092)
    if (D > 32 && D <= 64 && min_E(16, D) && chains * 16 / 32 >= 148LL * 4) { *G = 16; *E = min_E(16, D); return 0; }
    int gmax = 1;
    while (gmax < D && gmax < 32) gmax <<= 1;          // no more lanes than coordinates
    static const int Gs[] = {1, 4, 8, 16, 32};
    int bestG = 0, bestE = 0;
    for (int G_ : Gs) {
        if (G_ > gmax) break;
        const int e = min_E(G_, D);
        if (!e) continue;
        bestG = G_; bestE = e;
        if (chains * G_ / 32 >= 148LL * 8) break;
    }
    if (!bestG) return fail("dim=%d exceeds the largest kernel configuration (256)", D);
    *G = bestG; *E = bestE;
    return 0;
}

// ---- kernel dispatch: built-in functors through the launcher tables, CPN_USER through the NVRTC module -----------
static int user_launch(cpn_ctx* c, const char* kind, int G, int E, const void* params, int n, int Dp, int threads) {
    if (n <= 0) return 0;
    const int cpb = threads / G;
    const int nb = (n + cpb - 1) / cpb;
    const size_t sm = (size_t)cpb * (kScratchRows + 1) * Dp * sizeof(double);     // functor scratch + the HMC kernel's mat-vec row
    char err[4096];
    if (rtc_launch(c->rtc, kind, G, E, params, nb, threads, sm, chain_stream(c), err, sizeof err)) return fail("%s", err);
    return 0;
}
static int run_eval(cpn_ctx* c, int G, int E, const EvalParams& p) {
    if (c->cfg.functor == CPN_USER) return user_launch(c, "eval", G, E, &p, p.n, p.Dp, kEvalThreads);
    if (kEval[c->cfg.functor](G, E, p, c->stream)) return fail("no eval kernel for G=%d E=%d", G, E);
    return 0;
}
static int run_mh(cpn_ctx* c, int G, int E, int replay, const MHParams& p) {
    if (c->cfg.functor == CPN_USER) {
        if (replay) return fail("replay entry points are not available for a user functor");
        if (p.cycle_len > kCycInline) return fail("a user functor takes proposal cycles of up to %d entries (got %d)", kCycInline, p.cycle_len);
        return user_launch(c, "mh", G, E, &p, p.j1 - p.j0, p.Dp, kMHThreads);
    }
    if (kMH[c->cfg.functor](G, E, replay, p, chain_stream(c))) return fail("no MH kernel for G=%d E=%d", G, E);
    return 0;
}
static int run_slice(cpn_ctx* c, int G, int E, int replay, const SliceParams& p) {
    if (c->cfg.functor == CPN_USER) {
        if (replay) return fail("replay entry points are not available for a user functor");
        return user_launch(c, "slice", G, E, &p, p.j1 - p.j0, p.Dp, kMHThreads);
    }
    if (kSlice[c->cfg.functor](G, E, replay, p, chain_stream(c))) return fail("no slice kernel for G=%d E=%d", G, E);
    return 0;
}
static int run_hmc(cpn_ctx* c, int G, int E, int replay, const HMCParams& p) {
    if (c->cfg.functor == CPN_USER) {
        if (replay) return fail("replay entry points are not available for a user functor");
        return user_launch(c, "hmc", G, E, &p, p.j1 - p.j0, p.Dp, kMHThreads);
    }
    if (kHMC[c->cfg.functor](G, E, replay, p, chain_stream(c))) return fail("no HMC kernel for G=%d E=%d", G, E);
    return 0;
}

extern "C" const char* cpn_last_error(void) { return g_err.c_str(); }

extern "C" int cpn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static int grow_logs(cpn_ctx* c, long long need) {
    if (need <= c->log_cap) return 0;
    long long cap = std::max(need, c->log_cap * 2);
    double *nx, *nl, *np, *hl, *hv, *ni;
    CK(cudaMalloc((void**)&nx, (size_t)cap * c->Dp * sizeof(double)));
    CK(cudaMalloc((void**)&nl, (size_t)cap * sizeof(double)));
    CK(cudaMalloc((void**)&np, (size_t)cap * sizeof(double)));
    CK(cudaMalloc((void**)&hl, (size_t)cap * sizeof(double)));
    CK(cudaMalloc((void**)&hv, (size_t)cap * sizeof(double)));
    CK(cudaMalloc((void**)&ni, (size_t)cap * sizeof(double)));
    CK(cudaStreamSynchronize(c->stream));
    if (c->log_cap > 0) {
        const size_t n = (size_t)c->log_cap;
        CK(cudaMemcpy(nx, c->d_dead_x, n * c->Dp * sizeof(double), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(nl, c->d_dead_logL, n * sizeof(double), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(np, c->d_dead_logP, n * sizeof(double), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(hl, c->d_hist_logL, n * sizeof(double), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(hv, c->d_hist_logvol, n * sizeof(double), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(ni, c->d_ins, n * sizeof(double), cudaMemcpyDeviceToDevice));
        cudaFree(c->d_dead_x); cudaFree(c->d_dead_logL); cudaFree(c->d_dead_logP);
        cudaFree(c->d_hist_logL); cudaFree(c->d_hist_logvol); cudaFree(c->d_ins);
    }
    c->d_dead_x = nx; c->d_dead_logL = nl; c->d_dead_logP = np;
    c->d_hist_logL = hl; c->d_hist_logvol = hv; c->d_ins = ni;
    c->log_cap = cap;
    return 0;
}

// everything that reads what k_ns_accumulate writes (counters, evidence, stop flag) first joins its stream
static int join_acc(cpn_ctx* c) {
    if (c->acc_pending) {
        CK(cudaStreamWaitEvent(c->stream, c->ev_acc, 0));
        c->acc_pending = false;
    }
    return 0;
}

static int fetch_state(cpn_ctx* c) {
    if (join_acc(c)) return 1;
    CK(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(NSState), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->dead_upper = c->h_state->iteration;
    return 0;
}

static int push_state(cpn_ctx* c) {
    if (join_acc(c)) return 1;
    CK(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(NSState), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

static void init_state(cpn_ctx* c, NSState* s) {
    memset(s, 0, sizeof *s);
    s->logZ = -INFINITY;
    s->B = 0.0;
    s->info = 0.0;
    s->logw = 0.0;
    s->logLmin = -INFINITY;
    s->logLmax = -INFINITY;
    s->condition = INFINITY;
    s->tolerance = 0.1;                                   // NestedSampler(stopping=0.1)
    const int n0 = c->cfg.nmcmc_init > 0 ? c->cfg.nmcmc_init : std::max(1, c->cfg.maxmcmc / 10);   // sampler.py:79
    for (int k = 0; k < kNumKinds; ++k) {
        s->ctl[k].nmcmc = std::min(n0, c->cfg.maxmcmc);
        s->ctl[k].nmcmc_exact = (double)s->ctl[k].nmcmc;
        // measured on the 50-D Gaussian (nlive 1e4, 8 seeds each): MH chains are unbiased at 0.1; slice chains show
        // +1.0 sigma (+0.09 nats) mean deviation of logZ at 0.1 and none at 0.05
        s->ctl[k].rho_target = c->cfg.rho_target > 0 ? c->cfg.rho_target : (k == CPN_SAMPLER_SLICE ? 0.05 : 0.1);
        s->ctl[k].rho_max = -1.0;
        s->ctl[k].rho_decay = -1.0;
        s->ctl[k].ne_ref = 0.0;
        s->ctl[k].rho_ref = 0.0;
    }
    s->nmcmc_safety = c->cfg.nmcmc_safety > 0 ? c->cfg.nmcmc_safety : 5.0;
    s->nmcmc_tau = (double)c->N / s->nmcmc_safety;        // tau = poolsize / safety (sampler.py:166)
    s->maxmcmc = c->cfg.maxmcmc;
    s->adapt = c->cfg.adapt_nmcmc;
    s->slice_mu = 1.0;                                    // SliceSampler.reset (sampler.py:422)
    s->hmc_dt = c->hmc_base_dt;                           // dt = base_dt * scale, scale = 1 (proposal.py:395-396, 550)
    s->sampler = c->primary;
}

static size_t stage_off(int N, int Dp, int which) {
    // byte offsets of the arrays inside a staging block for N chains
    size_t o = 0;
    const size_t sizes[11] = {(size_t)N * Dp * 8, (size_t)N * 8, (size_t)N * 8, (size_t)N * 4, (size_t)N * 4, (size_t)N * 4, (size_t)N * 4,
                              (size_t)N * 4, (size_t)N * 4, (size_t)N * Dp * 8, (size_t)N * 8};
    for (int i = 0; i < which; ++i) o += (sizes[i] + 255) & ~(size_t)255;
    return o;
}
static size_t stage_bytes(int N, int Dp) { return stage_off(N, Dp, 11); }
static void bind_staging(cpn_ctx* c, void* base) {
    char* b = (char*)base;
    c->d_new_x = (double*)(b + stage_off(c->N, c->Dp, 0));
    c->d_new_logL = (double*)(b + stage_off(c->N, c->Dp, 1));
    c->d_new_logP = (double*)(b + stage_off(c->N, c->Dp, 2));
    c->d_new_steps = (int*)(b + stage_off(c->N, c->Dp, 3));
    c->d_new_acc = (int*)(b + stage_off(c->N, c->Dp, 4));
    c->d_new_start = (int*)(b + stage_off(c->N, c->Dp, 5));
    c->d_new_evals = (int*)(b + stage_off(c->N, c->Dp, 6));
    c->d_new_ne = (int*)(b + stage_off(c->N, c->Dp, 7));
    c->d_new_nc = (int*)(b + stage_off(c->N, c->Dp, 8));
    c->d_mid_x = (double*)(b + stage_off(c->N, c->Dp, 9));
    c->d_mid_logL = (double*)(b + stage_off(c->N, c->Dp, 10));
}

// two halves: batch b uses half b & 1, so that a fast rank's chains of batch b+1 never store into a
// block that a slow rank is still reading for batch b (one barrier per batch is then enough)
extern "C" int64_t cpn_staging_bytes(cpn_ctx* c) { return c ? 2 * (int64_t)stage_bytes(c->N, c->Dp) : -1; }

// Multi-GPU: `peers[r]` is the address, valid on THIS device, of rank r's staging block (peer-mapped
// symmetric memory; peers[rank] is the local block).  From then on every chain writes its outcome
// into all blocks, i.e. the exchange of the replacements is done by the chain kernel itself.
extern "C" int cpn_set_staging(cpn_ctx* c, void* const* peers, int32_t npeers) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    if (!peers || npeers == 0) {
        c->n_peers = 0;
        bind_staging(c, c->d_stage_own);
        return 0;
    }
    if (npeers != c->world) return fail("cpn_set_staging: %d peer blocks given, world size is %d", npeers, c->world);
    for (int r = 0; r < npeers; ++r) {
        if (!peers[r]) return fail("cpn_set_staging: null block for rank %d", r);
        c->stage_peer[r] = peers[r];
    }
    c->n_peers = npeers;
    c->stage_half = 0;
    bind_staging(c, peers[c->rank]);
    return 0;
}

static int create_impl(cpn_ctx* c, const cpn_config* cfg, const double* lo, const double* hi, const void* blob, size_t blob_bytes,
                       size_t nb);

extern "C" int cpn_create(const cpn_config* cfg, const double* lo, const double* hi, const void* blob,
                          size_t blob_bytes, cpn_ctx** out) {
    if (!cfg || !lo || !hi || !out) return fail("cpn_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("cpn_create: no CUDA device available (cpnest_b200 has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("cpn_create: device %d out of range (%d devices)", cfg->device, ndev);
    if (cfg->dim < 1) return fail("cpn_create: dim must be >= 1");
    if (cfg->nlive < 4) return fail("cpn_create: nlive must be >= 4");
    if (cfg->maxmcmc < 1) return fail("cpn_create: maxmcmc must be >= 1");
    if (cfg->functor < 0 || cfg->functor >= CPN_NUM_FUNCTORS) return fail("cpn_create: unknown functor %d", cfg->functor);
    if (cfg->sampler < CPN_SAMPLER_MH || cfg->sampler > CPN_SAMPLER_HMC) return fail("cpn_create: unknown sampler %d", cfg->sampler);
    for (int k = 0; k < kNumKinds; ++k)
        if (cfg->mix[k] < 0) return fail("cpn_create: mix[%d] = %d must be >= 0", k, cfg->mix[k]);
    for (int d = 0; d < cfg->dim; ++d)
        if (!(lo[d] < hi[d])) return fail("cpn_create: bounds[%d] = (%g, %g) is empty", d, lo[d], hi[d]);
    // functor-specific blob checks
    const size_t nb = blob_bytes / sizeof(double);
    switch (cfg->functor) {
        case CPN_GAUSS_IID: if (nb != 3) return fail("gauss_iid blob must be {mu, 1/sigma, -log(sigma)-log(2pi)/2}"); break;
        case CPN_GAUSS_CORR:
            if (nb != 1 + (size_t)cfg->dim * cfg->dim + (size_t)cfg->dim * ((cfg->dim + 31) & ~31))
                return fail("gauss_corr blob must be {logdet, Linv[D*D], LinvT[D][S]}, S = D rounded up to 32");
            break;
        case CPN_RING: if (nb != 2 || cfg->dim != 2) return fail("ring needs dim=2 and blob {radius, thickness}"); break;
        case CPN_GAUSS_MEAN_SIGMA: if (cfg->dim != 2) return fail("gauss_mean_sigma needs dim=2"); break;
        case CPN_GAUSS_MIXTURE:
            if (cfg->dim != 5 || nb < 2 || (size_t)((const double*)blob)[0] + 1 != nb) return fail("gauss_mixture needs dim=5 and blob {n, data[n]}");
            break;
        case CPN_GAUSS_MIX_ND: if (nb != 5) return fail("gauss_mix_nd blob must be {w, m1, s1, m2, s2}"); break;
        default: break;
    }
    CK(cudaSetDevice(cfg->device));
    cpn_ctx* c = new cpn_ctx();
    memset(c, 0, sizeof *c);
    c->cfg = *cfg;
    // every failure below leaves through one place that releases what has been allocated so far
    if (create_impl(c, cfg, lo, hi, blob, blob_bytes, nb)) {
        const std::string keep = g_err;
        cpn_destroy(c);
        g_err = keep;
        return 1;
    }
    *out = c;
    return 0;
}

static int create_impl(cpn_ctx* c, const cpn_config* cfg, const double* lo, const double* hi, const void* blob, size_t blob_bytes,
                       size_t nb) {
    c->D = cfg->dim;
    c->Dp = (cfg->dim + 3) & ~3;
    // Row padding.  When the rows are zero-padded up to G*E doubles, every slot a lane group gathers lies inside its own
    // row and the MH kernel needs no masks for the slots beyond D (mh_kernel.cuh).  Taken when it costs at most a third
    // more memory, for the functor that profits (iid Gaussian centred at 0: Model::pad_ok): D = 50 -> rows of 64 doubles.
    if (cfg->functor == CPN_GAUSS_IID && blob && blob_bytes >= sizeof(double) && ((const double*)blob)[0] == 0.0 &&
        !getenv("CPN_NO_ROW_PAD")) {
        int pmin = 0;
        for (int i = 0; i < kNumCfgs; ++i) {
            const int prod = kCfgs[i][0] * kCfgs[i][1];
            if (prod >= cfg->dim && (pmin == 0 || prod < pmin)) pmin = prod;
        }
        if (pmin >= 4 && 3 * pmin <= 4 * c->Dp) c->Dp = pmin;
    }
    c->N = cfg->nlive;
    {
        int active = 0;
        for (int k = 0; k < kNumKinds; ++k) { c->mix[k] = cfg->mix[k]; active += cfg->mix[k] > 0; }
        if (active == 0) { c->mix[cfg->sampler] = 1; active = 1; }
        c->mixed = active > 1;
        c->primary = cfg->sampler;
        if (c->mix[c->primary] == 0)
            for (int k = kNumKinds - 1; k >= 0; --k) if (c->mix[k] > 0) c->primary = k;
        c->cfg.sampler = c->primary;
    }
    if (pick_config(c->D, cfg->lanes, cfg->functor, c->N, &c->G, &c->E)) return 1;
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_cov, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_eig, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->acc_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_thr, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_acc, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_ins, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_corr, cudaEventDisableTiming));
    if (c->mixed) {
        CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        for (int k = 0; k < kNumKinds; ++k) {
            CK(cudaStreamCreateWithFlags(&c->kind_stream[k], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c->ev_kind[k], cudaEventDisableTiming));
        }
    }
    const int N = c->N, D = c->D, Dp = c->Dp;
    if (dalloc(&c->d_lo, D) || dalloc(&c->d_hi, D) || dalloc(&c->d_blob, std::max<size_t>(nb, 1))) return 1;
    CK(h2d_sync(c->d_lo, lo, D * sizeof(double)));
    CK(h2d_sync(c->d_hi, hi, D * sizeof(double)));
    if (nb) CK(h2d_sync(c->d_blob, blob, nb * sizeof(double)));
    // the live rows and, behind them, the two buffers of eigen-directions: one array, so the MH kernel gathers every row of a
    // step from one base with a 32-bit row number (MHParams::eig_row0)
    if (dalloc(&c->d_x, (size_t)(N + 2 * D) * Dp + kRowSlack) || dalloc(&c->d_logL, N) || dalloc(&c->d_logP, N)) return 1;
    c->d_eigvec[0] = c->d_x + (size_t)N * Dp;
    c->d_eigvec[1] = c->d_x + (size_t)(N + D) * Dp;
    for (int b = 0; b < 2; ++b)
        if (dalloc(&c->d_order[b], N) || dalloc(&c->d_skeys[b], N)) return 1;
    c->Npad = 1;
    while (c->Npad < N) c->Npad <<= 1;
    if (dalloc(&c->d_sort_keys, c->Npad) || dalloc(&c->d_sort_idx, c->Npad)) return 1;
    {
        char* blk;
        if (dalloc(&blk, (size_t)stage_bytes(N, Dp))) return 1;
        c->d_stage_own = blk;
        bind_staging(c, blk);
    }
    c->rank = std::max(0, cfg->rank);
    c->world = std::max(1, cfg->world_size);
    if (c->world > kMaxPeers || c->rank >= c->world) return fail("cpn_create: rank %d / world_size %d not supported (max %d)", c->rank, c->world, kMaxPeers);
    if (dalloc(&c->d_rho, kNumKinds * (D + 1)) || dalloc(&c->d_rho2, kNumKinds * (D + 1)) || dalloc(&c->d_rank, N) || dalloc(&c->d_snew, N) ||
        dalloc(&c->d_rho_mid, kNumKinds * (D + 1)) || dalloc(&c->d_lag_ema, 2 * kNumKinds * (D + 1)))
        return 1;
    if (dalloc(&c->d_bucket, 3 * (size_t)(N + 1) + (size_t)(N / kScanTile + 2)) || dalloc(&c->d_nle, N)) return 1;
    c->corr_part_cap = (size_t)(D + 2) * ((N + kCorrSeg - 1) / kCorrSeg + 1) * kCorrMom;
    if (dalloc(&c->d_corr_part, c->corr_part_cap)) return 1;
    if (dalloc(&c->d_start_x, (size_t)N * Dp) || dalloc(&c->d_start_logL, N) || dalloc(&c->d_start_logP, N)) return 1;
    if (dalloc(&c->d_mean, D) || dalloc(&c->d_cov, (size_t)D * D) || dalloc(&c->d_covwork, 4 * (size_t)D * (D | 1)) ||
        dalloc(&c->d_eigval[0], D) || dalloc(&c->d_eigscale[0], D) ||
        dalloc(&c->d_eigval[1], D) || dalloc(&c->d_eigscale[1], D) ||
        dalloc(&c->d_eigmean[0], D) || dalloc(&c->d_eigmean[1], D) ||
        dalloc(&c->d_covbuf[0], (size_t)D * D) || dalloc(&c->d_covbuf[1], (size_t)D * D) || dalloc(&c->d_mean_partial, (size_t)stat_blocks(N) * D) ||
        dalloc(&c->d_cov_partial, (size_t)stat_blocks(N) * D * D))
        return 1;
    if (dalloc(&c->d_state, 1) || dalloc(&c->d_trap_partial, 256) || dalloc(&c->d_scalar, 4) ||
        dalloc(&c->d_rec, (size_t)N * (D + 2)) || dalloc(&c->d_sweeps, 1))
        return 1;
    CK(cudaMallocHost((void**)&c->h_state, sizeof(NSState)));
    CK(cudaMallocHost((void**)&c->h_idx, 2 * (size_t)N * sizeof(int)));
    // DefaultProposalCycle weights 5:5:10:10 laid out deterministically until cpn_set_cycle is called
    {
        std::vector<uint8_t> cyc;
        const uint8_t pat[6] = {CPN_WALK, CPN_STRETCH, CPN_DE, CPN_DE, CPN_EIGEN, CPN_EIGEN};
        for (int i = 0; i < 96; ++i) cyc.push_back(pat[i % 6]);
        c->cycle_len = (int)cyc.size();
        memcpy(c->h_cycle, cyc.data(), cyc.size());
        if (dalloc(&c->d_cycle, 4096)) return 1;
        CK(h2d_sync(c->d_cycle, cyc.data(), cyc.size()));
    }
    {
        // EnsembleSliceProposalCycle: equal weights (proposal.py:195), laid out deterministically until
        // cpn_set_slice_cycle is called
        std::vector<uint8_t> cyc;
        for (int i = 0; i < 99; ++i) cyc.push_back((uint8_t)(i % 3));
        c->slice_cycle_len = (int)cyc.size();
        if (dalloc(&c->d_slice_cycle, 4096)) return 1;
        CK(h2d_sync(c->d_slice_cycle, cyc.data(), cyc.size()));
    }
    {
        // base_L = 10 + int((h/l) d^(1/4)), base_dt = (1/base_L) l/h with l, h the shortest and longest prior range
        double l = INFINITY, h = 0.0;
        for (int d = 0; d < D; ++d) { l = std::min(l, hi[d] - lo[d]); h = std::max(h, hi[d] - lo[d]); }
        c->hmc_base_L = 10 + (int)((h / l) * std::pow((double)D, 0.25));
        c->hmc_base_dt = (1.0 / c->hmc_base_L) * l / h;
    }
    init_state(c, c->h_state);
    if (push_state(c)) return 1;
    // Dead-point log: a run produces about nlive * (H + a few sqrt(H)) dead points; HBM is plentiful
    // (180 GB), re-allocation mid-run is not free (allocation + copy + implicit synchronisation), so
    // reserve 128 * nlive rows up front (bounded by 8 GB) and double beyond that.
    {
        long long rows = 128LL * N;
        const long long cap_rows = (8LL << 30) / ((long long)Dp * 8 + 40);
        rows = std::max<long long>(std::min(rows, cap_rows), 4LL * N);
        if (grow_logs(c, rows)) return 1;
    }
    return 0;
}

extern "C" void cpn_destroy(cpn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    // (a context whose creation failed half-way has null handles: nothing to wait for / destroy there)
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->side) cudaStreamSynchronize(c->side);
    if (c->acc_stream) cudaStreamSynchronize(c->acc_stream);
    void* ptrs[] = {c->d_lo, c->d_hi, c->d_blob, c->d_x, c->d_logL, c->d_logP, c->d_order[0], c->d_order[1],
                    c->d_skeys[0], c->d_skeys[1], c->d_sort_keys, c->d_sort_idx, c->d_stage_own, c->d_rho, c->d_rho2, c->d_rho_mid, c->d_lag_ema, c->d_rank, c->d_snew, c->d_bucket, c->d_nle, c->d_corr_part, c->d_start_x, c->d_start_logL,
                    c->d_start_logP, c->d_mean, c->d_cov, c->d_covwork, c->d_eigval[0], c->d_eigscale[0], c->d_eigval[1], c->d_eigscale[1],
                    c->d_mean_partial, c->d_cov_partial, c->d_cycle, c->d_slice_cycle, c->d_eigmean[0], c->d_eigmean[1], c->d_covbuf[0], c->d_covbuf[1], c->d_state, c->d_dead_x, c->d_dead_logL,
                    c->d_dead_logP, c->d_hist_logL, c->d_hist_logvol, c->d_ins, c->d_trap_partial, c->d_scalar, c->d_rec};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->d_chain_log) cudaFree(c->d_chain_log);
    if (c->d_chain_log_count) cudaFree(c->d_chain_log_count);
    if (c->rtc) rtc_destroy(c->rtc);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_idx) cudaFreeHost(c->h_idx);
    cudaEvent_t evs[] = {c->ev, c->ev_main, c->ev_cov, c->ev_eig, c->ev_thr, c->ev_acc, c->ev_ins, c->ev_corr, c->ev_fork,
                         c->ev_kind[0], c->ev_kind[1], c->ev_kind[2]};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    cudaStream_t sts[] = {c->side, c->acc_stream, c->kind_stream[0], c->kind_stream[1], c->kind_stream[2], c->own_stream};
    for (cudaStream_t st : sts)
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    cudaGetLastError();
    delete c;
}

extern "C" int cpn_set_cycle(cpn_ctx* c, const uint8_t* cycle, int32_t len) {
    if (!c || !cycle || len < 1 || len > 4096) return fail("cpn_set_cycle: bad arguments");
    for (int i = 0; i < len; ++i)
        if (cycle[i] > CPN_EIGEN) return fail("cpn_set_cycle: unknown proposal id %d at %d", cycle[i], i);
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMemcpyAsync(c->d_cycle, cycle, len, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memcpy(c->h_cycle, cycle, len);
    c->cycle_len = len;
    return 0;
}

extern "C" int cpn_set_slice_cycle(cpn_ctx* c, const uint8_t* cycle, int32_t len) {
    if (!c || !cycle || len < 1 || len > 4096) return fail("cpn_set_slice_cycle: bad arguments");
    for (int i = 0; i < len; ++i)
        if (cycle[i] > CPN_SLICE_CORR) return fail("cpn_set_slice_cycle: unknown direction id %d at %d", cycle[i], i);
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMemcpyAsync(c->d_slice_cycle, cycle, len, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->slice_cycle_len = len;
    return 0;
}

// Model.log_likelihood / log_prior written by the user in CUDA C++ (csrc/user_model.cuh states the contract): the chain
// kernels are instantiated around it with NVRTC for sm_100a, for every lane configuration this context can pick.
extern "C" int cpn_set_user_functor(cpn_ctx* c, const char* source, const char* const* include_dirs, int32_t n_dirs) {
    if (!c || !source) return fail("null argument");
    if (c->cfg.functor != CPN_USER) return fail("cpn_set_user_functor: the context was not created with functor CPN_USER");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaFree(0));                                  // make sure the primary context is current for the driver API
    std::vector<std::array<int, 2>> cfgs;
    static const int Gs[] = {1, 4, 8, 16, 32};
    int gmax = 1;
    while (gmax < c->D && gmax < 32) gmax <<= 1;
    if (c->cfg.lanes > 0) {
        const int e = min_E(c->cfg.lanes, c->D);
        if (!e) return fail("no kernel configuration with %d lanes per chain covers dim=%d", c->cfg.lanes, c->D);
        cfgs.push_back({c->cfg.lanes, e});
    } else {
        for (int G_ : Gs) {
            if (G_ > gmax) break;
            const int e = min_E(G_, c->D);
            if (e) cfgs.push_back({G_, e});
        }
    }
    bool have = false;
    for (auto& g : cfgs) have = have || (g[0] == c->G && g[1] == c->E);
    if (!have) cfgs.push_back({c->G, c->E});
    int kinds = 1 | 2;                                // eval + MH (the prior burn-in of Sampler.reset runs MH chains)
    if (c->mix[CPN_SAMPLER_SLICE] > 0) kinds |= 4;
    if (c->mix[CPN_SAMPLER_HMC] > 0) kinds |= 8;
    std::vector<int> flat;
    for (auto& g : cfgs) { flat.push_back(g[0]); flat.push_back(g[1]); }
    char err[8192];
    err[0] = 0;
    if (c->rtc) { rtc_destroy(c->rtc); c->rtc = nullptr; }
    c->rtc = rtc_compile(source, include_dirs, n_dirs, (const int(*)[2])flat.data(), (int)cfgs.size(), kinds, err, sizeof err);
    if (!c->rtc) return fail("%s", err);
    c->user_flat_prior = strstr(source, "CPN_USER_HAS_PRIOR") == nullptr;
    return 0;
}

extern "C" int cpn_set_seed(cpn_ctx* c, uint64_t seed) {
    if (!c) return fail("null context");
    c->cfg.seed = seed;
    return 0;
}

extern "C" int cpn_set_slice_mu(cpn_ctx* c, double mu) {
    if (!c) return fail("null context");
    if (!(mu > 0.0)) return fail("cpn_set_slice_mu: mu must be > 0");
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    c->h_state->slice_mu = mu;
    return push_state(c);
}

// ---- sort the live set by logL (OrderedLivePoints.__init__) -------------------------------------------
static int sort_live(cpn_ctx* c) {
    const int T = 256, Np = c->Npad;
    k_sort_fill<<<(Np + T - 1) / T, T, 0, c->stream>>>(c->d_logL, c->N, Np, c->d_sort_keys, c->d_sort_idx);
    for (int k = 2; k <= Np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1)
            k_bitonic_step<<<(Np + T - 1) / T, T, 0, c->stream>>>(c->d_sort_keys, c->d_sort_idx, Np, j, k);
    CKL();
    CK(cudaMemcpyAsync(c->d_skeys[c->cur], c->d_sort_keys, c->N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_order[c->cur], c->d_sort_idx, c->N * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

// mean and covariance of the live ensemble (for eigen buffer `buf`) on stream `st`
static int launch_moments(cpn_ctx* c, cudaStream_t st, int buf) {
    const int D = c->D, N = c->N;
    const int kStatBlocks = stat_blocks(N);
    if (D > kStatThreads) return fail("ensemble statistics support dim <= %d", kStatThreads);
    k_mean_partial<<<kStatBlocks, kStatThreads, 0, st>>>(c->d_x, N, D, c->Dp, c->d_mean_partial);
    k_mean_final<<<(D + 127) / 128, 128, 0, st>>>(c->d_mean_partial, kStatBlocks, N, D, c->d_mean);
    CK(cudaMemcpyAsync(c->d_eigmean[buf], c->d_mean, D * sizeof(double), cudaMemcpyDeviceToDevice, st));
    const size_t cov_sm = (size_t)kMmaStage * cov_mma_ld(D) * sizeof(double);
    if (cov_sm > 48 * 1024) CK(cudaFuncSetAttribute(k_cov_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cov_sm));
    k_cov_partial<<<kStatBlocks, kStatThreads, cov_sm, st>>>(
        c->d_x, c->d_mean, N, D, c->Dp, c->d_cov_partial);
    k_cov_final<<<(D * D + 127) / 128, 128, 0, st>>>(c->d_cov_partial, kStatBlocks, N, D, c->d_cov);
    CK(cudaMemcpyAsync(c->d_covbuf[buf], c->d_cov, (size_t)D * D * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CKL();
    c->launches += 4;
    return 0;
}

// eigen-decomposition of the covariance launch_moments left in d_cov, into eigen buffer `buf`, on stream `st`
static int launch_eigen(cpn_ctx* c, cudaStream_t st, int buf, bool beside_chains) {
    const int D = c->D;
    const size_t sm = 4 * (size_t)D * (D | 1) * sizeof(double);
    const int use_smem = sm <= 200 * 1024;
    if (use_smem && sm > 40 * 1024)
        CK(cudaFuncSetAttribute(k_jacobi_eigh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    c->eig_calls += 1;
    // beside a running chain kernel only a small block finds room on an SM (registers); alone, use them all
    const int jt = beside_chains ? 256 : kJacobiThreads;
    // (warm start from the eigenvectors in use, the other buffer)
    k_jacobi_eigh<<<1, jt, use_smem ? sm : 0, st>>>(c->d_cov, c->d_covwork, use_smem, D, c->Dp,
                                                               (c->cfg.flags & CPN_COLD_EIGEN) ? nullptr : c->d_eigvec[buf ^ 1],
                                                               // beside the chain kernels the block slows its SM's chain blocks for as long
                                                               // as it runs (0.09 ms per batch it overlaps): five sweeps, one batch
                                                               beside_chains ? 5 : 30,
                                                               c->d_eigval[buf], c->d_eigvec[buf], c->d_eigscale[buf], c->d_sweeps);
    CKL();
    if (beside_chains) CK(cudaEventRecord(c->ev_eig, st));
    c->chains_since_stats = 0;
    c->launches += 1;
    return 0;
}

// a refresh started on the side stream during the previous batch becomes current
static int adopt_pending_stats(cpn_ctx* c) {
    if (!c->eig_pending) return 0;
    CK(cudaStreamWaitEvent(c->stream, c->ev_eig, 0));
    c->eig_cur ^= 1;
    c->eig_pending = false;
    return 0;
}

extern "C" int cpn_update_ensemble_stats(cpn_ctx* c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    if (adopt_pending_stats(c)) return 1;
    if (launch_moments(c, c->stream, c->eig_cur ^ 1) || launch_eigen(c, c->stream, c->eig_cur ^ 1, false)) return 1;
    c->eig_cur ^= 1;
    return 0;
}

// refresh on the side stream: reads the live set as it is now (after everything already enqueued on
// the main stream), overlaps with the chain kernels of the coming batches, is adopted by a later batch.
// (Measured: a refresh slows the two chain kernels it runs beside by 0.09 ms each - the eigen-decomposition's block
// shares an SM with three chain blocks for 1 ms.  Running the moments alone on the main stream with one block per SM did
// not change that and cost 55 us per refresh.)
static int start_async_stats(cpn_ctx* c) {
    CK(cudaEventRecord(c->ev_main, c->stream));
    CK(cudaStreamWaitEvent(c->side, c->ev_main, 0));
    if (launch_moments(c, c->side, c->eig_cur ^ 1)) return 1;
    CK(cudaEventRecord(c->ev_cov, c->side));     // the live rows have been read
    if (launch_eigen(c, c->side, c->eig_cur ^ 1, true)) return 1;
    c->eig_pending = true;
    c->eig_pending_age = 0;
    c->wait_cov = true;
    return 0;
}

extern "C" int cpn_get_ensemble_stats(cpn_ctx* c, double* mean, double* cov, double* eigval, double* eigvec) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    if (adopt_pending_stats(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    const int D = c->D;
    if (getenv("CPN_DEBUG")) {
        int sw = -1;
        CK(cudaMemcpy(&sw, c->d_sweeps, sizeof(int), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[cpn] jacobi sweeps of the last refresh: %d (refresh #%lld)\n", sw, c->eig_calls);
    }
    if (mean) CK(cudaMemcpy(mean, c->d_mean, D * sizeof(double), cudaMemcpyDeviceToHost));
    if (cov) CK(cudaMemcpy(cov, c->d_cov, (size_t)D * D * sizeof(double), cudaMemcpyDeviceToHost));
    if (eigval) CK(cudaMemcpy(eigval, c->d_eigval[c->eig_cur], D * sizeof(double), cudaMemcpyDeviceToHost));
    if (eigvec) {
        // numpy layout: eigvec[k][i] = component k of eigenvector i
        std::vector<double> rows((size_t)D * c->Dp);
        CK(cudaMemcpy(rows.data(), c->d_eigvec[c->eig_cur], rows.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < D; ++i)
            for (int k = 0; k < D; ++k) eigvec[(size_t)k * D + i] = rows[(size_t)i * c->Dp + k];
    }
    return 0;
}

// what every chain kernel needs: the ensemble, the box, the chain control and the staging outputs
static void fill_common(cpn_ctx* c, ChainParams* p) {
    p->ens_x = c->d_x; p->ens_logL = c->d_logL; p->ens_logP = c->d_logP;
    p->ens_n = c->N; p->Dp = c->Dp; p->D = c->D;
    p->order = c->d_order[c->cur];
    p->lo = c->d_lo; p->hi = c->d_hi;
    p->state = c->d_state;
    p->kind = c->primary;
    p->maxmcmc = c->cfg.maxmcmc;
    p->chain_base = c->chain_counter;
    p->seed_lo = (uint32_t)c->cfg.seed; p->seed_hi = (uint32_t)(c->cfg.seed >> 32);
    p->blob = c->d_blob;
    const int nout = c->n_peers > 0 ? c->n_peers : 1;
    p->n_out = nout;
    for (int o = 0; o < nout; ++o) {
        // slot 0 is always the local block, the other ranks follow
        const int r = (o == 0) ? c->rank : (o <= c->rank ? o - 1 : o);
        char* b = (c->n_peers > 0) ? (char*)c->stage_peer[r] + (size_t)c->stage_half * stage_bytes(c->N, c->Dp)
                                   : (char*)c->d_stage_own;
        p->out_x[o] = (double*)(b + stage_off(c->N, c->Dp, 0));
        p->out_logL[o] = (double*)(b + stage_off(c->N, c->Dp, 1));
        p->out_logP[o] = (double*)(b + stage_off(c->N, c->Dp, 2));
        p->out_steps[o] = (int*)(b + stage_off(c->N, c->Dp, 3));
        p->out_acc[o] = (int*)(b + stage_off(c->N, c->Dp, 4));
        p->out_start[o] = (int*)(b + stage_off(c->N, c->Dp, 5));
        p->out_evals[o] = (int*)(b + stage_off(c->N, c->Dp, 6));
        p->out_ne[o] = (int*)(b + stage_off(c->N, c->Dp, 7));
        p->out_nc[o] = (int*)(b + stage_off(c->N, c->Dp, 8));
        p->out_mid_x[o] = (double*)(b + stage_off(c->N, c->Dp, 9));
        p->out_mid_logL[o] = (double*)(b + stage_off(c->N, c->Dp, 10));
    }
}

// the proposal cycle as it travels in the launch parameters of the production MH kernel (cycles of up to kCycInline slots):
// the kinds themselves and, for every cycle position, the kinds of the kKindWin steps that start there packed into a word
static void fill_inline_cycle(MHParams* p, const uint8_t* cyc, int len) {
    if (len > kCycInline) return;                       // longer cycles are read from device memory (chain-history instantiation)
    memcpy(p->cyc, cyc, len);
    for (int q = 0; q < len; ++q) {
        uint32_t w = 0;
        for (int j = 0; j < kKindWin; ++j) w |= (uint32_t)(cyc[(q + j) % len] & 3) << (2 * j);
        p->cycw[q] = make_uint2(w, (uint32_t)((q + kKindWin) % len));
    }
}

static void fill_mh(cpn_ctx* c, MHParams* p) {
    memset(p, 0, sizeof *p);
    fill_common(c, p);
    p->cycle = c->d_cycle; p->cycle_len = c->cycle_len;
    fill_inline_cycle(p, c->h_cycle, c->cycle_len);
    p->cycle_pos0 = (int)((c->batches_enqueued * 37) % c->cycle_len);
    p->eig_scale = c->d_eigscale[c->eig_cur]; p->eig_vec = c->d_eigvec[c->eig_cur];
    p->eig_row0 = c->N + c->eig_cur * c->D;
    p->chain_log = c->d_chain_log; p->chain_log_count = c->d_chain_log_count;
    p->chain_log_n = c->chain_log_n; p->chain_log_cap = c->chain_log_cap;
}

static void fill_slice(cpn_ctx* c, SliceParams* p) {
    memset(p, 0, sizeof *p);
    fill_common(c, p);
    p->cycle = c->d_slice_cycle; p->cycle_len = c->slice_cycle_len;
    p->cycle_pos0 = (int)((c->batches_enqueued * 37) % c->slice_cycle_len);
    p->mean = c->d_eigmean[c->eig_cur]; p->eig_scale = c->d_eigscale[c->eig_cur]; p->eig_vec = c->d_eigvec[c->eig_cur];
}

static void fill_hmc(cpn_ctx* c, HMCParams* p) {
    memset(p, 0, sizeof *p);
    fill_common(c, p);
    p->cov = c->d_covbuf[c->eig_cur]; p->eig_scale = c->d_eigscale[c->eig_cur]; p->eig_vec = c->d_eigvec[c->eig_cur];
    p->base_L = c->hmc_base_L;
    p->max_traj = c->cfg.maxmcmc;
}

// chains [j0, j1) of a batch with sampler `kind`; START_SURVIVOR unless start rows are given (START_GIVEN: row c of the
// start arrays belongs to chain start_row0 + c of the launch)
static int launch_chains_kind(cpn_ctx* c, int kind, int j0, int j1, int start_mode, int n_skip, bool use_override,
                              double logLmin, int nmcmc, int n_out_override, int start_row0) {
    int G, E;
    if (j1 <= j0) return 0;
    if (pick_config(c->D, c->cfg.lanes, c->cfg.functor, j1 - j0, &G, &E)) return 1;
    const double* sx = c->d_start_x + (size_t)start_row0 * c->Dp;
    const double* sl = c->d_start_logL + start_row0;
    const double* sp = c->d_start_logP + start_row0;
    if (kind == CPN_SAMPLER_SLICE) {
        SliceParams p;
        fill_slice(c, &p);
        p.kind = kind;
        p.start_mode = start_mode; p.n_skip = n_skip; p.j0 = j0; p.j1 = j1;
        p.start_x = sx; p.start_logL = sl; p.start_logP = sp;
        if (use_override) {
            if (fetch_state(c)) return 1;
            p.use_override = 1; p.logLmin_override = logLmin; p.nmcmc_override = nmcmc; p.mu_override = c->h_state->slice_mu;
        }
        if (n_out_override > 0) p.n_out = n_out_override;
        if (run_slice(c, G, E, 0, p)) return 1;
    } else if (kind == CPN_SAMPLER_MH) {
        MHParams p;
        fill_mh(c, &p);
        p.kind = kind;
        p.padded = c->Dp >= G * E;
        p.start_mode = start_mode; p.n_skip = n_skip; p.j0 = j0; p.j1 = j1;
        p.start_x = sx; p.start_logL = sl; p.start_logP = sp;
        if (use_override) { p.use_override = 1; p.logLmin_override = logLmin; p.nmcmc_override = nmcmc; }
        if (n_out_override > 0) p.n_out = n_out_override;
        if (run_mh(c, G, E, 0, p)) return 1;
    } else {
        HMCParams p;
        fill_hmc(c, &p);
        p.kind = kind;
        p.start_mode = start_mode; p.n_skip = n_skip; p.j0 = j0; p.j1 = j1;
        p.start_x = sx; p.start_logL = sl; p.start_logP = sp;
        if (use_override) {
            if (fetch_state(c)) return 1;
            p.use_override = 1; p.logLmin_override = logLmin; p.nmcmc_override = nmcmc; p.dt_override = c->h_state->hmc_dt;
        }
        if (n_out_override > 0) p.n_out = n_out_override;
        if (run_hmc(c, G, E, 0, p)) return 1;
    }
    CKL();
    c->launches += 1;
    return 0;
}

// chains [j0, j1) of the K chains of a batch, each with the sampler kind that owns it (kind_ranges).  A mixed
// population forks onto one stream per kind so that the kinds' kernels share the GPU, and joins the main stream.
static int launch_chains(cpn_ctx* c, int K, int j0, int j1, int start_mode, int n_skip, bool use_override, double logLmin,
                         int nmcmc, int n_out_override) {
    int kb[kNumKinds + 1];
    if (kind_ranges(c, K, kb)) return 1;
    for (int k = 0; k <= kNumKinds; ++k) c->last_kb[k] = kb[k];
    if (!c->mixed)
        return launch_chains_kind(c, c->primary, j0, j1, start_mode, n_skip, use_override, logLmin, nmcmc, n_out_override, 0);
    CK(cudaEventRecord(c->ev_fork, c->stream));
    int rc = 0;
    for (int k = 0; k < kNumKinds && !rc; ++k) {
        const int a = std::max(j0, kb[k]), b = std::min(j1, kb[k + 1]);
        if (b <= a) continue;
        CK(cudaStreamWaitEvent(c->kind_stream[k], c->ev_fork, 0));
        c->chain_stream = c->kind_stream[k];
        rc = launch_chains_kind(c, k, a, b, start_mode, n_skip, use_override, logLmin, nmcmc, n_out_override, a - j0);
        c->chain_stream = nullptr;
        if (rc) break;
        CK(cudaEventRecord(c->ev_kind[k], c->kind_stream[k]));
        CK(cudaStreamWaitEvent(c->stream, c->ev_kind[k], 0));
    }
    return rc;
}

// start/end correlation + batch counters of the chains [j0, j1) of kind `kind` (ns_kernels.cuh), on stream `cs`
static int launch_chain_corr(cpn_ctx* c, cudaStream_t cs, int kind, int j0, int j1, double* rho, double* rho_mid = nullptr) {
    const int nseg = std::max(1, (j1 - j0 + kCorrSeg - 1) / kCorrSeg);
    const size_t need = (size_t)(c->D + 2) * nseg * kCorrMom;
    if (need > c->corr_part_cap) return fail("chain statistics: %zu partial sums exceed the scratch of %zu", need, c->corr_part_cap);
    const bool mid = rho_mid != nullptr && kind == CPN_SAMPLER_MH;
    k_chain_corr_partial<<<dim3((c->D + 1 + 31) / 32 + 1, nseg), kCorrThreads, 0, cs>>>(c->d_state, kind, j0, j1, c->D, c->Dp, c->d_new_start, c->d_x, c->d_logL,
                                                                      c->d_new_x, c->d_new_logL, mid ? c->d_mid_x : nullptr,
                                                                      mid ? c->d_mid_logL : nullptr, c->d_new_steps, c->d_new_acc,
                                                                      c->d_new_evals, c->d_corr_part);
    k_chain_corr_final<<<c->D + 2, 32, 0, cs>>>(c->d_state, kind, j1 - j0, c->D, nseg, c->d_corr_part, rho, rho_mid);
    c->launches += 2;
    return 0;
}

extern "C" int cpn_init_prior(cpn_ctx* c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    init_state(c, c->h_state);
    if (push_state(c)) return 1;
    c->dead_upper = 0; c->batches_enqueued = 0; c->chain_counter = 0; c->cur = 0; c->eig_calls = 0;
    EvalParams e;
    memset(&e, 0, sizeof e);
    e.x = c->d_x; e.logL = c->d_logL; e.logP = c->d_logP; e.n = c->N; e.D = c->D; e.Dp = c->Dp;
    e.lo = c->d_lo; e.hi = c->d_hi; e.blob = c->d_blob; e.draw = 1;
    e.seed_lo = (uint32_t)c->cfg.seed; e.seed_hi = (uint32_t)(c->cfg.seed >> 32);
    e.id_base = 0x4000000000000000ull;
    if (run_eval(c, c->G, c->E, e)) return 1;
    CKL();
    // Sampler.reset (sampler.py:129-135): evolve every member with logLmin = -inf so the set follows
    // the prior, adapting the chain length on the way.  Passes are repeated until the product of the
    // per-pass start/end correlations shows that the members have forgotten their uniform draw
    // (at least 2, at most 24 passes).
    // Functors with the flat prior of Model.log_prior (model.py:66-82) need no burn-in: the uniform
    // draw IS a draw from the prior.
    const bool flat = (c->cfg.functor == CPN_USER) ? c->user_flat_prior
                                                   : !(c->cfg.functor == CPN_GAUSS_MEAN_SIGMA || c->cfg.functor == CPN_GAUSS_MIXTURE);
    double memory = 1.0;
    for (int pass = 0; pass < (flat ? 0 : 24); ++pass) {
        if (cpn_update_ensemble_stats(c)) return 1;
        MHParams p;
        fill_mh(c, &p);
        p.start_mode = START_SELF;
        p.chain_log = nullptr;                      // the burn-in is not part of the run's chain history
        p.n_skip = 0;
        p.use_override = 0;
        p.j0 = 0; p.j1 = c->N;
        p.chain_base = 0x2000000000000000ull + (unsigned long long)pass * (unsigned long long)c->N;
        p.cycle_pos0 = pass * 13 % c->cycle_len;
        int G, E;
        if (pick_config(c->D, c->cfg.lanes, c->cfg.functor, c->N, &G, &E)) return 1;
        if (run_mh(c, G, E, 0, p)) return 1;
        CKL();
        // the burn-in chains (always MH) run under the primary kind's controller
        if (launch_chain_corr(c, c->stream, c->primary, 0, c->N, c->d_rho)) return 1;
        k_adapt_nmcmc<<<1, 1, 0, c->stream>>>(c->d_state, c->primary, c->N, c->d_rho, c->d_rho2, c->D);   // coordinates only: logL is unconstrained here
        CK(cudaMemcpyAsync(c->d_x, c->d_new_x, (size_t)c->N * c->Dp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_logL, c->d_new_logL, c->N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_logP, c->d_new_logP, c->N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        c->launches += 1;
        std::vector<double> rho(c->D);
        CK(cudaMemcpyAsync(rho.data(), c->d_rho, c->D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        double r = 0.0;
        for (double v : rho)
            if (v == v) r = std::max(r, std::fabs(v));
        memory *= std::min(1.0, r + 2.0 / std::sqrt((double)c->N));
        if (pass >= 1 && memory < 0.05) break;
    }
    // the burn-in statistics must not leak into the run's controller
    if (fetch_state(c)) return 1;
    for (int k = 0; k < kNumKinds; ++k) {
        c->h_state->ctl[k].rho_max = -1.0;
        c->h_state->ctl[k].nmcmc = c->h_state->ctl[c->primary].nmcmc;             // every kind starts from the burn-in's length
        c->h_state->ctl[k].nmcmc_exact = c->h_state->ctl[c->primary].nmcmc_exact;
    }
    if (push_state(c)) return 1;
    if (sort_live(c)) return 1;
    if (cpn_update_ensemble_stats(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    c->live_ready = true;
    return 0;
}

extern "C" int cpn_set_live(cpn_ctx* c, const double* rows, int32_t n) {
    if (!c || !rows) return fail("null argument");
    if (n != c->N) return fail("cpn_set_live: got %d rows, nlive is %d", n, c->N);
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp, N = c->N;
    std::vector<double> x((size_t)N * Dp, 0.0), l(N), lp(N);
    for (int i = 0; i < N; ++i) {
        for (int d = 0; d < D; ++d) x[(size_t)i * Dp + d] = rows[(size_t)i * (D + 2) + d];
        l[i] = rows[(size_t)i * (D + 2) + D];
        lp[i] = rows[(size_t)i * (D + 2) + D + 1];
    }
    CK(cudaStreamSynchronize(c->stream));
    CK(h2d_sync(c->d_x, x.data(), x.size() * sizeof(double)));
    CK(h2d_sync(c->d_logL, l.data(), N * sizeof(double)));
    CK(h2d_sync(c->d_logP, lp.data(), N * sizeof(double)));
    if (sort_live(c)) return 1;
    if (cpn_update_ensemble_stats(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    c->live_ready = true;
    return 0;
}

static int check_K(cpn_ctx* c, int K) {
    if (!c) return fail("null context");
    if (!c->live_ready) return fail("live set not initialised: call cpn_init_prior or cpn_set_live first");
    if (K < 1 || K > c->N - 3) return fail("K=%d must be in [1, nlive-3]", K);
    return 0;
}

extern "C" int cpn_select_and_accumulate(cpn_ctx* c, int32_t K) {
    if (check_K(c, K)) return 1;
    CK(cudaSetDevice(c->cfg.device));
    // Refresh cadence of the reference: every poolsize//4 chains (sampler.py:252-254).  The refresh runs
    // on the side stream beside the chain kernels of three batches and is adopted by the batch after those
    // (fixed latency, so the schedule - and with it every result - is deterministic); one refresh is in
    // flight at a time.
    if (c->eig_pending && ++c->eig_pending_age >= 3)
        if (adopt_pending_stats(c)) return 1;
    // Beyond nlive/4 chains the cadence follows what the statistics are used for: replacing K of N points shrinks
    // each axis of the ensemble by exp(-K/(N D)), so the jump scales sqrt|lambda_i| drift by 5% only after
    // 0.05 N D chains (5 batches at N = 1e4, D = 50, K = 4.7e3; nlive/4 again for D <= 5).
    const long long cadence = std::max<long long>(std::max(1, c->N / 4), (long long)(0.05 * (double)c->N * (double)c->D));
    if (!c->eig_pending && c->chains_since_stats >= cadence)
        if (start_async_stats(c)) return 1;

    if (grow_logs(c, c->dead_upper + K + c->N)) return 1;
    AccParams a;
    memset(&a, 0, sizeof a);
    a.st = c->d_state; a.skeys = c->d_skeys[c->cur]; a.N = c->N; a.K = K; a.mode = c->cfg.ns_mode;
    a.final_sweep = 0; a.nlive_run = c->N; a.hist_logL = c->d_hist_logL; a.hist_logvol = c->d_hist_logvol;
    a.update_threshold = 1;
    if (join_acc(c)) return 1;
    k_ns_threshold<<<1, 1, 0, c->stream>>>(c->d_state, c->d_skeys[c->cur], c->N, K);
    CK(cudaEventRecord(c->ev_thr, c->stream));
    CK(cudaStreamWaitEvent(c->acc_stream, c->ev_thr, 0));
    k_ns_accumulate<<<1, kAccThreads, 0, c->acc_stream>>>(a);
    CKL();
    CK(cudaEventRecord(c->ev_acc, c->acc_stream));
    c->acc_pending = true;
    c->dead_upper += K;
    c->launches += 2;
    return 0;
}

extern "C" int cpn_evolve_batch(cpn_ctx* c, int32_t K) {
    if (check_K(c, K)) return 1;
    CK(cudaSetDevice(c->cfg.device));
    if (c->n_peers > 0) {                       // next half of the peer-mapped staging blocks
        c->stage_half ^= 1;
        bind_staging(c, (char*)c->stage_peer[c->rank] + (size_t)c->stage_half * stage_bytes(c->N, c->Dp));
    }
    // this rank's slice of the K chains of the batch (all of them on a single GPU)
    const int j0 = (int)((long long)K * c->rank / c->world), j1 = (int)((long long)K * (c->rank + 1) / c->world);
    if (launch_chains(c, K, j0, j1, START_SURVIVOR, K, false, 0.0, 0, 0)) return 1;
    c->chain_counter += (unsigned long long)K;
    c->chains_since_stats += K;
    c->batches_enqueued += 1;
    return 0;
}

// NCCL path of a multi-GPU run: this rank's slice of the staging arrays as one contiguous record
// block [chains][Dp + 5] doubles {x.., logL, logP, (steps, acc), (start, evals), (ne, nc)}, and back.
__global__ void k_pack_shard(int j0, int n, int Dp, const double* x, const double* logL, const double* logP, const int* steps,
                             const int* acc, const int* start, const int* evals, const int* ne, const int* nc, double* rec) {
    const int S = Dp + 5;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * S) return;
    const int r = i / S, d = i % S, j = j0 + r;
    double v;
    if (d < Dp) v = x[(size_t)j * Dp + d];
    else if (d == Dp) v = logL[j];
    else if (d == Dp + 1) v = logP[j];
    else if (d == Dp + 2) v = __hiloint2double(acc[j], steps[j]);
    else if (d == Dp + 3) v = __hiloint2double(evals[j], start[j]);
    else v = __hiloint2double(nc[j], ne[j]);
    rec[i] = v;
}
__global__ void k_unpack_shards(int K, int Dp, const double* rec, double* x, double* logL, double* logP, int* steps, int* acc,
                                int* start, int* evals, int* ne, int* nc) {
    const int S = Dp + 5;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * S) return;
    const int j = i / S, d = i % S;
    const double v = rec[i];
    if (d < Dp) x[(size_t)j * Dp + d] = v;
    else if (d == Dp) logL[j] = v;
    else if (d == Dp + 1) logP[j] = v;
    else if (d == Dp + 2) { steps[j] = __double2loint(v); acc[j] = __double2hiint(v); }
    else if (d == Dp + 3) { start[j] = __double2loint(v); evals[j] = __double2hiint(v); }
    else { ne[j] = __double2loint(v); nc[j] = __double2hiint(v); }
}

extern "C" int64_t cpn_shard_record_doubles(cpn_ctx* c) { return c ? c->Dp + 5 : -1; }

extern "C" int cpn_shard_range(cpn_ctx* c, int32_t K, int32_t rank, int32_t* j0, int32_t* j1) {
    if (!c || !j0 || !j1 || rank < 0 || rank >= c->world) return fail("bad argument");
    *j0 = (int)((long long)K * rank / c->world);
    *j1 = (int)((long long)K * (rank + 1) / c->world);
    return 0;
}

extern "C" int cpn_pack_shard(cpn_ctx* c, int32_t K, void* dev_send) {
    if (check_K(c, K) || !dev_send) return fail("bad argument");
    CK(cudaSetDevice(c->cfg.device));
    const int j0 = (int)((long long)K * c->rank / c->world), j1 = (int)((long long)K * (c->rank + 1) / c->world);
    const int n = j1 - j0, S = c->Dp + 5, T = 256;
    if (n > 0)
        k_pack_shard<<<(n * S + T - 1) / T, T, 0, c->stream>>>(j0, n, c->Dp, c->d_new_x, c->d_new_logL, c->d_new_logP, c->d_new_steps,
                                                            c->d_new_acc, c->d_new_start, c->d_new_evals, c->d_new_ne, c->d_new_nc,
                                                            (double*)dev_send);
    CKL();
    c->launches += 1;
    return 0;
}

extern "C" int cpn_unpack_shards(cpn_ctx* c, int32_t K, const void* dev_recv) {
    if (check_K(c, K) || !dev_recv) return fail("bad argument");
    CK(cudaSetDevice(c->cfg.device));
    const int S = c->Dp + 5, T = 256;
    k_unpack_shards<<<(K * S + T - 1) / T, T, 0, c->stream>>>(K, c->Dp, (const double*)dev_recv, c->d_new_x, c->d_new_logL,
                                                           c->d_new_logP, c->d_new_steps, c->d_new_acc, c->d_new_start, c->d_new_evals,
                                                           c->d_new_ne, c->d_new_nc);
    CKL();
    c->launches += 1;
    return 0;
}

extern "C" int cpn_insert_batch(cpn_ctx* c, int32_t K) {
    if (check_K(c, K)) return 1;
    CK(cudaSetDevice(c->cfg.device));
    const int T = 256;
    if (join_acc(c)) return 1;                 // the merge appends at the counters the accumulation has advanced
    int kb[kNumKinds + 1];
    if (kind_ranges(c, K, kb)) return 1;
    const bool corr = !c->staged_external;
    // the mid-chain snapshots travel with the fused exchange (and trivially on one GPU), not with the NCCL record
    const bool twolag_ok = (c->world == 1 || c->n_peers > 0) && !(c->cfg.flags & CPN_NO_TWO_LAG);
    MergeParams m;
    memset(&m, 0, sizeof m);
    m.st = c->d_state; m.st_w = c->d_state; m.N = c->N; m.K = K; m.Dp = c->Dp; m.Dreal = c->D;
    m.dshift = 0;
    while ((1 << m.dshift) < c->Dp) ++m.dshift;
    m.rho = c->staged_external ? nullptr : c->d_rho;
    m.rho2_ema = c->d_rho2;
    m.rho_mid = (m.rho != nullptr && twolag_ok) ? c->d_rho_mid : nullptr;
    m.lag_ema = c->d_lag_ema;
    m.allow_twolag = twolag_ok ? 1 : 0;
    for (int k = 0; k < kNumKinds; ++k) m.kind_n[k] = kb[k + 1] - kb[k];
    {
        // dependence between chain start and chain end across the (whole, gathered) batch for the chain-length
        // control, the batch counters and the controllers themselves: on the accumulation stream (idle by now), beside
        // the ranking of the replacements below; the merge waits for both
        cudaStream_t cs = c->acc_stream;
        CK(cudaEventRecord(c->ev_ins, c->stream));
        CK(cudaStreamWaitEvent(cs, c->ev_ins, 0));
        for (int k = 0; corr && k < kNumKinds; ++k) {
            if (kb[k + 1] <= kb[k]) continue;
            // (the kinds' statistics share one scratch: they run one after the other on this stream)
            if (launch_chain_corr(c, cs, k, kb[k], kb[k + 1], c->d_rho + k * (c->D + 1), twolag_ok ? c->d_rho_mid + k * (c->D + 1) : nullptr))
                return 1;
            if (k == CPN_SAMPLER_SLICE) {
                k_adapt_slice_mu<<<1, 256, 0, cs>>>(c->d_state, kb[k], kb[k + 1], c->d_new_ne, c->d_new_nc);
                c->launches += 1;
            } else if (k == CPN_SAMPLER_HMC) {
                k_adapt_hmc_dt<<<1, 1, 0, cs>>>(c->d_state);
                c->launches += 1;
            }
        }
        k_adapt_batch<<<1, 32, 0, cs>>>(m);
        c->launches += 1;
        CK(cudaEventRecord(c->ev_corr, cs));
    }
    // rank of every replacement among the replacements: buckets between the sorted survivors (ns_kernels.cuh)
    {
        const int ns = c->N - K, nb = ns + 1;
        int* hist = c->d_bucket;                    // [nb] histogram -> bucket starts
        int* cursor = c->d_bucket + c->N + 1;       // [nb]
        int* members = c->d_bucket + 2 * (c->N + 1);   // [K]
        CK(cudaMemsetAsync(c->d_bucket, 0, 2 * (size_t)(c->N + 1) * sizeof(int), c->stream));
        const int gb = (K + T - 1) / T;
        k_rank_bucket<<<gb, T, 0, c->stream>>>(c->d_state, c->d_new_logL, K, c->d_skeys[c->cur] + K, ns, c->d_nle, hist);
        const int ntile = (nb + kScanTile - 1) / kScanTile;
        int* tile_sum = c->d_bucket + 3 * (c->N + 1);
        k_rank_scan_tiles<<<ntile, kScanThreads, 0, c->stream>>>(c->d_state, hist, nb, tile_sum);
        k_rank_scan_apply<<<ntile, kScanThreads, 0, c->stream>>>(c->d_state, hist, nb, tile_sum);
        k_rank_fill<<<gb, T, 0, c->stream>>>(c->d_state, K, c->d_nle, hist, cursor, members);
        k_rank_final<<<gb, T, 0, c->stream>>>(c->d_state, c->d_new_logL, K, c->d_nle, hist, cursor, members, c->d_rank, c->d_snew);
        c->launches += 5;
    }
    c->staged_external = false;
    m.order_in = c->d_order[c->cur]; m.skeys_in = c->d_skeys[c->cur];
    m.order_out = c->d_order[c->cur ^ 1]; m.skeys_out = c->d_skeys[c->cur ^ 1];
    m.rank_new = c->d_rank; m.snew = c->d_snew; m.nle = c->d_nle;
    m.live_x = c->d_x; m.live_logL = c->d_logL; m.live_logP = c->d_logP;
    m.new_x = c->d_new_x; m.new_logL = c->d_new_logL; m.new_logP = c->d_new_logP;
    m.dead_x = c->d_dead_x; m.dead_logL = c->d_dead_logL; m.dead_logP = c->d_dead_logP;
    m.insertion = c->d_ins;
    CK(cudaStreamWaitEvent(c->stream, c->ev_corr, 0));
    if (c->wait_cov) {     // an asynchronous refresh must have read the live rows before they are overwritten
        CK(cudaStreamWaitEvent(c->stream, c->ev_cov, 0));
        c->wait_cov = false;
    }
    k_merge_commit<<<(int)(((long long)c->N + ((long long)K << m.dshift) + T - 1) / T), T, 0, c->stream>>>(m);
    CKL();
    c->cur ^= 1;
    c->launches += 1;
    return 0;
}

extern "C" int cpn_ns_step(cpn_ctx* c, int32_t K) {
    if (check_K(c, K)) return 1;
    if (cpn_select_and_accumulate(c, K)) return 1;
    if (cpn_evolve_batch(c, K)) return 1;
    return cpn_insert_batch(c, K);
}

extern "C" int cpn_run(cpn_ctx* c, int32_t K, int64_t max_batches, double tolerance, int64_t* batches_done) {
    if (check_K(c, K)) return 1;
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    const long long b0 = c->h_state->batches;
    c->h_state->tolerance = tolerance;
    if (push_state(c)) return 1;
    const int chunk = 8;
    long long enq = 0;
    while (max_batches <= 0 || enq < max_batches) {
        const int n = (max_batches > 0) ? (int)std::min<long long>(chunk, max_batches - enq) : chunk;
        for (int i = 0; i < n; ++i)
            if (cpn_ns_step(c, K)) return 1;
        enq += n;
        if (fetch_state(c)) return 1;
        if (c->h_state->done || c->h_state->stop_next) break;
    }
    if (batches_done) *batches_done = c->h_state->batches - b0;
    return 0;
}

extern "C" int cpn_finalise(cpn_ctx* c, double* logZ_rect, double* logZ_trap) {
    if (!c) return fail("null context");
    if (!c->live_ready) return fail("live set not initialised");
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    const long long dead0 = c->h_state->iteration;
    if (grow_logs(c, dead0 + c->N + 1)) return 1;
    AccParams a;
    memset(&a, 0, sizeof a);
    a.st = c->d_state; a.skeys = c->d_skeys[c->cur]; a.N = c->N; a.K = c->N; a.mode = CPN_NS_SEQUENTIAL;
    a.final_sweep = 1; a.nlive_run = c->N; a.hist_logL = c->d_hist_logL; a.hist_logvol = c->d_hist_logvol;
    a.update_threshold = 1;
    k_ns_accumulate<<<1, kAccThreads, 0, c->stream>>>(a);
    const int T = 256;
    k_append_live_sorted<<<(c->N + T - 1) / T, T, 0, c->stream>>>(c->d_order[c->cur], c->N, c->Dp, c->d_x, c->d_logL,
                                                                 c->d_logP, c->d_dead_x, c->d_dead_logL, c->d_dead_logP, dead0);
    CKL();
    if (fetch_state(c)) return 1;
    if (logZ_rect) *logZ_rect = c->h_state->logZ;
    const long long nh = c->h_state->n_hist;
    k_trap_partial<<<256, 256, 0, c->stream>>>(c->d_hist_logL, c->d_hist_logvol, nh, c->d_trap_partial);
    k_trap_final<<<1, 32, 0, c->stream>>>(c->d_trap_partial, 256, c->d_scalar);
    CKL();
    double z;
    CK(cudaMemcpyAsync(&z, c->d_scalar, sizeof z, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (logZ_trap) *logZ_trap = z;
    c->h_state->logZ = z;                       // self.logZ = self.state.finalise() (NestedSampling.py:479-480)
    c->h_state->done = 1;
    if (push_state(c)) return 1;
    c->live_ready = false;                      // the live set has been consumed
    return 0;
}

// ---- checkpoint / resume (NestedSampler.checkpoint / resume + Sampler.checkpoint / resume,
// cpnest/NestedSampling.py:495-537, cpnest/sampler.py:275-314): the device state is a handful of flat arrays and
// counters; it is written into / read from one caller-owned host blob.  A resumed run continues bit-identically.
namespace {
struct CkptHeader {
    char magic[8];                 // "CPNB200\0"
    int32_t version, dim, nlive, functor, sampler, ns_mode, maxmcmc, cycle_len, slice_cycle_len, cur, eig_cur, eig_pending,
        eig_pending_age, live_ready, mix[kNumKinds];
    uint64_t seed;
    long long dead_upper, batches_enqueued, chains_since_stats, eig_calls, n_dead, n_hist, n_ins;
    unsigned long long chain_counter;
    NSState state;
};
struct CkptSection { void* dev; size_t bytes; };

static std::vector<CkptSection> ckpt_sections(cpn_ctx* c, long long n_dead, long long n_hist, long long n_ins) {
    const size_t N = c->N, D = c->D, Dp = c->Dp;
    std::vector<CkptSection> v = {
        {c->d_x, N * Dp * 8}, {c->d_logL, N * 8}, {c->d_logP, N * 8},
        {c->d_order[c->cur], N * 4}, {c->d_skeys[c->cur], N * 8},
        {c->d_dead_x, (size_t)n_dead * Dp * 8}, {c->d_dead_logL, (size_t)n_dead * 8}, {c->d_dead_logP, (size_t)n_dead * 8},
        {c->d_hist_logL, (size_t)n_hist * 8}, {c->d_hist_logvol, (size_t)n_hist * 8}, {c->d_ins, (size_t)n_ins * 8},
        {c->d_rho2, kNumKinds * (D + 1) * 8}, {c->d_lag_ema, 2 * kNumKinds * (D + 1) * 8}, {c->d_cycle, (size_t)c->cycle_len}, {c->d_slice_cycle, (size_t)c->slice_cycle_len},
        {c->d_mean, D * 8}, {c->d_cov, D * D * 8}};
    for (int b = 0; b < 2; ++b) {
        v.push_back({c->d_eigval[b], D * 8});
        v.push_back({c->d_eigvec[b], D * Dp * 8});
        v.push_back({c->d_eigscale[b], D * 8});
        v.push_back({c->d_eigmean[b], D * 8});
        v.push_back({c->d_covbuf[b], D * D * 8});
    }
    return v;
}
}  // namespace

// finish what is in flight (a statistics refresh on the side stream) so that the arrays are final
static int ckpt_quiesce(cpn_ctx* c) {
    CK(cudaSetDevice(c->cfg.device));
    if (join_acc(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->side));
    c->wait_cov = false;
    return fetch_state(c);
}

extern "C" int64_t cpn_checkpoint_bytes(cpn_ctx* c) {
    if (!c) return -1;
    if (ckpt_quiesce(c)) return -1;
    const NSState* s = c->h_state;
    size_t tot = sizeof(CkptHeader);
    for (auto& sec : ckpt_sections(c, s->iteration, s->n_hist, s->n_ins)) tot += (sec.bytes + 7) & ~(size_t)7;
    return (int64_t)tot;
}

extern "C" int cpn_save_checkpoint(cpn_ctx* c, void* blob, int64_t bytes) {
    if (!c || !blob) return fail("null argument");
    if (ckpt_quiesce(c)) return 1;
    const NSState* s = c->h_state;
    CkptHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "CPNB200", 8);
    h.version = 3; h.dim = c->D; h.nlive = c->N; h.functor = c->cfg.functor; h.sampler = c->cfg.sampler; h.ns_mode = c->cfg.ns_mode;
    h.maxmcmc = c->cfg.maxmcmc; h.cycle_len = c->cycle_len; h.slice_cycle_len = c->slice_cycle_len; h.cur = c->cur;
    h.eig_cur = c->eig_cur; h.eig_pending = c->eig_pending ? 1 : 0; h.eig_pending_age = c->eig_pending_age;
    for (int k = 0; k < kNumKinds; ++k) h.mix[k] = c->mix[k];
    h.live_ready = c->live_ready ? 1 : 0; h.seed = c->cfg.seed; h.dead_upper = c->dead_upper; h.batches_enqueued = c->batches_enqueued;
    h.chains_since_stats = c->chains_since_stats; h.eig_calls = c->eig_calls; h.n_dead = s->iteration; h.n_hist = s->n_hist;
    h.n_ins = s->n_ins; h.chain_counter = c->chain_counter; h.state = *s;
    auto secs = ckpt_sections(c, h.n_dead, h.n_hist, h.n_ins);
    size_t tot = sizeof h;
    for (auto& sec : secs) tot += (sec.bytes + 7) & ~(size_t)7;
    if ((size_t)bytes < tot) return fail("cpn_save_checkpoint: blob of %lld bytes, %zu needed", (long long)bytes, tot);
    char* p = (char*)blob;
    memcpy(p, &h, sizeof h);
    p += sizeof h;
    for (auto& sec : secs) {
        if (sec.bytes) CK(cudaMemcpy(p, sec.dev, sec.bytes, cudaMemcpyDeviceToHost));
        p += (sec.bytes + 7) & ~(size_t)7;
    }
    return 0;
}

extern "C" int cpn_load_checkpoint(cpn_ctx* c, const void* blob, int64_t bytes) {
    if (!c || !blob) return fail("null argument");
    if ((size_t)bytes < sizeof(CkptHeader)) return fail("cpn_load_checkpoint: blob too small");
    CkptHeader h;
    memcpy(&h, blob, sizeof h);
    if (memcmp(h.magic, "CPNB200", 8) != 0 || h.version != 3) return fail("cpn_load_checkpoint: not a cpnest_b200 checkpoint (version 3)");
    if (h.dim != c->D || h.nlive != c->N || h.functor != c->cfg.functor || h.sampler != c->cfg.sampler || h.ns_mode != c->cfg.ns_mode ||
        h.maxmcmc != c->cfg.maxmcmc || h.mix[0] != c->mix[0] || h.mix[1] != c->mix[1] || h.mix[2] != c->mix[2])
        return fail("cpn_load_checkpoint: checkpoint of a different run (dim %d nlive %d functor %d sampler %d maxmcmc %d)", h.dim,
                    h.nlive, h.functor, h.sampler, h.maxmcmc);
    // every field that sizes a copy or indexes a buffer is checked before the context is touched: a corrupt or truncated
    // file leaves the context as it was
    if (h.cycle_len < 1 || h.cycle_len > 4096 || h.slice_cycle_len < 1 || h.slice_cycle_len > 4096)
        return fail("cpn_load_checkpoint: corrupt header (cycle lengths %d, %d)", h.cycle_len, h.slice_cycle_len);
    if ((h.cur != 0 && h.cur != 1) || (h.eig_cur != 0 && h.eig_cur != 1))
        return fail("cpn_load_checkpoint: corrupt header (buffer indices %d, %d)", h.cur, h.eig_cur);
    if (h.n_dead < 0 || h.n_hist < 0 || h.n_ins < 0 || h.dead_upper < h.n_dead || h.batches_enqueued < 0 || h.chains_since_stats < 0 ||
        h.eig_calls < 0 || h.eig_pending_age < 0)
        return fail("cpn_load_checkpoint: corrupt header (negative or inconsistent counters)");
    if (h.n_dead != h.state.iteration || h.n_hist != h.state.n_hist || h.n_ins != h.state.n_ins)
        return fail("cpn_load_checkpoint: corrupt header (log lengths %lld/%lld/%lld do not match the saved state %lld/%lld/%lld)",
                    h.n_dead, h.n_hist, h.n_ins, h.state.iteration, h.state.n_hist, h.state.n_ins);
    {
        // size of the sections with the header's counts, computed without touching the context
        const size_t N = c->N, D = c->D, Dp = c->Dp;
        const size_t sizes[] = {N * Dp * 8, N * 8, N * 8, N * 4, N * 8, (size_t)h.n_dead * Dp * 8, (size_t)h.n_dead * 8, (size_t)h.n_dead * 8,
                                (size_t)h.n_hist * 8, (size_t)h.n_hist * 8, (size_t)h.n_ins * 8, kNumKinds * (D + 1) * 8,
                                2 * kNumKinds * (D + 1) * 8, (size_t)h.cycle_len, (size_t)h.slice_cycle_len, D * 8, D * D * 8,
                                2 * (D * 8), 2 * (D * Dp * 8), 2 * (D * 8), 2 * (D * 8), 2 * (D * D * 8)};
        size_t tot = sizeof h;
        for (int i = 0; i < 17; ++i) tot += (sizes[i] + 7) & ~(size_t)7;
        for (int i = 17; i < 22; ++i) tot += 2 * (((sizes[i] / 2) + 7) & ~(size_t)7);
        if ((size_t)bytes < tot) return fail("cpn_load_checkpoint: truncated blob (%lld of %zu bytes)", (long long)bytes, tot);
    }
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->side));
    if (grow_logs(c, std::max(h.n_dead, std::max(h.n_hist, h.n_ins)) + 2LL * c->N)) return 1;
    c->cycle_len = h.cycle_len; c->slice_cycle_len = h.slice_cycle_len; c->cur = h.cur; c->eig_cur = h.eig_cur;
    c->eig_pending = h.eig_pending != 0; c->eig_pending_age = h.eig_pending_age; c->wait_cov = false;
    c->live_ready = h.live_ready != 0; c->cfg.seed = h.seed; c->dead_upper = h.dead_upper; c->batches_enqueued = h.batches_enqueued;
    c->chains_since_stats = h.chains_since_stats; c->eig_calls = h.eig_calls; c->chain_counter = h.chain_counter;
    c->staged_external = false;
    auto secs = ckpt_sections(c, h.n_dead, h.n_hist, h.n_ins);
    size_t tot = sizeof h;
    for (auto& sec : secs) tot += (sec.bytes + 7) & ~(size_t)7;
    if ((size_t)bytes < tot) return fail("cpn_load_checkpoint: truncated blob (%lld of %zu bytes)", (long long)bytes, tot);
    const char* p = (const char*)blob + sizeof h;
    for (auto& sec : secs) {
        if (sec.bytes) CK(h2d_sync(sec.dev, p, sec.bytes));
        p += (sec.bytes + 7) & ~(size_t)7;
    }
    CK(cudaMemcpy(c->h_cycle, c->d_cycle, (size_t)c->cycle_len, cudaMemcpyDeviceToHost));
    *c->h_state = h.state;
    if (push_state(c)) return 1;
    if (c->eig_pending) {
        // the refresh that was in flight had finished before the state was saved: make its event "already happened"
        CK(cudaEventRecord(c->ev_eig, c->stream));
    }
    return 0;
}

extern "C" int cpn_synchronize(cpn_ctx* c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    if (join_acc(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int cpn_set_stream(cpn_ctx* c, void* stream) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    if (join_acc(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return 0;
}

// ---- chain history (Sampler.samples / mcmc_chain_*.dat, cpnest/sampler.py:86, 261-267, 345) ----
extern "C" int cpn_enable_chain_log(cpn_ctx* c, int32_t n_chains, int32_t n_events) {
    if (!c) return fail("null context");
    if (n_chains < 0 || n_chains > c->N || (n_chains > 0 && n_events < 8)) return fail("cpn_enable_chain_log: bad sizes");
    CK(cudaSetDevice(c->cfg.device));
    if (join_acc(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    if (c->d_chain_log) { cudaFree(c->d_chain_log); cudaFree(c->d_chain_log_count); c->d_chain_log = nullptr; c->d_chain_log_count = nullptr; }
    c->chain_log_n = 0; c->chain_log_cap = 0;
    if (n_chains == 0) return 0;
    if (dalloc(&c->d_chain_log, (size_t)n_chains * n_events * (c->Dp + 4)) || dalloc(&c->d_chain_log_count, n_chains)) return 1;
    c->chain_log_n = n_chains; c->chain_log_cap = n_events;
    return 0;
}

extern "C" int cpn_get_chain_log(cpn_ctx* c, int32_t slot, double* events, int32_t max_events, int32_t* n_out) {
    if (!c || !events || !n_out) return fail("null argument");
    if (!c->d_chain_log || slot < 0 || slot >= c->chain_log_n) return fail("cpn_get_chain_log: no such chain log (enable it with cpn_enable_chain_log)");
    CK(cudaSetDevice(c->cfg.device));
    if (join_acc(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long cnt = 0;
    CK(cudaMemcpy(&cnt, c->d_chain_log_count + slot, sizeof cnt, cudaMemcpyDeviceToHost));
    const int cap = c->chain_log_cap, W = c->Dp + 4, D = c->D;
    const int n = (int)std::min<unsigned long long>(cnt, (unsigned long long)cap);
    if (n > max_events) return fail("cpn_get_chain_log: %d events, room for %d", n, max_events);
    std::vector<double> ring((size_t)cap * W);
    CK(cudaMemcpy(ring.data(), c->d_chain_log + (size_t)slot * cap * W, ring.size() * sizeof(double), cudaMemcpyDeviceToHost));
    // oldest event first; rows out: {tag, step, logL, logP, x[D]}
    for (int i = 0; i < n; ++i) {
        const unsigned long long src = (cnt - n + i) % (unsigned long long)cap;
        const double* r = ring.data() + (size_t)src * W;
        double* o = events + (size_t)i * (D + 4);
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
        for (int d = 0; d < D; ++d) o[4 + d] = r[4 + d];
    }
    *n_out = n;
    return 0;
}

extern "C" int64_t cpn_launch_count(cpn_ctx* c) { return c ? c->launches : -1; }

extern "C" int cpn_pick_lanes(cpn_ctx* c, int32_t chains, int32_t* lanes, int32_t* per_lane) {
    if (!c || !lanes || chains < 1) return fail("bad argument");
    int G, E;
    if (pick_config(c->D, c->cfg.lanes, c->cfg.functor, chains, &G, &E)) return 1;
    *lanes = G;
    if (per_lane) *per_lane = E;
    return 0;
}

extern "C" int cpn_get_state(cpn_ctx* c, cpn_state* o) {
    if (!c || !o) return fail("null argument");
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    const NSState* s = c->h_state;
    o->logZ = s->logZ; o->info = s->info; o->logw = s->logw; o->logLmin = s->logLmin; o->logLmax = s->logLmax;
    const KindCtl& pk = s->ctl[c->primary];
    o->condition = s->condition; o->nmcmc_exact = pk.nmcmc_exact; o->iteration = s->iteration; o->n_hist = s->n_hist;
    o->batches = s->batches;
    unsigned long long t[4];
    for (int i = 0; i < 4; ++i) {
        t[i] = s->totals[i];
        for (int k = 0; k < kNumKinds; ++k) t[i] += s->ctl[k].batch[i];
    }
    o->total_steps = (int64_t)t[0]; o->total_accepted = (int64_t)t[1]; o->total_evals = (int64_t)t[2]; o->failed_chains = (int64_t)t[3];
    o->nmcmc = pk.nmcmc; o->done = s->done; o->rho_max = pk.rho_max;
    for (int k = 0; k < kNumKinds; ++k) {
        o->rho_kind[k] = s->ctl[k].rho_max;
        o->nmcmc_kind[k] = s->ctl[k].nmcmc;
        o->chains_kind[k] = c->last_kb[k + 1] - c->last_kb[k];
        o->two_lag_kind[k] = s->ctl[k].twolag;
        o->rho_decay_kind[k] = s->ctl[k].rho_decay;
    }
    o->slice_mu = s->slice_mu; o->hmc_dt = s->hmc_dt;
    return 0;
}

extern "C" int cpn_set_nmcmc(cpn_ctx* c, int32_t nmcmc) {
    if (!c) return fail("null context");
    if (nmcmc < 1 || nmcmc > c->cfg.maxmcmc) return fail("nmcmc=%d must be in [1, maxmcmc=%d]", nmcmc, c->cfg.maxmcmc);
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    for (int k = 0; k < kNumKinds; ++k) {
        c->h_state->ctl[k].nmcmc = nmcmc;
        c->h_state->ctl[k].nmcmc_exact = (double)nmcmc;
    }
    return push_state(c);
}

extern "C" int cpn_set_tolerance(cpn_ctx* c, double tolerance) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    c->h_state->tolerance = tolerance;
    return push_state(c);
}

extern "C" int cpn_get_live(cpn_ctx* c, double* rows) {
    if (!c || !rows) return fail("null argument");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    const int N = c->N, D = c->D, Dp = c->Dp;
    std::vector<double> x((size_t)N * Dp), l(N), lp(N);
    std::vector<int> ord(N);
    CK(cudaMemcpy(x.data(), c->d_x, x.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(l.data(), c->d_logL, N * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lp.data(), c->d_logP, N * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ord.data(), c->d_order[c->cur], N * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i) {
        const int s = ord[i];
        if (s < 0 || s >= N) return fail("corrupt live order at %d: %d", i, s);
        for (int d = 0; d < D; ++d) rows[(size_t)i * (D + 2) + d] = x[(size_t)s * Dp + d];
        rows[(size_t)i * (D + 2) + D] = l[s];
        rows[(size_t)i * (D + 2) + D + 1] = lp[s];
    }
    return 0;
}

extern "C" int64_t cpn_num_dead(cpn_ctx* c) {
    if (!c) return -1;
    cudaSetDevice(c->cfg.device);
    if (fetch_state(c)) return -1;
    return c->h_state->iteration;
}

extern "C" int cpn_get_dead(cpn_ctx* c, int64_t first, int64_t n, double* rows, double* log_vols) {
    if (!c || !rows) return fail("null argument");
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    if (first < 0 || n < 0 || first + n > c->h_state->iteration) return fail("cpn_get_dead: range [%lld,%lld) outside %lld dead points", (long long)first, (long long)(first + n), c->h_state->iteration);
    const int D = c->D, Dp = c->Dp;
    std::vector<double> x((size_t)n * Dp), l(n), lp(n);
    CK(cudaMemcpy(x.data(), c->d_dead_x + (size_t)first * Dp, x.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(l.data(), c->d_dead_logL + first, n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lp.data(), c->d_dead_logP + first, n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) {
        for (int d = 0; d < D; ++d) rows[(size_t)i * (D + 2) + d] = x[(size_t)i * Dp + d];
        rows[(size_t)i * (D + 2) + D] = l[i];
        rows[(size_t)i * (D + 2) + D + 1] = lp[i];
    }
    if (log_vols) {
        if (c->cfg.ns_mode != CPN_NS_SEQUENTIAL) return fail("per-point log volumes exist only in sequential mode");
        CK(cudaMemcpy(log_vols, c->d_hist_logvol + first, n * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}

extern "C" int64_t cpn_num_insertion_indices(cpn_ctx* c) {
    if (!c) return -1;
    cudaSetDevice(c->cfg.device);
    if (fetch_state(c)) return -1;
    return c->h_state->n_ins;
}

extern "C" int cpn_get_insertion_indices(cpn_ctx* c, int64_t first, int64_t n, double* out) {
    if (!c || !out) return fail("null argument");
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    if (first < 0 || n < 0 || first + n > c->h_state->n_ins) return fail("insertion index range out of bounds");
    CK(cudaMemcpy(out, c->d_ins + first, n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int cpn_get_last_batch(cpn_ctx* c, int32_t K, double* rows, int32_t* steps, int32_t* accepted) {
    if (!c) return fail("null context");
    if (K < 1 || K > c->N) return fail("K out of range");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    const int D = c->D, Dp = c->Dp;
    if (rows) {
        std::vector<double> x((size_t)K * Dp), l(K), lp(K);
        CK(cudaMemcpy(x.data(), c->d_new_x, x.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(l.data(), c->d_new_logL, K * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(lp.data(), c->d_new_logP, K * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < K; ++i) {
            for (int d = 0; d < D; ++d) rows[(size_t)i * (D + 2) + d] = x[(size_t)i * Dp + d];
            rows[(size_t)i * (D + 2) + D] = l[i];
            rows[(size_t)i * (D + 2) + D + 1] = lp[i];
        }
    }
    if (steps) CK(cudaMemcpy(steps, c->d_new_steps, K * sizeof(int), cudaMemcpyDeviceToHost));
    if (accepted) CK(cudaMemcpy(accepted, c->d_new_acc, K * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

// ---- host protocol: requests in, replies out (sampler.py:233-246) ---------------------------------------
__global__ void k_unpack_records(const double* rec, int n, int D, int Dp, double* x, double* logL, double* logP) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * Dp) return;
    const int r = i / Dp, d = i % Dp;
    x[i] = (d < D) ? rec[(size_t)r * (D + 2) + d] : 0.0;
    if (d == 0) {
        logL[r] = rec[(size_t)r * (D + 2) + D];
        logP[r] = rec[(size_t)r * (D + 2) + D + 1];
    }
}
__global__ void k_pack_records(const double* x, const double* logL, const double* logP, int n, int D, int Dp, double* rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (D + 2)) return;
    const int r = i / (D + 2), d = i % (D + 2);
    rec[i] = (d < D) ? x[(size_t)r * Dp + d] : (d == D ? logL[r] : logP[r]);
}

__global__ void k_pool_append(const int* slots, int K, int N, int Dp, const double* new_x, const double* new_logL,
                              const double* new_logP, double* x, double* logL, double* logP) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Dp) return;
    const int j = i / Dp, d = i % Dp;
    const int s = slots[j];
    if (s < 0 || s >= N) return;
    x[(size_t)s * Dp + d] = new_x[i];
    if (d == 0) { logL[s] = new_logL[j]; logP[s] = new_logP[j]; }
}

extern "C" int cpn_evolve_host(cpn_ctx* c, int32_t K, double logLmin, int32_t nmcmc, const double* requests,
                               const int32_t* slots, double* replies, int32_t* steps, int32_t* accepted) {
    if (!c || !requests || !replies) return fail("null argument");
    if (!c->live_ready) return fail("live set not initialised");
    if (K < 1 || K > c->N) return fail("K=%d must be in [1, nlive]", K);
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    double* d_rec = c->d_rec;
    const size_t need = (size_t)K * (D + 2);
    CK(cudaMemcpyAsync(d_rec, requests, need * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int T = 256;
    k_unpack_records<<<(K * Dp + T - 1) / T, T, 0, c->stream>>>(d_rec, K, D, Dp, c->d_start_x, c->d_start_logL, c->d_start_logP);
    // the host protocol is served by one GPU (n_out = 1)
    if (launch_chains(c, K, 0, K, START_GIVEN, 0, true, logLmin, std::max(1, std::min(nmcmc, c->cfg.maxmcmc)), 1)) return 1;
    for (int k = 0; k < kNumKinds; ++k) {
        const int a = c->last_kb[k], b = c->last_kb[k + 1];
        if (b <= a) continue;
        k_batch_counters<<<1, kCorrThreads, 0, c->stream>>>(c->d_state, k, a, b, c->d_new_steps, c->d_new_acc, c->d_new_evals);
        k_adapt_nmcmc<<<1, 1, 0, c->stream>>>(c->d_state, k, 0, nullptr, nullptr, 0);      // folds the batch counters into the totals
        k_adapt_nmcmc_on_the_fly<<<1, kCorrThreads, 0, c->stream>>>(c->d_state, k, a, b, c->d_new_steps, c->d_new_acc);
        if (k == CPN_SAMPLER_SLICE) k_adapt_slice_mu<<<1, 256, 0, c->stream>>>(c->d_state, a, b, c->d_new_ne, c->d_new_nc);
        c->launches += 3 + (k == CPN_SAMPLER_SLICE);
    }
    if (slots) {
        CK(cudaMemcpyAsync(c->d_rank, slots, K * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        k_pool_append<<<(K * Dp + T - 1) / T, T, 0, c->stream>>>(c->d_rank, K, c->N, Dp, c->d_new_x, c->d_new_logL,
                                                              c->d_new_logP, c->d_x, c->d_logL, c->d_logP);
        c->launches += 1;
    }
    c->launches += 2;
    k_pack_records<<<((int)need + T - 1) / T, T, 0, c->stream>>>(c->d_new_x, c->d_new_logL, c->d_new_logP, K, D, Dp, d_rec);
    CKL();
    CK(cudaMemcpyAsync(replies, d_rec, need * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (steps) CK(cudaMemcpyAsync(steps, c->d_new_steps, K * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (accepted) CK(cudaMemcpyAsync(accepted, c->d_new_acc, K * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->chain_counter += (unsigned long long)K;
    c->batches_enqueued += 1;
    return 0;
}

// ---- host protocol over a pinned host table: the device gathers the requests and scatters the replies itself ----
__global__ void k_gather_table(const double* __restrict__ table, const int* __restrict__ idx, int K, int N, int D, int Dp,
                               double* __restrict__ x, double* __restrict__ logL, double* __restrict__ logP) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Dp) return;
    const int j = i / Dp, d = i % Dp;
    const int r = min(max(idx[j], 0), N - 1);
    const double* row = table + (size_t)r * (D + 2);
    x[i] = (d < D) ? row[d] : 0.0;
    if (d == 0) { logL[j] = row[D]; logP[j] = row[D + 1]; }
}
// columns [d0, d1) of the K reply records into rows slots[j] of the host table
__global__ void k_scatter_table(double* __restrict__ table, const int* __restrict__ slots, int K, int N, int D, int Dp, int d0, int d1,
                                const double* __restrict__ x, const double* __restrict__ logL, const double* __restrict__ logP) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = d1 - d0;
    if (i >= K * w) return;
    const int j = i / w, d = d0 + i % w;
    const int r = slots[j];
    if (r < 0 || r >= N) return;
    table[(size_t)r * (D + 2) + d] = (d < D) ? x[(size_t)j * Dp + d] : (d == D ? logL[j] : logP[j]);
}

extern "C" int cpn_evolve_host_table(cpn_ctx* c, int32_t K, double logLmin, int32_t nmcmc, double* table,
                                     const int32_t* start_idx, const int32_t* slots, int32_t* steps, int32_t* accepted,
                                     int32_t flags) {
    if (!c || !table || !start_idx || !slots) return fail("null argument");
    if (!c->live_ready) return fail("live set not initialised");
    if (K < 1 || K > c->N) return fail("K=%d must be in [1, nlive]", K);
    CK(cudaSetDevice(c->cfg.device));
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, table) != cudaSuccess || at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) {
        cudaGetLastError();
        return fail("cpn_evolve_host_table: the table must be pinned host memory (cudaHostAlloc / cudaHostRegister, e.g. "
                    "torch.Tensor.pin_memory()); use cpn_evolve_host for pageable buffers");
    }
    double* d_table = (double*)at.devicePointer;
    // checked and moved into pinned memory in one pass (the previous call has synchronised: the buffer is free)
    int* h_start = c->h_idx;
    int* h_slots = c->h_idx + c->N;
    for (int j = 0; j < K; ++j) {
        if (start_idx[j] < 0 || start_idx[j] >= c->N) return fail("cpn_evolve_host_table: start_idx[%d] = %d out of range", j, start_idx[j]);
        if (slots[j] < 0 || slots[j] >= c->N) return fail("cpn_evolve_host_table: slots[%d] = %d out of range", j, slots[j]);
        h_start[j] = start_idx[j];
        h_slots[j] = slots[j];
    }
    const int D = c->D, Dp = c->Dp, T = 256;
    int* d_idx = c->d_sort_idx;                         // scratch of the initial sort, free during a run
    CK(cudaMemcpyAsync(d_idx, h_start, K * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_rank, h_slots, K * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_gather_table<<<(K * Dp + T - 1) / T, T, 0, c->stream>>>(d_table, d_idx, K, c->N, D, Dp, c->d_start_x, c->d_start_logL, c->d_start_logP);
    if (launch_chains(c, K, 0, K, START_GIVEN, 0, true, logLmin, std::max(1, std::min(nmcmc, c->cfg.maxmcmc)), 1)) return 1;
    for (int k = 0; k < kNumKinds; ++k) {
        const int a = c->last_kb[k], b = c->last_kb[k + 1];
        if (b <= a) continue;
        k_batch_counters<<<1, kCorrThreads, 0, c->stream>>>(c->d_state, k, a, b, c->d_new_steps, c->d_new_acc, c->d_new_evals);
        k_adapt_nmcmc<<<1, 1, 0, c->stream>>>(c->d_state, k, 0, nullptr, nullptr, 0);      // folds the batch counters into the totals
        k_adapt_nmcmc_on_the_fly<<<1, kCorrThreads, 0, c->stream>>>(c->d_state, k, a, b, c->d_new_steps, c->d_new_acc);
        if (k == CPN_SAMPLER_SLICE) k_adapt_slice_mu<<<1, 256, 0, c->stream>>>(c->d_state, a, b, c->d_new_ne, c->d_new_nc);
        c->launches += 3 + (k == CPN_SAMPLER_SLICE);
    }
    k_pool_append<<<(K * Dp + T - 1) / T, T, 0, c->stream>>>(c->d_rank, K, c->N, Dp, c->d_new_x, c->d_new_logL, c->d_new_logP,
                                                          c->d_x, c->d_logL, c->d_logP);
    // the keys first (logL, logP: what the caller's next selection reads), then the coordinates
    k_scatter_table<<<(K * 2 + T - 1) / T, T, 0, c->stream>>>(d_table, c->d_rank, K, c->N, D, Dp, D, D + 2, c->d_new_x, c->d_new_logL,
                                                             c->d_new_logP);
    if (steps) CK(cudaMemcpyAsync(steps, c->d_new_steps, K * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (accepted) CK(cudaMemcpyAsync(accepted, c->d_new_acc, K * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev, c->stream));
    k_scatter_table<<<(K * D + T - 1) / T, T, 0, c->stream>>>(d_table, c->d_rank, K, c->N, D, Dp, 0, D, c->d_new_x, c->d_new_logL,
                                                             c->d_new_logP);
    CKL();
    c->launches += 4;
    if (flags & CPN_TABLE_ROWS_ASYNC) CK(cudaEventSynchronize(c->ev));     // the rows land while the caller selects
    else CK(cudaStreamSynchronize(c->stream));
    c->chain_counter += (unsigned long long)K;
    c->batches_enqueued += 1;
    return 0;
}

// the caller's side of the protocol: rows of a host table by index (see the header)
extern "C" int cpn_host_gather_rows(const double* table, int64_t nrows, int32_t row_doubles, const int64_t* idx, int32_t n,
                                    double* out) {
    if (!table || !idx || !out || row_doubles < 1 || n < 0) return fail("cpn_host_gather_rows: bad argument");
    const size_t rb = (size_t)row_doubles * sizeof(double);
    for (int32_t j = 0; j < n; ++j) {
        const int64_t i = idx[j];
        if (i < 0 || i >= nrows) return fail("cpn_host_gather_rows: index %lld at %d out of range [0, %lld)", (long long)i, j, (long long)nrows);
        memcpy(out + (size_t)j * row_doubles, table + (size_t)i * row_doubles, rb);
    }
    return 0;
}
extern "C" int cpn_host_scatter_rows(double* table, int64_t nrows, int32_t row_doubles, const int64_t* idx, int32_t n,
                                     const double* rows) {
    if (!table || !idx || !rows || row_doubles < 1 || n < 0) return fail("cpn_host_scatter_rows: bad argument");
    const size_t rb = (size_t)row_doubles * sizeof(double);
    for (int32_t j = 0; j < n; ++j) {
        const int64_t i = idx[j];
        if (i < 0 || i >= nrows) return fail("cpn_host_scatter_rows: index %lld at %d out of range [0, %lld)", (long long)i, j, (long long)nrows);
    }
    for (int32_t j = 0; j < n; ++j) memcpy(table + (size_t)idx[j] * row_doubles, rows + (size_t)j * row_doubles, rb);
    return 0;
}

// ---- parity entry points --------------------------------------------------------------------------------------
extern "C" int cpn_eval_model(cpn_ctx* c, const double* x, int32_t n, double* logL, double* logP) {
    if (!c || !x || !logL || !logP || n < 1) return fail("bad argument");
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    std::vector<double> xp((size_t)n * Dp, 0.0);
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < D; ++d) xp[(size_t)i * Dp + d] = x[(size_t)i * D + d];
    double *dx, *dl, *dp;
    if (dalloc(&dx, xp.size()) || dalloc(&dl, n) || dalloc(&dp, n)) return 1;
    CK(h2d_sync(dx, xp.data(), xp.size() * sizeof(double)));
    EvalParams e;
    memset(&e, 0, sizeof e);
    e.x = dx; e.logL = dl; e.logP = dp; e.n = n; e.D = D; e.Dp = Dp; e.lo = c->d_lo; e.hi = c->d_hi; e.blob = c->d_blob;
    if (run_eval(c, c->G, c->E, e)) return 1;
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(logL, dl, n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(logP, dp, n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dl); cudaFree(dp);
    return 0;
}

extern "C" int cpn_eval_gradients(cpn_ctx* c, const double* x, int32_t n, double* grad_logL, double* grad_V) {
    if (!c || !x || !grad_logL || !grad_V || n < 1) return fail("bad argument");
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    std::vector<double> xp((size_t)n * Dp, 0.0);
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < D; ++d) xp[(size_t)i * Dp + d] = x[(size_t)i * D + d];
    double *dx, *dl, *dp, *dgl, *dgv;
    if (dalloc(&dx, xp.size()) || dalloc(&dl, n) || dalloc(&dp, n) || dalloc(&dgl, xp.size()) || dalloc(&dgv, xp.size())) return 1;
    CK(h2d_sync(dx, xp.data(), xp.size() * sizeof(double)));
    EvalParams e;
    memset(&e, 0, sizeof e);
    e.x = dx; e.logL = dl; e.logP = dp; e.n = n; e.D = D; e.Dp = Dp; e.lo = c->d_lo; e.hi = c->d_hi; e.blob = c->d_blob;
    e.grad_logL = dgl; e.grad_V = dgv;
    if (run_eval(c, c->G, c->E, e)) return 1;
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    std::vector<double> a(xp.size()), b(xp.size());
    CK(cudaMemcpy(a.data(), dgl, a.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), dgv, b.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < D; ++d) {
            grad_logL[(size_t)i * D + d] = a[(size_t)i * Dp + d];
            grad_V[(size_t)i * D + d] = b[(size_t)i * Dp + d];
        }
    cudaFree(dx); cudaFree(dl); cudaFree(dp); cudaFree(dgl); cudaFree(dgv);
    return 0;
}

extern "C" int cpn_replay_mh(cpn_ctx* c, int32_t P, const double* ensemble, int32_t C, const double* start,
                             const double* eigval, const double* eigvec, int32_t cycle_idx, const double* tape,
                             int32_t nmcmc, int32_t maxmcmc, double logLmin, int32_t compat, int32_t lanes,
                             double* out, int32_t* steps, int32_t* accepted, double* states) {
    if (!c || !ensemble || !start || !tape || !out) return fail("null argument");
    if (P < 3 || C < 1 || maxmcmc < 1) return fail("bad sizes");
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    int G, E;
    if (pick_config(D, lanes > 0 ? lanes : c->cfg.lanes, c->cfg.functor, C, &G, &E)) return 1;
    std::vector<double> ex((size_t)P * Dp, 0.0), sx((size_t)C * Dp, 0.0), sl(C), sp(C), ev((size_t)D * Dp, 0.0), es(D, 0.0);
    for (int i = 0; i < P; ++i)
        for (int d = 0; d < D; ++d) ex[(size_t)i * Dp + d] = ensemble[(size_t)i * D + d];
    for (int i = 0; i < C; ++i) {
        for (int d = 0; d < D; ++d) sx[(size_t)i * Dp + d] = start[(size_t)i * (D + 2) + d];
        sl[i] = start[(size_t)i * (D + 2) + D];
        sp[i] = start[(size_t)i * (D + 2) + D + 1];
    }
    if (eigval && eigvec)
        for (int i = 0; i < D; ++i) {
            es[i] = sqrt(fabs(eigval[i]));                                   // proposal.py:339
            for (int k = 0; k < D; ++k) ev[(size_t)i * Dp + k] = eigvec[(size_t)k * D + i];
        }
    double *dex, *dsx, *dsl, *dsp, *dev, *des, *dtape, *dox, *dol, *dop, *dstates = nullptr, *dEL, *dEP;
    int *dst, *dac;
    if (dalloc(&dex, ex.size()) || dalloc(&dsx, sx.size()) || dalloc(&dsl, C) || dalloc(&dsp, C) || dalloc(&dev, ev.size()) ||
        dalloc(&des, D) || dalloc(&dtape, (size_t)C * maxmcmc * 8) || dalloc(&dox, (size_t)C * Dp) || dalloc(&dol, C) ||
        dalloc(&dop, C) || dalloc(&dst, C) || dalloc(&dac, C) || dalloc(&dEL, P) || dalloc(&dEP, P))
        return 1;
    if (states && dalloc(&dstates, (size_t)C * maxmcmc * D)) return 1;
    CK(h2d_sync(dex, ex.data(), ex.size() * sizeof(double)));
    CK(h2d_sync(dsx, sx.data(), sx.size() * sizeof(double)));
    CK(h2d_sync(dsl, sl.data(), C * sizeof(double)));
    CK(h2d_sync(dsp, sp.data(), C * sizeof(double)));
    CK(h2d_sync(dev, ev.data(), ev.size() * sizeof(double)));
    CK(h2d_sync(des, es.data(), D * sizeof(double)));
    CK(h2d_sync(dtape, tape, (size_t)C * maxmcmc * 8 * sizeof(double)));
    MHParams p;
    memset(&p, 0, sizeof p);
    p.ens_x = dex; p.ens_logL = dEL; p.ens_logP = dEP; p.ens_n = P; p.Dp = Dp; p.D = D;
    p.start_mode = START_GIVEN; p.start_x = dsx; p.start_logL = dsl; p.start_logP = dsp;
    p.lo = c->d_lo; p.hi = c->d_hi;
    p.cycle = c->d_cycle; p.cycle_len = c->cycle_len; p.cycle_pos0 = cycle_idx;
    p.eig_scale = des; p.eig_vec = dev;
    p.state = nullptr; p.use_override = 1; p.logLmin_override = logLmin; p.nmcmc_override = nmcmc; p.maxmcmc = maxmcmc;
    p.j0 = 0; p.j1 = C; p.blob = c->d_blob;
    p.n_out = 1; p.out_x[0] = dox; p.out_logL[0] = dol; p.out_logP[0] = dop; p.out_steps[0] = dst; p.out_acc[0] = dac;
    p.tape = dtape; p.compat = compat; p.states = dstates;
    if (run_mh(c, G, E, 1, p)) return 1;
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    std::vector<double> ox((size_t)C * Dp), ol(C), op(C);
    CK(cudaMemcpy(ox.data(), dox, ox.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ol.data(), dol, C * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(op.data(), dop, C * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < C; ++i) {
        for (int d = 0; d < D; ++d) out[(size_t)i * (D + 2) + d] = ox[(size_t)i * Dp + d];
        out[(size_t)i * (D + 2) + D] = ol[i];
        out[(size_t)i * (D + 2) + D + 1] = op[i];
    }
    if (steps) CK(cudaMemcpy(steps, dst, C * sizeof(int), cudaMemcpyDeviceToHost));
    if (accepted) CK(cudaMemcpy(accepted, dac, C * sizeof(int), cudaMemcpyDeviceToHost));
    if (states) CK(cudaMemcpy(states, dstates, (size_t)C * maxmcmc * D * sizeof(double), cudaMemcpyDeviceToHost));
    void* fr[] = {dex, dsx, dsl, dsp, dev, des, dtape, dox, dol, dop, dst, dac, dEL, dEP, dstates};
    for (void* q : fr)
        if (q) cudaFree(q);
    return 0;
}

// The PRODUCTION MH kernel (the instantiation every run and the benchmark launch) on given chains, with a record of
// what every step consumed: see the header.
extern "C" int cpn_trace_mh(cpn_ctx* c, int32_t P, const double* ensemble, int32_t C, const double* start, const double* eigval,
                            const double* eigvec, int32_t cycle_idx, int32_t nmcmc, int32_t maxmcmc, double logLmin, int32_t lanes,
                            uint64_t chain_base, double* out, int32_t* steps, int32_t* accepted, int32_t* evals, double* draws) {
    if (!c || !ensemble || !start || !out || !draws) return fail("null argument");
    if (P < 4 || C < 1 || maxmcmc < 1 || nmcmc < 1) return fail("bad sizes");
    if (c->cfg.functor == CPN_USER) return fail("cpn_trace_mh is not available for a user functor");
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    int G, E;
    if (pick_config(D, lanes > 0 ? lanes : c->cfg.lanes, c->cfg.functor, C, &G, &E)) return 1;
    std::vector<double> ex((size_t)P * Dp, 0.0), sx((size_t)C * Dp, 0.0), sl(C), sp(C), ev((size_t)D * Dp, 0.0), es(D, 0.0);
    for (int i = 0; i < P; ++i)
        for (int d = 0; d < D; ++d) ex[(size_t)i * Dp + d] = ensemble[(size_t)i * D + d];
    for (int i = 0; i < C; ++i) {
        for (int d = 0; d < D; ++d) sx[(size_t)i * Dp + d] = start[(size_t)i * (D + 2) + d];
        sl[i] = start[(size_t)i * (D + 2) + D];
        sp[i] = start[(size_t)i * (D + 2) + D + 1];
    }
    if (eigval && eigvec)
        for (int i = 0; i < D; ++i) {
            es[i] = sqrt(fabs(eigval[i]));                                   // proposal.py:339
            for (int k = 0; k < D; ++k) ev[(size_t)i * Dp + k] = eigvec[(size_t)k * D + i];
        }
    const size_t ndraw = (size_t)C * maxmcmc * 8;
    double *dex = nullptr, *dsx = nullptr, *dsl = nullptr, *dsp = nullptr, *dev = nullptr, *des = nullptr, *dox = nullptr, *dol = nullptr,
           *dop = nullptr, *ddraw = nullptr;
    int *dst = nullptr, *dac = nullptr, *dnev = nullptr;
    int rc = 1;
    do {
        // (the eigen-directions behind the ensemble rows, as in the engine's own arrays)
        if (dalloc(&dex, ex.size() + ev.size() + kRowSlack) || dalloc(&dsx, sx.size()) || dalloc(&dsl, C) || dalloc(&dsp, C) ||
            dalloc(&des, D) || dalloc(&dox, (size_t)C * Dp) || dalloc(&dol, C) ||
            dalloc(&dop, C) || dalloc(&dst, C) || dalloc(&dac, C) || dalloc(&dnev, C) || dalloc(&ddraw, ndraw))
            break;
        dev = dex + ex.size();
        if (h2d_sync(dex, ex.data(), ex.size() * sizeof(double)) != cudaSuccess || h2d_sync(dsx, sx.data(), sx.size() * sizeof(double)) != cudaSuccess ||
            h2d_sync(dsl, sl.data(), C * sizeof(double)) != cudaSuccess || h2d_sync(dsp, sp.data(), C * sizeof(double)) != cudaSuccess ||
            h2d_sync(dev, ev.data(), ev.size() * sizeof(double)) != cudaSuccess || h2d_sync(des, es.data(), D * sizeof(double)) != cudaSuccess) {
            fail("cpn_trace_mh: host to device copy failed");
            break;
        }
        MHParams p;
        memset(&p, 0, sizeof p);
        p.ens_x = dex; p.ens_n = P; p.Dp = Dp; p.D = D;
        p.start_mode = START_GIVEN; p.start_x = dsx; p.start_logL = dsl; p.start_logP = dsp;
        p.lo = c->d_lo; p.hi = c->d_hi;
        p.cycle = c->d_cycle; p.cycle_len = c->cycle_len; p.cycle_pos0 = cycle_idx % c->cycle_len;
        fill_inline_cycle(&p, c->h_cycle, c->cycle_len);
        p.eig_scale = des; p.eig_vec = dev; p.eig_row0 = P;
        p.state = nullptr; p.use_override = 1; p.logLmin_override = logLmin; p.nmcmc_override = nmcmc; p.maxmcmc = maxmcmc;
        p.j0 = 0; p.j1 = C; p.blob = c->d_blob;
        p.padded = Dp >= G * E;
        p.chain_base = chain_base;
        p.seed_lo = (uint32_t)c->cfg.seed; p.seed_hi = (uint32_t)(c->cfg.seed >> 32);
        p.n_out = 1; p.out_x[0] = dox; p.out_logL[0] = dol; p.out_logP[0] = dop; p.out_steps[0] = dst; p.out_acc[0] = dac;
        p.out_evals[0] = dnev;
        p.draws = ddraw;
        if (run_mh(c, G, E, 0, p)) break;
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { fail("cpn_trace_mh: launch failed"); break; }
        std::vector<double> ox((size_t)C * Dp), ol(C), op(C);
        if (cudaMemcpy(ox.data(), dox, ox.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(ol.data(), dol, C * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(op.data(), dop, C * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(draws, ddraw, ndraw * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            (steps && cudaMemcpy(steps, dst, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (accepted && cudaMemcpy(accepted, dac, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (evals && cudaMemcpy(evals, dnev, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)) {
            fail("cpn_trace_mh: device to host copy failed");
            break;
        }
        for (int i = 0; i < C; ++i) {
            for (int d = 0; d < D; ++d) out[(size_t)i * (D + 2) + d] = ox[(size_t)i * Dp + d];
            out[(size_t)i * (D + 2) + D] = ol[i];
            out[(size_t)i * (D + 2) + D + 1] = op[i];
        }
        rc = 0;
    } while (0);
    void* fr[] = {dex, dsx, dsl, dsp, des, dox, dol, dop, dst, dac, dnev, ddraw};
    for (void* q : fr)
        if (q) cudaFree(q);
    return rc;
}

static int slice_scratch_launch(cpn_ctx* c, SliceParams& p, int replay, int lanes, int C) {
    int G, E;
    if (pick_config(c->D, lanes > 0 ? lanes : c->cfg.lanes, c->cfg.functor, C, &G, &E)) return 1;
    if (run_slice(c, G, E, replay, p)) return 1;
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// slice chains on a given ensemble from given starts: the replay instantiation on explicit tapes, or the production
// instantiation (Philox draws) recording what it consumed (trace)
static int slice_on_given(cpn_ctx* c, bool replay, int32_t P, const double* ensemble, int32_t C, const double* start,
                          int32_t cycle_idx, const double* tape, const int64_t* tape_off, double mu, int32_t nmcmc,
                          int32_t maxmcmc, double logLmin, int32_t compat, int32_t lanes, uint64_t chain_base, const double* mean,
                          const double* eigval, const double* eigvec, double* out, int32_t* steps, int32_t* accepted,
                          int32_t* evals, double* last, double* trace, int32_t trace_cap) {
    if (P < 2 || C < 1 || maxmcmc < 1) return fail("bad sizes");
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    std::vector<double> ex((size_t)P * Dp, 0.0), sx((size_t)C * Dp, 0.0), sl(C), sp(C), ev((size_t)D * Dp, 0.0), es(D, 0.0), mn(D, 0.0);
    for (int i = 0; i < P; ++i)
        for (int d = 0; d < D; ++d) ex[(size_t)i * Dp + d] = ensemble[(size_t)i * D + d];
    for (int i = 0; i < C; ++i) {
        for (int d = 0; d < D; ++d) sx[(size_t)i * Dp + d] = start[(size_t)i * (D + 2) + d];
        sl[i] = start[(size_t)i * (D + 2) + D];
        sp[i] = start[(size_t)i * (D + 2) + D + 1];
    }
    if (mean && eigval && eigvec)
        for (int i = 0; i < D; ++i) {
            mn[i] = mean[i];
            es[i] = sqrt(fabs(eigval[i]));
            for (int k = 0; k < D; ++k) ev[(size_t)i * Dp + k] = eigvec[(size_t)k * D + i];   // numpy layout: column i = eigenvector i
        }
    const size_t ntape = replay ? (size_t)tape_off[C] : 0;
    const size_t ntrace = trace ? (size_t)C * trace_cap : 0;
    double *dex = nullptr, *dsx = nullptr, *dsl = nullptr, *dsp = nullptr, *dtape = nullptr, *dox = nullptr, *dol = nullptr, *dop = nullptr,
           *dlast = nullptr, *dEL = nullptr, *dEP = nullptr, *dmn = nullptr, *des = nullptr, *dvec = nullptr, *dtr = nullptr;
    long long* doff = nullptr;
    int *dst = nullptr, *dac = nullptr, *dev = nullptr;
    int rc = 1;
    do {
        if (dalloc(&dex, ex.size()) || dalloc(&dsx, sx.size()) || dalloc(&dsl, C) || dalloc(&dsp, C) || dalloc(&dtape, ntape + 8) ||
            dalloc(&doff, C + 1) || dalloc(&dox, (size_t)C * Dp) || dalloc(&dol, C) || dalloc(&dop, C) || dalloc(&dst, C) ||
            dalloc(&dac, C) || dalloc(&dev, C) || dalloc(&dlast, (size_t)C * 4) || dalloc(&dEL, P) || dalloc(&dEP, P) ||
            dalloc(&dmn, D) || dalloc(&des, D) || dalloc(&dvec, ev.size()) || dalloc(&dtr, ntrace + 1))
            break;
        if (h2d_sync(dex, ex.data(), ex.size() * sizeof(double)) != cudaSuccess || h2d_sync(dsx, sx.data(), sx.size() * sizeof(double)) != cudaSuccess ||
            h2d_sync(dsl, sl.data(), C * sizeof(double)) != cudaSuccess || h2d_sync(dsp, sp.data(), C * sizeof(double)) != cudaSuccess ||
            h2d_sync(dmn, mn.data(), D * sizeof(double)) != cudaSuccess || h2d_sync(des, es.data(), D * sizeof(double)) != cudaSuccess ||
            h2d_sync(dvec, ev.data(), ev.size() * sizeof(double)) != cudaSuccess ||
            (replay && (h2d_sync(dtape, tape, ntape * sizeof(double)) != cudaSuccess ||
                        h2d_sync(doff, tape_off, (C + 1) * sizeof(long long)) != cudaSuccess))) {
            fail("slice chains on a given ensemble: host to device copy failed");
            break;
        }
        SliceParams p;
        memset(&p, 0, sizeof p);
        p.ens_x = dex; p.ens_logL = dEL; p.ens_logP = dEP; p.ens_n = P; p.Dp = Dp; p.D = D;
        p.start_mode = START_GIVEN; p.start_x = dsx; p.start_logL = dsl; p.start_logP = dsp;
        p.lo = c->d_lo; p.hi = c->d_hi;
        p.cycle = c->d_slice_cycle; p.cycle_len = c->slice_cycle_len; p.cycle_pos0 = cycle_idx % c->slice_cycle_len;
        p.mean = dmn; p.eig_scale = des; p.eig_vec = dvec;
        p.state = nullptr; p.use_override = 1; p.logLmin_override = logLmin; p.nmcmc_override = nmcmc; p.mu_override = mu;
        p.maxmcmc = maxmcmc; p.j0 = 0; p.j1 = C; p.blob = c->d_blob;
        p.chain_base = chain_base;
        p.seed_lo = (uint32_t)c->cfg.seed; p.seed_hi = (uint32_t)(c->cfg.seed >> 32);
        p.n_out = 1; p.out_x[0] = dox; p.out_logL[0] = dol; p.out_logP[0] = dop; p.out_steps[0] = dst; p.out_acc[0] = dac;
        p.out_evals[0] = dev;
        p.tape = dtape; p.tape_off = doff; p.compat = compat; p.last = dlast;
        if (trace) { p.trace = dtr; p.trace_cap = trace_cap; }
        if (slice_scratch_launch(c, p, replay ? 1 : 0, lanes, C)) break;
        std::vector<double> ox((size_t)C * Dp), ol(C), op(C);
        if (cudaMemcpy(ox.data(), dox, ox.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(ol.data(), dol, C * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(op.data(), dop, C * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            (steps && cudaMemcpy(steps, dst, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (accepted && cudaMemcpy(accepted, dac, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (evals && cudaMemcpy(evals, dev, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (last && cudaMemcpy(last, dlast, (size_t)C * 4 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (trace && cudaMemcpy(trace, dtr, ntrace * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)) {
            fail("slice chains on a given ensemble: device to host copy failed");
            break;
        }
        for (int i = 0; i < C; ++i) {
            for (int d = 0; d < D; ++d) out[(size_t)i * (D + 2) + d] = ox[(size_t)i * Dp + d];
            out[(size_t)i * (D + 2) + D] = ol[i];
            out[(size_t)i * (D + 2) + D + 1] = op[i];
        }
        rc = 0;
    } while (0);
    void* fr[] = {dex, dsx, dsl, dsp, dtape, doff, dox, dol, dop, dst, dac, dev, dlast, dEL, dEP, dmn, des, dvec, dtr};
    for (void* q : fr)
        if (q) cudaFree(q);
    return rc;
}

extern "C" int cpn_replay_slice(cpn_ctx* c, int32_t P, const double* ensemble, int32_t C, const double* start,
                                int32_t cycle_idx, const double* tape, const int64_t* tape_off, double mu, int32_t nmcmc,
                                int32_t maxmcmc, double logLmin, int32_t compat, int32_t lanes, double* out, int32_t* steps,
                                int32_t* accepted, int32_t* evals, double* last) {
    if (!c || !ensemble || !start || !tape || !tape_off || !out) return fail("null argument");
    return slice_on_given(c, true, P, ensemble, C, start, cycle_idx, tape, tape_off, mu, nmcmc, maxmcmc, logLmin, compat, lanes, 0,
                          nullptr, nullptr, nullptr, out, steps, accepted, evals, last, nullptr, 0);
}

extern "C" int cpn_trace_slice(cpn_ctx* c, int32_t P, const double* ensemble, int32_t C, const double* start,
                               const double* mean, const double* eigval, const double* eigvec, int32_t cycle_idx, double mu,
                               int32_t nmcmc, int32_t maxmcmc, double logLmin, int32_t lanes, uint64_t chain_base, double* out,
                               int32_t* steps, int32_t* accepted, int32_t* evals, double* last, double* trace, int32_t trace_cap) {
    if (!c || !ensemble || !start || !out || !trace) return fail("null argument");
    if (trace_cap < c->D + 16) return fail("cpn_trace_slice: trace_cap too small");
    if (c->cfg.functor == CPN_USER) return fail("cpn_trace_slice is not available for a user functor");
    return slice_on_given(c, false, P, ensemble, C, start, cycle_idx, nullptr, nullptr, mu, nmcmc, maxmcmc, logLmin, 0, lanes,
                          chain_base, mean, eigval, eigvec, out, steps, accepted, evals, last, trace, trace_cap);
}

extern "C" int cpn_replay_hmc(cpn_ctx* c, int32_t C, const double* start, const double* cov, double logdet,
                              const double* tape, const int64_t* tape_off, double dt, int32_t L, int32_t max_traj,
                              double logLmin, int32_t lanes, double* out, int32_t* steps, int32_t* accepted, int32_t* evals,
                              double* log_J) {
    if (!c || !start || !cov || !tape || !tape_off || !out) return fail("null argument");
    if (C < 1 || L < 1 || max_traj < 1) return fail("bad sizes");
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    std::vector<double> sx((size_t)C * Dp, 0.0), sl(C), sp(C);
    for (int i = 0; i < C; ++i) {
        for (int d = 0; d < D; ++d) sx[(size_t)i * Dp + d] = start[(size_t)i * (D + 2) + d];
        sl[i] = start[(size_t)i * (D + 2) + D];
        sp[i] = start[(size_t)i * (D + 2) + D + 1];
    }
    const size_t ntape = (size_t)tape_off[C];
    const size_t ntraj = (size_t)C * (2 * L + 1) * (Dp + 3);
    double *dsx, *dsl, *dsp, *dcov, *dtape, *dtraj, *dox, *dol, *dop, *dlj;
    long long* doff;
    int *dst, *dac, *dev;
    if (dalloc(&dsx, sx.size()) || dalloc(&dsl, C) || dalloc(&dsp, C) || dalloc(&dcov, (size_t)D * D) || dalloc(&dtape, ntape + 8) ||
        dalloc(&doff, C + 1) || dalloc(&dtraj, ntraj) || dalloc(&dox, (size_t)C * Dp) || dalloc(&dol, C) || dalloc(&dop, C) ||
        dalloc(&dst, C) || dalloc(&dac, C) || dalloc(&dev, C) || dalloc(&dlj, C))
        return 1;
    CK(h2d_sync(dsx, sx.data(), sx.size() * sizeof(double)));
    CK(h2d_sync(dsl, sl.data(), C * sizeof(double)));
    CK(h2d_sync(dsp, sp.data(), C * sizeof(double)));
    CK(h2d_sync(dcov, cov, (size_t)D * D * sizeof(double)));
    CK(h2d_sync(dtape, tape, ntape * sizeof(double)));
    CK(h2d_sync(doff, tape_off, (C + 1) * sizeof(long long)));
    HMCParams p;
    memset(&p, 0, sizeof p);
    p.ens_n = 0; p.Dp = Dp; p.D = D;
    p.start_mode = START_GIVEN; p.start_x = dsx; p.start_logL = dsl; p.start_logP = dsp;
    p.lo = c->d_lo; p.hi = c->d_hi;
    p.cov = dcov; p.logdet = logdet;
    p.state = nullptr; p.use_override = 1; p.logLmin_override = logLmin; p.dt_override = dt; p.L_override = L;
    p.nmcmc_override = 1;
    p.max_traj = max_traj; p.j0 = 0; p.j1 = C; p.blob = c->d_blob;
    p.n_out = 1; p.out_x[0] = dox; p.out_logL[0] = dol; p.out_logP[0] = dop; p.out_steps[0] = dst; p.out_acc[0] = dac;
    p.out_evals[0] = dev;
    p.tape = dtape; p.tape_off = doff; p.traj = dtraj; p.logJ_out = dlj;
    int G, E;
    if (pick_config(D, lanes > 0 ? lanes : c->cfg.lanes, c->cfg.functor, C, &G, &E)) return 1;
    if (run_hmc(c, G, E, 1, p)) return 1;
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    std::vector<double> ox((size_t)C * Dp), ol(C), op(C);
    CK(cudaMemcpy(ox.data(), dox, ox.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ol.data(), dol, C * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(op.data(), dop, C * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < C; ++i) {
        for (int d = 0; d < D; ++d) out[(size_t)i * (D + 2) + d] = ox[(size_t)i * Dp + d];
        out[(size_t)i * (D + 2) + D] = ol[i];
        out[(size_t)i * (D + 2) + D + 1] = op[i];
    }
    if (steps) CK(cudaMemcpy(steps, dst, C * sizeof(int), cudaMemcpyDeviceToHost));
    if (accepted) CK(cudaMemcpy(accepted, dac, C * sizeof(int), cudaMemcpyDeviceToHost));
    if (evals) CK(cudaMemcpy(evals, dev, C * sizeof(int), cudaMemcpyDeviceToHost));
    if (log_J) CK(cudaMemcpy(log_J, dlj, C * sizeof(double), cudaMemcpyDeviceToHost));
    void* fr[] = {dsx, dsl, dsp, dcov, dtape, doff, dtraj, dox, dol, dop, dst, dac, dev, dlj};
    for (void* q : fr)
        if (q) cudaFree(q);
    return 0;
}

extern "C" int cpn_trace_hmc(cpn_ctx* c, int32_t C, const double* start, const double* cov, const double* eigval,
                             const double* eigvec, double dt, int32_t nmcmc, int32_t max_traj, double logLmin, int32_t lanes,
                             uint64_t chain_base, double* out, int32_t* steps, int32_t* accepted, int32_t* evals, double* trace,
                             int32_t trace_cap) {
    if (!c || !start || !cov || !eigval || !eigvec || !out || !trace) return fail("null argument");
    if (C < 1 || nmcmc < 1 || max_traj < 1 || trace_cap < c->D + 8) return fail("bad sizes");
    if (c->cfg.functor == CPN_USER) return fail("cpn_trace_hmc is not available for a user functor");
    CK(cudaSetDevice(c->cfg.device));
    const int D = c->D, Dp = c->Dp;
    std::vector<double> sx((size_t)C * Dp, 0.0), sl(C), sp(C), ev((size_t)D * Dp, 0.0), es(D, 0.0);
    for (int i = 0; i < C; ++i) {
        for (int d = 0; d < D; ++d) sx[(size_t)i * Dp + d] = start[(size_t)i * (D + 2) + d];
        sl[i] = start[(size_t)i * (D + 2) + D];
        sp[i] = start[(size_t)i * (D + 2) + D + 1];
    }
    for (int i = 0; i < D; ++i) {
        es[i] = sqrt(fabs(eigval[i]));
        for (int k = 0; k < D; ++k) ev[(size_t)i * Dp + k] = eigvec[(size_t)k * D + i];       // numpy layout: column i = eigenvector i
    }
    const size_t ntrace = (size_t)C * trace_cap;
    double *dsx = nullptr, *dsl = nullptr, *dsp = nullptr, *dcov = nullptr, *des = nullptr, *dvec = nullptr, *dox = nullptr, *dol = nullptr,
           *dop = nullptr, *dtr = nullptr;
    int *dst = nullptr, *dac = nullptr, *dev = nullptr;
    int rc = 1;
    do {
        if (dalloc(&dsx, sx.size()) || dalloc(&dsl, C) || dalloc(&dsp, C) || dalloc(&dcov, (size_t)D * D) || dalloc(&des, D) ||
            dalloc(&dvec, ev.size()) || dalloc(&dox, (size_t)C * Dp) || dalloc(&dol, C) || dalloc(&dop, C) || dalloc(&dst, C) ||
            dalloc(&dac, C) || dalloc(&dev, C) || dalloc(&dtr, ntrace))
            break;
        if (h2d_sync(dsx, sx.data(), sx.size() * sizeof(double)) != cudaSuccess || h2d_sync(dsl, sl.data(), C * sizeof(double)) != cudaSuccess ||
            h2d_sync(dsp, sp.data(), C * sizeof(double)) != cudaSuccess || h2d_sync(dcov, cov, (size_t)D * D * sizeof(double)) != cudaSuccess ||
            h2d_sync(des, es.data(), D * sizeof(double)) != cudaSuccess || h2d_sync(dvec, ev.data(), ev.size() * sizeof(double)) != cudaSuccess) {
            fail("cpn_trace_hmc: host to device copy failed");
            break;
        }
        HMCParams p;
        memset(&p, 0, sizeof p);
        p.ens_n = 0; p.Dp = Dp; p.D = D;
        p.start_mode = START_GIVEN; p.start_x = dsx; p.start_logL = dsl; p.start_logP = dsp;
        p.lo = c->d_lo; p.hi = c->d_hi;
        p.cov = dcov; p.eig_scale = des; p.eig_vec = dvec;
        p.state = nullptr; p.use_override = 1; p.logLmin_override = logLmin; p.dt_override = dt; p.L_override = 0;
        p.nmcmc_override = nmcmc; p.base_L = c->hmc_base_L;
        p.max_traj = max_traj; p.j0 = 0; p.j1 = C; p.blob = c->d_blob;
        p.chain_base = chain_base;
        p.seed_lo = (uint32_t)c->cfg.seed; p.seed_hi = (uint32_t)(c->cfg.seed >> 32);
        p.n_out = 1; p.out_x[0] = dox; p.out_logL[0] = dol; p.out_logP[0] = dop; p.out_steps[0] = dst; p.out_acc[0] = dac;
        p.out_evals[0] = dev;
        p.trace = dtr; p.trace_cap = trace_cap;
        int G, E;
        if (pick_config(D, lanes > 0 ? lanes : c->cfg.lanes, c->cfg.functor, C, &G, &E)) break;
        if (run_hmc(c, G, E, 0, p)) break;
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { fail("cpn_trace_hmc: launch failed"); break; }
        std::vector<double> ox((size_t)C * Dp), ol(C), op(C);
        if (cudaMemcpy(ox.data(), dox, ox.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(ol.data(), dol, C * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(op.data(), dop, C * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(trace, dtr, ntrace * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            (steps && cudaMemcpy(steps, dst, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (accepted && cudaMemcpy(accepted, dac, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) ||
            (evals && cudaMemcpy(evals, dev, C * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)) {
            fail("cpn_trace_hmc: device to host copy failed");
            break;
        }
        for (int i = 0; i < C; ++i) {
            for (int d = 0; d < D; ++d) out[(size_t)i * (D + 2) + d] = ox[(size_t)i * Dp + d];
            out[(size_t)i * (D + 2) + D] = ol[i];
            out[(size_t)i * (D + 2) + D + 1] = op[i];
        }
        rc = 0;
    } while (0);
    void* fr[] = {dsx, dsl, dsp, dcov, des, dvec, dox, dol, dop, dst, dac, dev, dtr};
    for (void* q : fr)
        if (q) cudaFree(q);
    return rc;
}

extern "C" int cpn_slice_directions(cpn_ctx* c, int32_t direction, int32_t C, double mu, double* out) {
    if (!c || !out) return fail("null argument");
    if (!c->live_ready) return fail("live set not initialised");
    if (direction < CPN_SLICE_DIFF || direction > CPN_SLICE_CORR || C < 1 || C > c->N) return fail("bad argument");
    CK(cudaSetDevice(c->cfg.device));
    if (adopt_pending_stats(c)) return 1;
    uint8_t* dcyc;
    double* ddir;
    char* scratch;
    if (dalloc(&dcyc, 4) || dalloc(&ddir, (size_t)C * c->D) || dalloc(&scratch, stage_bytes(c->N, c->Dp))) return 1;
    CK(cudaMemset(dcyc, direction, 4));
    CK(cudaStreamSynchronize(0));
    SliceParams p;
    fill_slice(c, &p);
    p.cycle = dcyc; p.cycle_len = 4; p.cycle_pos0 = 0;
    p.start_mode = START_SELF; p.j0 = 0; p.j1 = C;
    p.use_override = 1; p.logLmin_override = -INFINITY; p.nmcmc_override = 0; p.mu_override = mu;
    p.maxmcmc = 1;
    p.n_out = 1;
    p.out_x[0] = (double*)(scratch + stage_off(c->N, c->Dp, 0));
    p.out_logL[0] = (double*)(scratch + stage_off(c->N, c->Dp, 1));
    p.out_logP[0] = (double*)(scratch + stage_off(c->N, c->Dp, 2));
    p.out_steps[0] = (int*)(scratch + stage_off(c->N, c->Dp, 3));
    p.out_acc[0] = (int*)(scratch + stage_off(c->N, c->Dp, 4));
    p.out_start[0] = nullptr; p.out_evals[0] = nullptr; p.out_ne[0] = nullptr; p.out_nc[0] = nullptr;
    p.dir_out = ddir;
    if (slice_scratch_launch(c, p, 0, 0, C)) return 1;
    CK(cudaMemcpy(out, ddir, (size_t)C * c->D * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dcyc); cudaFree(ddir); cudaFree(scratch);
    return 0;
}

extern "C" int cpn_stage_replacements(cpn_ctx* c, int32_t K, const double* rows) {
    if (!c || !rows) return fail("null argument");
    if (K < 1 || K > c->N) return fail("K out of range");
    CK(cudaSetDevice(c->cfg.device));
    const size_t need = (size_t)K * (c->D + 2);
    CK(cudaMemcpyAsync(c->d_rec, rows, need * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int T = 256;
    k_unpack_records<<<(K * c->Dp + T - 1) / T, T, 0, c->stream>>>(c->d_rec, K, c->D, c->Dp, c->d_new_x, c->d_new_logL,
                                                                   c->d_new_logP);
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    c->staged_external = true;
    return 0;
}

extern "C" int cpn_ns_reset(cpn_ctx* c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->cfg.device));
    init_state(c, c->h_state);
    c->dead_upper = 0;
    return push_state(c);
}

extern "C" int cpn_ns_increment(cpn_ctx* c, const double* logL, int32_t K, int32_t nlive, int32_t ns_mode) {
    if (!c || !logL || K < 1 || nlive < K) return fail("bad argument");
    CK(cudaSetDevice(c->cfg.device));
    if (grow_logs(c, c->dead_upper + K + 1)) return 1;
    double* d;
    if (dalloc(&d, K)) return 1;
    CK(h2d_sync(d, logL, K * sizeof(double)));
    AccParams a;
    memset(&a, 0, sizeof a);
    a.st = c->d_state; a.skeys = d; a.N = nlive; a.K = K; a.mode = ns_mode; a.final_sweep = 1; a.nlive_run = c->N;
    a.hist_logL = c->d_hist_logL; a.hist_logvol = c->d_hist_logvol; a.update_threshold = 0;
    k_ns_accumulate<<<1, kAccThreads, 0, c->stream>>>(a);
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    c->dead_upper += K;
    return 0;
}

extern "C" int cpn_ns_trapezoid(cpn_ctx* c, double* logZ) {
    if (!c || !logZ) return fail("null argument");
    CK(cudaSetDevice(c->cfg.device));
    if (fetch_state(c)) return 1;
    k_trap_partial<<<256, 256, 0, c->stream>>>(c->d_hist_logL, c->d_hist_logvol, c->h_state->n_hist, c->d_trap_partial);
    k_trap_final<<<1, 32, 0, c->stream>>>(c->d_trap_partial, 256, c->d_scalar);
    CKL();
    CK(cudaMemcpyAsync(logZ, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

__global__ void k_philox_kat(const uint32_t* ctr, uint32_t k0, uint32_t k1, int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Philox4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], k0, k1);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = r.v[k];
}

extern "C" int cpn_philox(cpn_ctx* c, const uint32_t* ctr, const uint32_t* key, int32_t n, uint32_t* out) {
    if (!c || !ctr || !key || !out || n < 1) return fail("bad argument");
    CK(cudaSetDevice(c->cfg.device));
    uint32_t *dc, *dout;
    if (dalloc(&dc, (size_t)4 * n) || dalloc(&dout, (size_t)4 * n)) return 1;
    CK(h2d_sync(dc, ctr, (size_t)4 * n * sizeof(uint32_t)));
    k_philox_kat<<<(n + 127) / 128, 128, 0, c->stream>>>(dc, key[0], key[1], n, dout);
    CKL();
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, dout, (size_t)4 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    cudaFree(dc); cudaFree(dout);
    return 0;
}
