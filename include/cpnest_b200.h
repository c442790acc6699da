/* cpnest_b200 -- C-ABI of the B200-native nested-sampling hot path.
 *
 * Drop-in boundary.  The reference (johnveitch/cpnest) has no FFI:
 * its seams are the Python duck types `Model`, `Proposal`/`ProposalCycle` and the
 * sampler <-> NestedSampler pipe protocol.  Each entry point below cites the reference
 * interface (path:line in the reference tree) whose work it performs on the device.
 *
 * Conventions
 *   - every call returns 0 on success, non-zero on failure; cpn_last_error() gives the text.
 *   - one context per GPU; calls on one context are not thread-safe; contexts are independent.
 *   - all `double*` / `int*` arguments are HOST pointers, caller-owned, copied synchronously,
 *     unless the name says `dev_`.  Rows are row-major.  A "record" row is [x_0..x_{D-1}, logL,
 *     logP] (the column order of LivePoint.asnparray, cpnest/parameter.pyx:148-157).
 *   - no CPU fallback exists: without a CUDA device cpn_create fails.
 */
#ifndef CPNEST_B200_H
#define CPNEST_B200_H
#ifndef __CUDACC_RTC__
#include <stddef.h>
#include <stdint.h>
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpn_ctx cpn_ctx;

/* device functors for Model.log_prior / Model.log_likelihood (cpnest/model.py:55-82) */
enum cpn_functor {
    CPN_GAUSS_IID = 0,        /* examples/test_50d.py:18-19, tests/test_2d.py:15-16, tests/test_1d.py:20-21;
                                 blob: {mu, sigma} (2 doubles) */
    CPN_GAUSS_CORR = 1,       /* SURVEY 8d config 4, N(0,Sigma); blob: {logdet_L, Linv[D][D] row-major lower, LinvT[D][S] its transpose, rows zero-padded to
                                 S = D rounded up to 32} */
    CPN_EGGBOX = 2,           /* examples/eggbox.py:19-23 */
    CPN_ROSENBROCK = 3,       /* examples/rosenbrock.py:20-22 */
    CPN_ACKLEY = 4,           /* standard D-dim Ackley (examples/ackley.py:21-28 is non-finite everywhere) */
    CPN_RING = 5,             /* tests/test_ring.py:19-23; blob: {radius, thickness} */
    CPN_GAUSS_MEAN_SIGMA = 6, /* tests/test_gaussian.py:15-20 (1/sigma prior) */
    CPN_GAUSS_MIXTURE = 7,    /* examples/gaussianmixture.py:22-31; blob: {n_data, data[n_data]} */
    CPN_USER = 8,             /* user-written CUDA functor, compiled at run time: cpn_set_user_functor (any subclass of
                                 cpnest.model.Model with its own log_likelihood / log_prior, cpnest/model.py:55-82); blob: free */
    CPN_GAUSS_MIX_ND = 9,     /* two isotropic Gaussians in parameter space (BASELINE config 5's "20-D Gaussian mixture";
                                 not in the reference); blob: {w, m1, s1, m2, s2} */
    CPN_NUM_FUNCTORS = 10
};

/* proposal ids, order of DefaultProposalCycle (cpnest/proposal.py:354-361) */
enum cpn_proposal { CPN_WALK = 0, CPN_STRETCH = 1, CPN_DE = 2, CPN_EIGEN = 3 };

/* sampler choice: CPNest(nthreads - nslice - nhamiltonian, nslice, nhamiltonian) (cpnest/cpnest.py:197-261) */
enum cpn_sampler {
    CPN_SAMPLER_MH = 0,    /* MetropolisHastingsSampler + DefaultProposalCycle (cpnest/sampler.py:316-358) */
    CPN_SAMPLER_SLICE = 1, /* SliceSampler + EnsembleSliceProposalCycle (cpnest/sampler.py:414-569) */
    CPN_SAMPLER_HMC = 2    /* HamiltonianMonteCarloSampler + ConstrainedLeapFrog (cpnest/sampler.py:360-412) */
};

/* slice direction ids, order of EnsembleSliceProposalCycle (cpnest/proposal.py:196) */
enum cpn_slice_direction { CPN_SLICE_DIFF = 0, CPN_SLICE_GAUSS = 1, CPN_SLICE_CORR = 2 };

/* evidence accumulation rule for one batch of K dead points */
enum cpn_ns_mode {
    CPN_NS_SEQUENTIAL = 0, /* per-point shrinkage -1/(N-i), the reference's final-sweep rule
                              (cpnest/NestedSampling.py:474-476); default */
    CPN_NS_REFERENCE = 1   /* one increment, logt=-K/N at the K-th worst (NestedSampling.py:324-326) */
};

typedef struct cpn_config {
    int32_t device;        /* CUDA device ordinal */
    int32_t dim;           /* len(Model.names) */
    int32_t nlive;         /* CPNest(nlive=) */
    int32_t maxmcmc;       /* CPNest(maxmcmc=) */
    int32_t nmcmc_init;    /* initial chain length; <=0: maxmcmc/10 (cpnest/sampler.py:79) */
    int32_t functor;       /* enum cpn_functor */
    int32_t lanes;         /* lanes per chain G in {0=auto,1,4,8,16,32} */
    int32_t ns_mode;       /* enum cpn_ns_mode */
    uint64_t seed;         /* CPNest(seed=), Philox key */
    double nmcmc_safety;   /* safety factor of estimate_nmcmc_on_the_fly (cpnest/sampler.py:156); <=0: 5 */
    int32_t adapt_nmcmc;   /* 1: adapt chain length from the batch acceptance; 0: keep nmcmc_init */
    int32_t world_size;    /* number of ranks sharing the live set (1 = single GPU) */
    int32_t rank;
    int32_t sampler;       /* enum cpn_sampler */
    double rho_target;     /* chain-length control: start/end correlation to reach; <=0: 0.1 (MH, HMC), 0.05 (slice) */
    int32_t mix[3];        /* mixed sampler population, CPNest(nthreads, nslice, nhamiltonian) (cpnest/cpnest.py:197-261):
                              relative numbers of MH / slice / HMC chains in every batch (indexed by enum cpn_sampler);
                              the K chains of a batch are split in these proportions, MH chains first.  All zero: every
                              chain is of kind `sampler` */
    int32_t flags;         /* enum cpn_config_flags */
} cpn_config;

enum cpn_config_flags {
    CPN_NO_TWO_LAG = 1,    /* keep the single-lag chain-length control (start/end correlation only) even when it saturates */
    CPN_COLD_EIGEN = 2     /* eigen-decomposition of the ensemble covariance from the identity at every refresh (default: warm
                              start from the eigenvectors in use, a few Jacobi sweeps instead of ~9) */
};

/* NestedSampler / _NSintegralState observable state (cpnest/NestedSampling.py:88-144, 254-262) */
typedef struct cpn_state {
    double logZ;           /* NS.state.logZ */
    double info;           /* NS.state.info */
    double logw;           /* NS.state.logw */
    double logLmin;        /* manager.logLmin */
    double logLmax;        /* manager.logLmax */
    double condition;      /* NS.condition */
    double nmcmc_exact;    /* Sampler.Nmcmc_exact */
    double rho_max;        /* smoothed max_d |corr(chain start, chain end)| across the last batches */
    int64_t iteration;     /* NS.iteration = number of dead points */
    int64_t n_hist;        /* len(state.logLs) - 1 */
    int64_t batches;       /* consume_sample calls */
    int64_t total_steps;   /* sum of Sampler.mcmc_counter */
    int64_t total_accepted;/* sum of Sampler.mcmc_accepted */
    int64_t total_evals;   /* log_likelihood calls */
    int64_t failed_chains; /* chains that returned with no accepted step (NestedSampling.py:354-357) */
    int32_t nmcmc;         /* Sampler.Nmcmc */
    int32_t done;          /* condition <= tolerance reached */
    double slice_mu;       /* SliceSampler.mu */
    double hmc_dt;         /* HamiltonianProposal.dt */
    /* per sampler kind (enum cpn_sampler); nmcmc / nmcmc_exact / rho_max above are those of the context's primary kind */
    double rho_kind[3];
    int32_t nmcmc_kind[3];
    int32_t chains_kind[3];/* chains of each kind in the last batch */
    int32_t two_lag_kind[3];/* two-lag chain-length control: 0 off, else batches since it was switched on */
    double rho_decay_kind[3];/* its estimate of the decaying part of the start/end correlation (-1: none) */
} cpn_state;

const char* cpn_last_error(void);
int cpn_device_count(void);

/* CPNest.__init__ (cpnest/cpnest.py:103-261): one context = NestedSampler + its samplers. */
int cpn_create(const cpn_config* cfg, const double* lo, const double* hi,
               const void* functor_blob, size_t blob_bytes, cpn_ctx** out);
void cpn_destroy(cpn_ctx* ctx);

/* ProposalCycle.set_cycle (cpnest/proposal.py:71-75): the host draws the cycle (np.random.choice)
 * and hands the device the proposal id per slot. */
int cpn_set_cycle(cpn_ctx* ctx, const uint8_t* cycle, int32_t len);
/* EnsembleSliceProposalCycle (cpnest/proposal.py:189-207): direction id (enum cpn_slice_direction) per slot. */
int cpn_set_slice_cycle(cpn_ctx* ctx, const uint8_t* cycle, int32_t len);
/* SliceSampler.mu (cpnest/sampler.py:422): initial / current length scale of the slice directions. */
int cpn_set_slice_mu(cpn_ctx* ctx, double mu);

/* Device form of a user's Model subclass (cpnest/model.py:55-82): `source` is CUDA C++ defining the __device__
 * function `cpn_user_log_likelihood` (x, D, params) -> double and optionally, announced with
 * #define CPN_USER_HAS_PRIOR / CPN_USER_HAS_GRADIENT, `cpn_user_log_prior` and `cpn_user_grad_log_likelihood`
 * (exact prototypes: cpnest_b200/csrc/user_model.cuh).  Compiled with NVRTC for sm_100a into the same chain kernels
 * as the built-in functors; include_dirs must contain cpnest_b200/csrc and include/.  The context must have been
 * created with functor CPN_USER. */
int cpn_set_user_functor(cpn_ctx* ctx, const char* source, const char* const* include_dirs, int32_t n_dirs);
/* CPNest(seed=): Philox key of everything drawn from now on (takes effect with the next cpn_init_prior). */
int cpn_set_seed(cpn_ctx* ctx, uint64_t seed);
/* Sampler.reset + NestedSampler.reset (cpnest/sampler.py:110-154, NestedSampling.py:404-431):
 * draw nlive points uniformly in bounds (Model.new_point, model.py:35-53), evaluate, evolve each
 * with logLmin=-inf so they follow the prior, sort by logL. */
int cpn_init_prior(cpn_ctx* ctx);
/* Load a given live set instead (resume, tests, host protocol): rows[n][D+2] records. */
int cpn_set_live(cpn_ctx* ctx, const double* rows, int32_t n);

/* EnsembleEigenVector.update_eigenvectors (cpnest/proposal.py:306-323): mean, covariance (np.cov)
 * and its eigen-decomposition of the live ensemble, on the device. */
int cpn_update_ensemble_stats(cpn_ctx* ctx);
int cpn_get_ensemble_stats(cpn_ctx* ctx, double* mean, double* cov, double* eigval, double* eigvec);

/* NestedSampler.get_worst_n_live_points + _NSintegralState.increment + termination condition
 * (cpnest/NestedSampling.py:324-332, 368-375, 111-134).  Asynchronous. */
int cpn_select_and_accumulate(cpn_ctx* ctx, int32_t K);
/* MetropolisHastingsSampler / SliceSampler / HamiltonianMonteCarloSampler .yield_sample for the K chains of the batch
 * (cpnest/sampler.py:322-358, 465-569, 365-402).  With a mixed population (cpn_config.mix) the chains [0, K) are cut
 * into one contiguous range per kind and the kinds' kernels run concurrently on their own streams.  Asynchronous. */
int cpn_evolve_batch(cpn_ctx* ctx, int32_t K);
/* OrderedLivePoints.insert_live_point / remove_n_worst_points + insertion indices
 * (cpnest/NestedSampling.py:44-85, 343-366).  Asynchronous. */
int cpn_insert_batch(cpn_ctx* ctx, int32_t K);
/* ---- multi-GPU (one process per GPU; every rank holds a replica of the live set and evolves its
 * slice [K*rank/world, K*(rank+1)/world) of the K chains of a batch).  The exchange step between
 * cpn_evolve_batch and cpn_insert_batch is either
 *   fused : cpn_set_staging() with peer-mapped staging blocks - the chain kernel itself stores every
 *           finished chain into all ranks' blocks over NVLink; the caller only needs a barrier; or
 *   NCCL  : cpn_pack_shard() -> all-gather of [chains][cpn_shard_record_doubles()] doubles ->
 *           cpn_unpack_shards().
 * This replaces the per-sampler pipes of the reference (cpnest/cpnest.py:553-584). */
int64_t cpn_staging_bytes(cpn_ctx* ctx);   /* size of one peer block (two halves, alternating per batch) */
int cpn_set_staging(cpn_ctx* ctx, void* const* peer_blocks /*device addresses, [world]*/, int32_t npeers);
int64_t cpn_shard_record_doubles(cpn_ctx* ctx);
int cpn_shard_range(cpn_ctx* ctx, int32_t K, int32_t rank, int32_t* j0, int32_t* j1);
/* Which chains of a batch run which sampler kind (cpn_config.mix; pure host function, no context): kind s owns the chains
 * [kb[s], kb[s+1]) of [0, K), kb has 4 entries.  Largest-remainder split of K in the proportions mix[0] : mix[1] : mix[2],
 * at least one chain for every kind present - the device form of "each of the nthreads sampler processes gets one of the
 * worst points per iteration" (cpnest/NestedSampling.py:324-342, cpnest/cpnest.py:197-261).  Independent of the number of GPUs. */
int cpn_kind_ranges(const int32_t* mix, int32_t K, int32_t* kb);
int cpn_pack_shard(cpn_ctx* ctx, int32_t K, void* dev_send);
int cpn_unpack_shards(cpn_ctx* ctx, int32_t K, const void* dev_recv);
/* One consume_sample() = the three calls above (+ ensemble refresh on the sampler's cadence,
 * cpnest/sampler.py:252-254). */
int cpn_ns_step(cpn_ctx* ctx, int32_t K);
/* `while self.condition > self.tolerance: consume_sample()` (NestedSampling.py:451): runs batches
 * until the device-side condition drops below `tolerance` or max_batches are done; returns the
 * number of batches that did work in *batches_done. */
int cpn_run(cpn_ctx* ctx, int32_t K, int64_t max_batches, double tolerance, int64_t* batches_done);
/* final sweep over the remaining live points + trapezoid refinement
 * (NestedSampling.py:472-480, nest2pos.py:33-42). */
int cpn_finalise(cpn_ctx* ctx, double* logZ_rect, double* logZ_trap);

/* NestedSampler.checkpoint / resume + Sampler.checkpoint / resume (cpnest/NestedSampling.py:495-537,
 * cpnest/sampler.py:275-314): the whole run state (live set, order, dead-point log, integrator history, insertion
 * indices, ensemble statistics, controllers, RNG counters) as one host blob.  The context that loads it must have
 * been created with the same dim / nlive / functor / sampler / maxmcmc; the run then continues bit-identically. */
int64_t cpn_checkpoint_bytes(cpn_ctx* ctx);
int cpn_save_checkpoint(cpn_ctx* ctx, void* blob, int64_t bytes);
int cpn_load_checkpoint(cpn_ctx* ctx, const void* blob, int64_t bytes);

int cpn_synchronize(cpn_ctx* ctx);
/* Enqueue all further work on a caller-owned CUDA stream (cudaStream_t as void*), e.g. the stream
 * of the host framework so that its events and collectives order with the kernels.  NULL restores
 * the context's own stream. */
int cpn_set_stream(cpn_ctx* ctx, void* cuda_stream);
/* Sampler.samples / mcmc_chain_<pid>.dat (cpnest/sampler.py:86, 261-267, 345): the reference keeps the last 5*maxmcmc states
 * of every sampler's chains and writes them out at verbose >= 3.  cpn_enable_chain_log makes the MH kernel record the history
 * of the first n_chains chains of every batch (slot i = chain i of each batch in turn, like sampler process i) as events in a
 * ring of n_events rows per slot; cpn_get_chain_log returns a slot's events, oldest first, rows {tag, step, logL, logP, x[D]}:
 * tag 0 - this is the chain's state from `step` on (step -1: its start point), tag 1 - the chain ended after `step` steps.
 * The per-step history follows by repeating each state up to the next event (cpnest_b200/engine.py chain_history).
 * n_chains = 0 switches the log off. */
int cpn_enable_chain_log(cpn_ctx* ctx, int32_t n_chains, int32_t n_events);
int cpn_get_chain_log(cpn_ctx* ctx, int32_t slot, double* events /*[max_events][D+4]*/, int32_t max_events, int32_t* n_out);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t cpn_launch_count(cpn_ctx* ctx);
/* Lanes per chain (and coordinates per lane) a launch of `chains` chains uses when cpn_config.lanes == 0.  The likelihood sum is
 * reduced over the lanes of a chain, so its rounding - not the random stream - depends on this number: two runs agree bit
 * for bit only if they use the same one (a sharded run launches K / world_size chains per GPU, a single GPU all K). */
int cpn_pick_lanes(cpn_ctx* ctx, int32_t chains, int32_t* lanes, int32_t* per_lane);
int cpn_get_state(cpn_ctx* ctx, cpn_state* out);
int cpn_set_nmcmc(cpn_ctx* ctx, int32_t nmcmc);
/* NestedSampler(stopping=) (cpnest/NestedSampling.py:229,250) */
int cpn_set_tolerance(cpn_ctx* ctx, double tolerance);
int cpn_get_live(cpn_ctx* ctx, double* rows /*[nlive][D+2], ascending logL*/);
int64_t cpn_num_dead(cpn_ctx* ctx);
int cpn_get_dead(cpn_ctx* ctx, int64_t first, int64_t n, double* rows /*[n][D+2]*/, double* log_vols /*[n] or NULL*/);
int64_t cpn_num_insertion_indices(cpn_ctx* ctx);
int cpn_get_insertion_indices(cpn_ctx* ctx, int64_t first, int64_t n, double* out);
/* per-chain outcome of the last evolve: record rows, steps and accepted counts */
int cpn_get_last_batch(cpn_ctx* ctx, int32_t K, double* rows, int32_t* steps, int32_t* accepted);

/* The sampler <-> NestedSampler pipe protocol with HOST buffers (cpnest/sampler.py:233-246):
 * `requests[K][D+2]` are the points sent down the pipes (used as chain start points),
 * `replies[K][D+2]` the evolved points sent back.  Copies are inside the call. */
int cpn_evolve_host(cpn_ctx* ctx, int32_t K, double logLmin, int32_t nmcmc,
                    const double* requests, const int32_t* slots, double* replies, int32_t* steps,
                    int32_t* accepted);
/* `slots` (may be NULL): pool slot that reply j replaces afterwards, the device form of
 * `self.evolution_points.append(p)` (sampler.py:243, 351), so the proposal ensemble follows a
 * live set that the caller keeps in host memory. */
/* The same protocol without host-side row moves: the caller's live set is ONE pinned host table[nlive][D+2]
 * (cudaHostAlloc / cudaHostRegister memory, e.g. a torch tensor after .pin_memory()).  Request j is the row
 * table[start_idx[j]]; the device reads those rows from host memory itself, evolves them, and writes reply j into
 * row table[slots[j]] (and into pool slot slots[j] of the device ensemble).  start_idx / slots are host int32[K].
 * When the call returns the table holds the replies.  Same bytes over the bus as cpn_evolve_host, no CPU gather/scatter.
 * flags: CPN_TABLE_ROWS_ASYNC - return as soon as the logL / logP columns of the replaced rows (all the caller needs to pick the
 * next worst points, NestedSampling.py:368-375) and steps / accepted are in host memory; the coordinate columns of those rows
 * keep arriving while the caller works and are complete after the next call on this context or cpn_synchronize(). */
enum cpn_table_flags { CPN_TABLE_ROWS_ASYNC = 1 };
int cpn_evolve_host_table(cpn_ctx* ctx, int32_t K, double logLmin, int32_t nmcmc, double* table,
                          const int32_t* start_idx, const int32_t* slots, int32_t* steps, int32_t* accepted, int32_t flags);
/* Host-side row moves of a caller that keeps the live set in host memory as one table[nrows][row_doubles] (the
 * NestedSampler side of the protocol: `self.params[k]` handed to the pipes and replaced by the replies,
 * cpnest/NestedSampling.py:324-366): out[j] = table[idx[j]] and table[idx[j]] = rows[j].  Plain memcpy loops (numpy's
 * fancy indexing moves these rows at ~4 GB/s); pure host code, no context needed.  Indices out of range fail. */
int cpn_host_gather_rows(const double* table, int64_t nrows, int32_t row_doubles, const int64_t* idx, int32_t n, double* out);
int cpn_host_scatter_rows(double* table, int64_t nrows, int32_t row_doubles, const int64_t* idx, int32_t n, const double* rows);

/* ---- parity / test entry points ------------------------------------------------------------ */
/* Model.log_prior / log_likelihood on given points (functor parity vs the Python callables). */
int cpn_eval_model(cpn_ctx* ctx, const double* x /*[n][D]*/, int32_t n, double* logL, double* logP);
/* Gradient functors on given points: d log_likelihood / dx (the reflection normal of the constrained leapfrog,
 * cpnest/proposal.py:469-486, 766-772) and Model.force = d(-log_prior)/dx (cpnest/model.py:97-106). */
int cpn_eval_gradients(cpn_ctx* ctx, const double* x /*[n][D]*/, int32_t n, double* grad_logL /*[n][D]*/,
                       double* grad_V /*[n][D]*/);
/* MH chains driven by an explicit random tape instead of Philox (SURVEY 7 step 3 "replay mode").
 * ensemble[P][D]; start[C][D+2]; per chain c and step s the tape row tape[(c*maxmcmc+s)*8+..] is
 * {i0,i1,i2 (as doubles), g0,g1,g2, u_stretch, u_mh}.  compat=1 applies the reference's fp32
 * rounding of LivePoint scalars (cpnest/parameter.pyx:96-120). */
int cpn_replay_mh(cpn_ctx* ctx, int32_t P, const double* ensemble, int32_t C, const double* start,
                  const double* eigval, const double* eigvec /*[D][D] row k, col i as numpy*/,
                  int32_t cycle_idx, const double* tape, int32_t nmcmc, int32_t maxmcmc,
                  double logLmin, int32_t compat, int32_t lanes,
                  double* out /*[C][D+2]*/, int32_t* steps, int32_t* accepted,
                  double* states /*[C][maxmcmc][D] or NULL*/);
/* The PRODUCTION MH kernel - the instantiation cpn_evolve_batch launches, Philox draws and all - on C given chains, with a
 * record of what each step consumed besides the chain state: draws[(c*maxmcmc+s)*8+..] = {proposal kind, i0, i1, i2, c0, c1,
 * c2, thr}: the ensemble members (or the eigen-direction) the step gathered, its scalars (WALK: c_j = g_j - mean(g);
 * STRETCH: c0 = Z; DE: c0 = g; EIGEN: c0 = sqrt|lambda_i| N(0,1)) and thr = log(u) - log_J of the prior MH test
 * (cpnest/sampler.py:337; -1 where a flat prior makes the test depend on the box alone).  Feeding these to a restatement of
 * MetropolisHastingsSampler.yield_sample (cpnest/sampler.py:322-358) must reproduce the chain: the step-level parity test
 * of the timed code path.  Chain c uses the Philox stream of chain id chain_base + c; ensemble[P][D], start[C][D+2] as in
 * cpn_replay_mh; the cycle is the context's (cpn_set_cycle), entered at cycle_idx. */
int cpn_trace_mh(cpn_ctx* ctx, int32_t P, const double* ensemble, int32_t C, const double* start,
                 const double* eigval, const double* eigvec, int32_t cycle_idx, int32_t nmcmc, int32_t maxmcmc,
                 double logLmin, int32_t lanes, uint64_t chain_base, double* out /*[C][D+2]*/, int32_t* steps,
                 int32_t* accepted, int32_t* evals, double* draws /*[C][maxmcmc][8]*/);
/* Slice chains driven by an explicit random tape (SliceSampler.yield_sample, cpnest/sampler.py:465-569).
 * Per chain c the doubles tape[tape_off[c] .. tape_off[c+1]) hold, slice after slice,
 *   { u of reset_boundaries, direction data, Exp(1) draw, u of J, X'_1, X'_2, ... }
 * with direction data = {i0, i1} (EnsembleSliceDifferential: the two `sample`d members) or the D-vector numpy
 * returned (np.random.normal for EnsembleSliceGaussian, multivariate_normal for ...CorrelatedGaussian), and X'_k
 * the values np.random.uniform(L, R) returned.  cycle_idx indexes the slice cycle.  last[C][4] = {Ne, Nc, L, R}
 * of each chain's last slice. */
int cpn_replay_slice(cpn_ctx* ctx, int32_t P, const double* ensemble, int32_t C, const double* start,
                     int32_t cycle_idx, const double* tape, const int64_t* tape_off, double mu,
                     int32_t nmcmc, int32_t maxmcmc, double logLmin, int32_t compat, int32_t lanes,
                     double* out /*[C][D+2]*/, int32_t* steps, int32_t* accepted, int32_t* evals,
                     double* last /*[C][4]*/);
/* The PRODUCTION slice kernel (the instantiation every run launches: Philox draws, closed-form stepping out under a flat
 * prior, fused multiply-adds) on given chains, recording what every slice consumed besides the chain state, in the order of
 * the replay tape above: trace[c][trace_cap] = { n, then per slice: direction kind, nX, u of reset_boundaries, direction data
 * ({i0, i1} or the D-vector), Exp(1) draw, u of J, X'_1 .. X'_nX }, n = doubles written after the first (-1: trace_cap was
 * too small).  Feeding these to a restatement of SliceSampler.yield_sample (cpnest/sampler.py:465-569, with its stepping-out
 * loops :495-524) must reproduce the chain: the step-level parity test of the production path.  mean[D], eigval[D],
 * eigvec[D][D] (numpy layout) feed the correlated-Gaussian direction (cpnest/proposal.py:167-173), NULL: zeros.  Chain c uses
 * the Philox stream of chain id chain_base + c; the cycle is the context's (cpn_set_slice_cycle), entered at cycle_idx. */
int cpn_trace_slice(cpn_ctx* ctx, int32_t P, const double* ensemble, int32_t C, const double* start,
                    const double* mean, const double* eigval, const double* eigvec, int32_t cycle_idx, double mu,
                    int32_t nmcmc, int32_t maxmcmc, double logLmin, int32_t lanes, uint64_t chain_base,
                    double* out /*[C][D+2]*/, int32_t* steps, int32_t* accepted, int32_t* evals, double* last /*[C][4]*/,
                    double* trace /*[C][trace_cap]*/, int32_t trace_cap);
/* HMC chains driven by an explicit tape (HamiltonianMonteCarloSampler.yield_sample, cpnest/sampler.py:365-402, with
 * ConstrainedLeapFrog, cpnest/proposal.py:793-844).  Per chain c the doubles tape[tape_off[c] .. tape_off[c+1]) hold,
 * trajectory after trajectory, { p0[D] (momenta_distribution.rvs()), normal[D] for every reflection off the likelihood
 * boundary in the order they occur (unit_normal), u of np.random.choice in sample_trajectory, u of the acceptance test }.
 * cov[D][D] is the ensemble covariance (= inverse mass matrix), logdet = log det of the mass matrix. */
int cpn_replay_hmc(cpn_ctx* ctx, int32_t C, const double* start /*[C][D+2]*/, const double* cov, double logdet,
                   const double* tape, const int64_t* tape_off, double dt, int32_t L, int32_t max_traj,
                   double logLmin, int32_t lanes, double* out /*[C][D+2]*/, int32_t* steps, int32_t* accepted,
                   int32_t* evals, double* log_J /*[C]*/);
/* The PRODUCTION HMC kernel (the instantiation every run launches: Philox draws, momenta through the eigen-decomposition of
 * the covariance, analytic reflection normals, reflection in the metric of the kinetic energy, one-pass weighted draw of the
 * trajectory point, nmcmc trajectories per chain - hmc_kernel.cuh header) on given chains, recording what every trajectory
 * consumed besides the chain state: trace[c][trace_cap] = { n, then per trajectory: L, p0[D], the uniforms of the weighted
 * draw for the points 1 .. 2L-1 in the order they are visited, the uniform of the acceptance test }, n = doubles written
 * after the first (-1: trace_cap was too small).  A restatement of the kernel's trajectory on those inputs must reproduce the
 * chain: the step-level test of the production path (it differs from the reference's HamiltonianMonteCarloSampler by design,
 * so the restatement is of THIS algorithm, built on the reference's leapfrog, cpnest/proposal.py:714-835).
 * cov[D][D], eigval[D], eigvec[D][D] (numpy layout): the ensemble covariance and its eigen-decomposition. */
int cpn_trace_hmc(cpn_ctx* ctx, int32_t C, const double* start /*[C][D+2]*/, const double* cov, const double* eigval,
                  const double* eigvec, double dt, int32_t nmcmc, int32_t max_traj, double logLmin, int32_t lanes,
                  uint64_t chain_base, double* out /*[C][D+2]*/, int32_t* steps, int32_t* accepted, int32_t* evals,
                  double* trace /*[C][trace_cap]*/, int32_t trace_cap);
/* Directions of the production slice kernel for chains [0, C) of the next batch's first slice, all of kind
 * `direction` (test entry: distribution of the three direction proposals, cpnest/proposal.py:121-187). */
int cpn_slice_directions(cpn_ctx* ctx, int32_t direction, int32_t C, double mu, double* out /*[C][D]*/);
/* Put K given records into the replacement staging area, as if the chains had returned them
 * (`proposed` of NestedSampling.py:342): lets a test script the samplers' replies. */
int cpn_stage_replacements(cpn_ctx* ctx, int32_t K, const double* rows /*[K][D+2]*/);
/* _NSintegralState.increment on given sorted dead logL values (integrator parity). */
int cpn_ns_reset(cpn_ctx* ctx);
int cpn_ns_increment(cpn_ctx* ctx, const double* logL, int32_t K, int32_t nlive, int32_t ns_mode);
int cpn_ns_trapezoid(cpn_ctx* ctx, double* logZ);
/* Philox4x32-10 known-answer test: n counters (4 words each) -> n outputs (4 words each). */
int cpn_philox(cpn_ctx* ctx, const uint32_t* ctr, const uint32_t* key, int32_t n, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
