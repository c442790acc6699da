// Nested-sampling driver kernels: what the parent process of the reference does between two
// rounds of sampler requests (cpnest/NestedSampling.py):
//   k_ns_accumulate   get_worst_n_live_points (:368-375) + _NSintegralState.increment (:111-134)
//                     + termination condition (:332, :451) + Nmcmc moving average (sampler.py:156-176)
//   k_rank_bucket/scan/fill/final  ordering of the K replacements (KeyOrderedList.search, :42-46)
//   k_merge_commit    insert_live_point / remove_n_worst_points (:74-85) + insertion indices (:347-350)
//                     + nested_samples.extend (:327)
//   k_bitonic_step    initial sort of the live set (OrderedLivePoints.__init__, :71-72)
//   k_trap_*          _NSintegralState.finalise (:136-144) = nest2pos.log_integrate_log_trap
// The live set stays in HBM as rows live_x[N][Dp] + logL[N] + logP[N]; its order is a
// permutation `order` (ascending logL) with the sorted keys beside it, double-buffered.
#pragma once
#include <math.h>
#include <stdint.h>
#include "cpnest_b200.h"
#include "state.cuh"

namespace cpn {

constexpr int kAccThreads = 1024;

__device__ __forceinline__ double d_logaddexp(double a, double b) {
    // numpy logaddexp semantics
    if (a == b) return a + 0.69314718055994530942;
    double d = a - b;
    if (d > 0) return a + log1p(exp(-d));
    if (d <= 0) return b + log1p(exp(d));
    return a + b;  // NaN
}

// online log-sum-exp accumulator of (weight, weight*value) pairs in the log domain
struct LSE {
    double m, s, t;   // max, sum exp(w - m), sum exp(w - m) * v
    __device__ __forceinline__ void init() { m = -INFINITY; s = 0.0; t = 0.0; }
    __device__ __forceinline__ void push(double w, double v) {
        if (!(w > -INFINITY)) return;
        if (w > m) {
            double f = exp(m - w);
            s = s * f + 1.0;
            t = t * f + v;
            m = w;
        } else {
            double f = exp(w - m);
            s += f;
            t += f * v;
        }
    }
    __device__ __forceinline__ void merge(const LSE& o) {
        if (!(o.m > -INFINITY)) return;
        if (!(m > -INFINITY)) { *this = o; return; }
        if (o.m > m) {
            double f = exp(m - o.m);
            s = s * f + o.s;
            t = t * f + o.t;
            m = o.m;
        } else {
            double f = exp(o.m - m);
            s += o.s * f;
            t += o.t * f;
        }
    }
};

__device__ __forceinline__ LSE block_reduce_lse(LSE v, LSE* sh) {
    // fixed-shape tree: deterministic for a given block size
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            LSE a = sh[tid];
            a.merge(sh[tid + s]);
            sh[tid] = a;
        }
        __syncthreads();
    }
    LSE r = sh[0];
    __syncthreads();
    return r;
}

// exclusive block scan of one double per thread (Hillis-Steele in shared memory)
__device__ __forceinline__ double block_exclusive_scan(double v, double* sh, double* total) {
    const int tid = threadIdx.x, n = blockDim.x;
    sh[tid] = v;
    __syncthreads();
    for (int off = 1; off < n; off <<= 1) {
        double add = (tid >= off) ? sh[tid - off] : 0.0;
        __syncthreads();
        sh[tid] += add;
        __syncthreads();
    }
    double incl = sh[tid];
    *total = sh[n - 1];
    __syncthreads();
    return incl - v;
}

struct AccParams {
    NSState* st;
    const double* skeys;   // sorted live logL (ascending), or the given dead logL values
    int N;                 // live points enclosing the first dead point of this batch
    int K;                 // dead points of this batch
    int mode;              // cpn_ns_mode
    int final_sweep;       // 1: NestedSampling.py:472-476 (ignores done, no threshold update)
    int nlive_run;         // Nlive of the run (denominator of the termination condition)
    double* hist_logL;     // state.logLs[1:]
    double* hist_logvol;   // state.log_vols[1:]
    int update_threshold;  // 0 for the raw integrator parity entry
};

// What the chains of a batch need from the selection: whether the run goes on (`while self.condition > self.tolerance`,
// NestedSampling.py:451: a condition <= tolerance seen by the previous batch ends the run), the threshold = K-th worst
// logL (:374) and logLmax.  One thread; the evidence accumulation of the same batch (k_ns_accumulate) then runs on its
// own stream beside the chain kernel, which does not depend on it.
__global__ void k_ns_threshold(NSState* st, const double* __restrict__ skeys, int N, int K) {
    if (st->done) return;
    if (st->stop_next) { st->done = 1; return; }
    st->logLmin = skeys[K - 1];
    st->logLmax = skeys[N - 1];
}

__global__ void __launch_bounds__(kAccThreads) k_ns_accumulate(const AccParams p) {
    __shared__ LSE sh_lse[kAccThreads];
    __shared__ double sh_scan[kAccThreads];
    NSState* st = p.st;
    const int tid = threadIdx.x;
    if (!p.final_sweep && st->done) return;     // k_ns_threshold of this batch has turned stop_next into done
    const double logw0 = st->logw, logZ0 = st->logZ, B0 = st->B;
    const long long hist0 = st->n_hist;
    const int K = p.K, N = p.N;
    LSE acc;
    acc.init();
    double logw_end;
    if (p.mode == CPN_NS_REFERENCE) {
        // one increment at the K-th worst with logt = -K/N (NestedSampling.py:121, 324-326)
        const double logt = -(double)K / (double)N;
        const double logL = p.skeys[K - 1];
        const double Wt = logw0 + logL + log1p(-exp(logt));
        if (tid == 0) {
            const bool first = !(logZ0 > -INFINITY);
            acc.push(Wt, first ? Wt : logL);
            p.hist_logL[hist0] = logL;
            p.hist_logvol[hist0] = logw0 + logt;
        }
        logw_end = logw0 + logt;
    } else {
        // per-point shrinkage -1/(N-j) (the reference's final-sweep rule, NestedSampling.py:474-476).  Thread t owns
        // the consecutive points [t*C, (t+1)*C): a serial prefix inside the chunk, one block scan over the chunk totals
        const int C = (K + kAccThreads - 1) / kAccThreads;
        const int j0 = min(tid * C, K), j1 = min(j0 + C, K);
        double mine = 0.0;
        for (int j = j0; j < j1; ++j) mine += -1.0 / (double)(N - j);
        double total;
        const double before = block_exclusive_scan(mine, sh_scan, &total);
        double run = before;
        for (int j = j0; j < j1; ++j) {
            const double logt = -1.0 / (double)(N - j);
            const double logw_j = logw0 + run;
            const double logL = p.skeys[j];
            const double Wt = logw_j + logL + log1p(-exp(logt));
            // the reference leaves info at 0 on the increment that makes logZ finite, i.e. its
            // recurrence starts from info + logZ = Wt rather than logL (NestedSampling.py:125-128)
            bool first = false;
            if (!(logZ0 > -INFINITY) && Wt > -INFINITY) {
                if (j == 0) first = true;
                else first = !(p.skeys[j - 1] > -INFINITY);
            }
            acc.push(Wt, first ? Wt : logL);
            p.hist_logL[hist0 + j] = logL;
            p.hist_logvol[hist0 + j] = logw_j + logt;
            run += logt;
        }
        logw_end = logw0 + total;
    }
    const LSE tot = block_reduce_lse(acc, sh_lse);
    if (tid != 0) return;

    double logZ = logZ0, B = B0, info = st->info;
    if (tot.m > -INFINITY) {
        const double Wsum = tot.m + log(tot.s);
        logZ = d_logaddexp(logZ0, Wsum);
        const double fnew = exp(tot.m - logZ);
        B = fnew * tot.t + ((logZ0 > -INFINITY) ? exp(logZ0 - logZ) * B0 : 0.0);
        info = B - logZ;
        if (isnan(info)) { info = 0.0; B = logZ; }
    }
    st->logZ = logZ;
    st->B = B;
    st->info = info;
    st->logw = logw_end;
    st->n_hist = hist0 + ((p.mode == CPN_NS_REFERENCE) ? 1 : K);
    if (!p.update_threshold) { st->iteration += K; return; }

    const long long it0 = st->iteration;
    st->iteration = it0 + K;
    if (p.final_sweep) return;
    st->batches += 1;
    st->n_ins += K;                                                  // one insertion index per replacement
    const double logLmax = p.skeys[N - 1];
    const double cond = d_logaddexp(logZ, logLmax - (double)it0 / (double)p.nlive_run) - logZ;   // :332
    st->condition = cond;
    if (!(cond > st->tolerance)) st->stop_next = 1;
}

// Chain length control (replaces Sampler.estimate_nmcmc, sampler.py:178-198, whose FFT
// autocorrelation needs a serial chain history).  Two ingredients, evaluated once per batch:
//  (1) the reference's acceptance rule (sampler.py:156-176): ACL = 2/acc - 1, times `safety`;
//  (2) the measured dependence between where the K chains of the batch started and where they
//      ended: rho = max_d |corr_d(start, end)| over the coordinates and logL.  With geometric decay
//      rho = exp(-n/tau) the length that reaches rho_target is n * ln(rho_target) / ln(rho).
// The chain length follows the larger of the two with a moving average.
// One controller per sampler kind (KindCtl): `chains` chains of kind `kind` returned in this batch.
// WARP: called by the 32 lanes of one warp - the passes over the nrho correlations (two global loads per entry, otherwise a
// serial chain of L2 round trips: 17 us for 51 entries) are split over the lanes, the maxima meet by shuffles, lane 0 does the
// scalar rest.  The per-entry averages and the maxima do not depend on the order: same numbers as the one-thread form.
template <bool WARP = false>
__device__ __forceinline__ void adapt_nmcmc(NSState* st, int kind, int chains, const double* rho, double* rho2_ema, int nrho,
                                            const double* rho_mid = nullptr, double* lag_ema = nullptr, int allow_twolag = 0) {
    KindCtl& k = st->ctl[kind];
    const int lane = WARP ? ((int)threadIdx.x & 31) : 0, nl = WARP ? 32 : 1;
    const unsigned long long steps = k.batch[0], accd = k.batch[1];
    const bool first = k.rho_max < 0.0;
    const int twolag_in = k.twolag;
    if (WARP) __syncwarp();
    if (lane == 0)
        for (int i = 0; i < 4; ++i) { st->totals[i] += k.batch[i]; k.batch[i] = 0; }
    if (steps == 0 || chains <= 0) return;
    const double a = 0.2;
    double qmax = -1.0;
    if (rho != nullptr) {
        // q_d = EMA(rho_d^2 - 1/K): the sampling noise of a correlation over K chains is removed in
        // expectation and averaged down BEFORE the maximum over d is taken
        for (int d = lane; d < nrho; d += nl) {
            double r2 = rho[d] * rho[d] - 1.0 / (double)chains;
            if (!(r2 == r2)) r2 = 0.0;                       // zero variance in this coordinate
            const double q = first ? r2 : (1.0 - a) * rho2_ema[d] + a * r2;
            rho2_ema[d] = q;
            if (q > qmax) qmax = q;
        }
        if (WARP)
            for (int m = 16; m > 0; m >>= 1) qmax = fmax(qmax, __shfl_xor_sync(0xffffffffu, qmax, m));
        if (lane == 0) k.rho_max = sqrt(fmax(qmax, 0.0));
    }
    // Two-lag estimate (MH chains that leave a mid-chain snapshot, KindCtl::twolag).  A chain in a multimodal live set is
    // x_t = (centre of its mode) + z_t: the centre never changes, so corr(start, end) keeps a part no chain length removes.
    // The increments d1 = mid - start and d2 = end - mid do not contain the centre.  For a stationary z with autocovariance
    // g(h) = g(0) q^(h / (n/2)):   corr(d1, d2) = (2 g(n/2) - g(0) - g(n)) / (2 (g(0) - g(n/2))) = -(1 - q) / 2,
    // i.e. q = 1 + 2 corr(d1, d2) is the within-mode correlation over half a chain (0: forgotten, 1: frozen) and q^2 the part
    // of the start/end correlation that decays with the chain length: the target applies to it.
    double dmax = -1.0;
    if (rho != nullptr && rho_mid != nullptr && lag_ema != nullptr && twolag_in >= 2) {
        const bool first2 = twolag_in == 2;
        dmax = 0.0;
        for (int d = lane; d < nrho; d += nl) {
            double ci = rho_mid[d];                          // corr(d1, d2) of this batch
            if (!(ci == ci)) ci = -0.5;                      // a coordinate that does not move at all carries no information
            const double e1 = first2 ? ci : (1.0 - a) * lag_ema[d] + a * ci;
            lag_ema[d] = e1;
            const double q = fmin(fmax(1.0 + 2.0 * e1, 0.0), 1.0);
            if (q * q > dmax) dmax = q * q;
        }
        if (WARP)
            for (int m = 16; m > 0; m >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, m));
    }
    if (lane != 0) return;
    k.rho_decay = dmax;
    if (k.twolag > 0 && k.twolag < 1000) k.twolag += 1;
    if (!st->adapt) return;
    const double safety = st->nmcmc_safety;
    const double sub_acc = (double)accd / (double)steps;
    const double mean_len = (double)steps / (double)chains;
    double ne = k.nmcmc_exact;
    double target = (sub_acc > 0.0) ? safety * (2.0 / sub_acc - 1.0) : 2.0 * ne;      // sampler.py:171
    if (rho != nullptr && k.rho_target > 0.0) {
        // resolution of q_max: z * sigma_q, sigma_q = sqrt(2)/K * sqrt(a/(2-a)), z = sqrt(2 ln nrho)
        const double floor2 = 2.0 * sqrt(2.0 * log((double)nrho + 1.0)) * 1.41421356 / (double)chains * sqrt(a / (2.0 - a));
        const double tgt2 = fmax(k.rho_target * k.rho_target, floor2);
        const double r2 = fmin(fmax(dmax >= 0.0 ? dmax * dmax : qmax, 1e-6), 0.998);
        double t2 = mean_len * log(tgt2) / log(r2);
        t2 = fmin(fmax(t2, 0.5 * ne), 2.0 * ne);
        target = fmax(target, t2);
        // Switch to the two-lag estimate when lengthening the chains no longer lowers the correlation: every time the chain
        // length has grown by 1.6x since the reference batch, rho_max is compared with its value there; still above the
        // target and not even 20% lower means that what is left does not decay with the chain length.
        if (allow_twolag && k.twolag == 0) {
            if (!(k.ne_ref > 0.0) || ne < k.ne_ref) { k.ne_ref = ne; k.rho_ref = k.rho_max; }
            else if (ne >= 1.6 * k.ne_ref || ne >= 0.999 * (double)st->maxmcmc) {
                if (k.rho_max > 1.1 * k.rho_target && k.rho_max > 0.8 * k.rho_ref && ne > 1.2 * k.ne_ref) k.twolag = 1;
                else { k.ne_ref = ne; k.rho_ref = k.rho_max; }
            }
        }
    }
    ne = (1.0 - a) * ne + a * target;
    ne = fmin(ne, (double)st->maxmcmc);
    k.nmcmc_exact = ne;
    int n = (int)ne;
    if (n < (int)safety) n = (int)safety;
    if (n > st->maxmcmc) n = st->maxmcmc;
    k.nmcmc = n;
}
__global__ void k_adapt_nmcmc(NSState* st, int kind, int chains, const double* rho, double* rho2_ema, int nrho) {
    if (st->done) return;
    adapt_nmcmc(st, kind, chains, rho, rho2_ema, nrho);
}

// Sums of the per-chain outcome arrays of the chains [j0, j1) of a batch (those of sampler kind `kind`)
// -> st->ctl[kind].batch[] = {steps, accepted, evals, failed chains}.  Integer sums: the result does not depend on the order, nor on how the chains were split
// over GPUs.
constexpr int kCorrThreads = 1024;    // block size of k_chain_corr / k_batch_counters (fixed: the reduction shape must not vary)
__device__ __forceinline__ void block_batch_counters(NSState* st, int kind, int j0, int j1, const int* steps, const int* acc,
                                                     const int* evals, unsigned long long (*shc)[kCorrThreads]) {
    const int tid = threadIdx.x;
    unsigned long long a[4] = {0, 0, 0, 0};
    for (int j = j0 + tid; j < j1; j += kCorrThreads) {
        a[0] += (unsigned long long)steps[j];
        a[1] += (unsigned long long)acc[j];
        a[2] += (unsigned long long)(evals != nullptr ? evals[j] : 0);
        a[3] += (acc[j] == 0) ? 1ull : 0ull;
    }
    for (int k = 0; k < 4; ++k) shc[k][tid] = a[k];
    __syncthreads();
    for (int w = kCorrThreads / 2; w > 0; w >>= 1) {
        if (tid < w)
            for (int k = 0; k < 4; ++k) shc[k][tid] += shc[k][tid + w];
        __syncthreads();
    }
    if (tid == 0)
        for (int k = 0; k < 4; ++k) st->ctl[kind].batch[k] = shc[k][0];
}
__global__ void __launch_bounds__(kCorrThreads) k_batch_counters(NSState* st, int kind, int j0, int j1, const int* steps,
                                                                 const int* acc, const int* evals) {
    __shared__ unsigned long long shc[4][kCorrThreads];
    block_batch_counters(st, kind, j0, j1, steps, acc, evals, shc);
}

// Sampler.estimate_nmcmc_on_the_fly (cpnest/sampler.py:156-176), the rule the reference applies after every chain of
// its burn-in, here for the chains [j0, j1) of a batch in chain order, each with its own sub-acceptance a_j = accepted_j /
// steps_j:   Nmcmc_exact <- min(maxmcmc, (1 - 1/tau) Nmcmc_exact + (safety/tau) (2/a_j - 1))      (a_j > 0)
//            Nmcmc_exact <- min(maxmcmc, (1 + 1/tau) Nmcmc_exact)                                 (a_j = 0)
// Every update is a map x -> min(m x + b, c) with m > 0, and such maps compose into one of the same form,
//   (m2, b2, c2) o (m1, b1, c1) = (m2 m1, m2 b1 + b2, min(m2 c1 + b2, c2)),
// so the K sequential updates are one ordered parallel reduction: each thread composes a run of consecutive chains, a
// fixed-shape tree composes the runs in order.  Used where no start/end correlation is available (host protocol); the
// result equals the sequential rule up to the rounding of the re-associated products (tests: 1e-12).
struct ClampAffine {
    double m, b, c;
    __device__ __forceinline__ void then(const ClampAffine& g) {      // this := g o this
        const double nb = fma(g.m, b, g.b);
        c = fmin(fma(g.m, c, g.b), g.c);
        m = g.m * m;
        b = nb;
    }
};
__global__ void __launch_bounds__(kCorrThreads) k_adapt_nmcmc_on_the_fly(NSState* st, int kind, int j0, int j1, const int* steps,
                                                                         const int* acc) {
    if (st->done) return;
    __shared__ ClampAffine sh[kCorrThreads];
    const int tid = threadIdx.x, n = j1 - j0;
    const double tau = st->nmcmc_tau, safety = st->nmcmc_safety, M = (double)st->maxmcmc;
    const int per = (n + kCorrThreads - 1) / kCorrThreads;
    ClampAffine f = {1.0, 0.0, INFINITY};
    for (int j = j0 + tid * per; j < min(j1, j0 + (tid + 1) * per); ++j) {
        ClampAffine u;
        if (acc[j] == 0) u = {1.0 + 1.0 / tau, 0.0, M};
        else u = {1.0 - 1.0 / tau, (safety / tau) * (2.0 / ((double)acc[j] / (double)steps[j]) - 1.0), M};
        f.then(u);
    }
    sh[tid] = f;
    __syncthreads();
    for (int w = 1; w < kCorrThreads; w <<= 1) {          // ordered tree: sh[i] := sh[i + w] o sh[i]
        if ((tid & (2 * w - 1)) == 0) {
            ClampAffine a = sh[tid];
            a.then(sh[tid + w]);
            sh[tid] = a;
        }
        __syncthreads();
    }
    if (tid == 0 && n > 0 && st->adapt) {
        KindCtl& k = st->ctl[kind];
        const ClampAffine t = sh[0];
        const double ne = fmin(fma(t.m, k.nmcmc_exact, t.b), t.c);
        k.nmcmc_exact = ne;
        int nn = (int)ne;
        if (nn < (int)safety) nn = (int)safety;                      // max(safety, int(Nmcmc_exact))
        k.nmcmc = nn;
    }
}

// SliceSampler.adapt_length_scale (sampler.py:429-437): mu *= 2 Ne / (Ne + Nc).  The reference applies it per
// chain with the counts of that chain's last slice, during its first 10*poolsize slices; here the counts of
// ALL slices of the batch are pooled (integer sums: independent of order and of the split over GPUs), which
// makes the estimate of the expansion/contraction balance precise enough to keep following the shrinking
// constrained region for the whole run.
__global__ void __launch_bounds__(256) k_adapt_slice_mu(NSState* st, int j0, int j1, const int* ne, const int* nc) {
    if (st->done) return;
    __shared__ unsigned long long sh[2][256];
    const int tid = threadIdx.x;
    unsigned long long a = 0, b = 0;
    for (int j = j0 + tid; j < j1; j += 256) { a += (unsigned long long)ne[j]; b += (unsigned long long)nc[j]; }
    sh[0][tid] = a; sh[1][tid] = b;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) { sh[0][tid] += sh[0][tid + w]; sh[1][tid] += sh[1][tid + w]; }
        __syncthreads();
    }
    if (tid == 0) {
        const double Ne = fmax(1.0, (double)sh[0][0]), Nc = fmax(1.0, (double)sh[1][0]);
        st->slice_mu *= 2.0 * (Ne / (Ne + Nc));
    }
}

// HamiltonianProposal.update_time_step (proposal.py:539-550): log(scale) += ADAPTATIONSIZE * (acceptance - TARGET),
// TARGET = 0.5.  The reference applies it once per returned chain with ADAPTATIONSIZE = 0.001 and its running
// acceptance; here once per batch with the batch's trajectory acceptance (accepted / attempted trajectories) and a
// gain of 0.05.  Must run after k_chain_corr (batch counters) and before k_merge_commit (which folds them).
__global__ void k_adapt_hmc_dt(NSState* st) {
    if (st->done) return;
    const unsigned long long traj = st->ctl[CPN_SAMPLER_HMC].batch[0], accd = st->ctl[CPN_SAMPLER_HMC].batch[1];
    if (traj == 0) return;
    const double acc = (double)accd / (double)traj;
    st->hmc_dt *= exp(0.05 * (acc - 0.5));
}

// corr_d(start, end) across the chains [j0, j1) of the batch (those of sampler kind `kind`), d in [0, D] (d == D is
// logL).  Two kernels: k_chain_corr_partial - block (d, s) sums the five moments of coordinate d over its segment of
// kCorrSeg chains (fixed-shape tree: deterministic), block row D+1 sums the batch counters; k_chain_corr_final adds the
// segments in order and writes rho[d] and st->ctl[kind].batch[].  The grid grows with the batch, so the replicated
// statistics of a multi-GPU batch (K = all ranks' chains) do not serialise on D+2 blocks.
constexpr int kCorrSeg = 512;          // chains per block of k_chain_corr_partial: 16 per warp (K = 37632: 74 segments x 3 block rows)
constexpr int kCorrMom = 9;      // s, e, ss, ee, se and - two-lag control - m, mm, sm, me (m: the mid-chain snapshot)
__global__ void __launch_bounds__(kCorrThreads) k_chain_corr_partial(const NSState* st, int kind, int j0, int j1, int D, int Dp,
                                                                     const int* __restrict__ start_slot, const double* __restrict__ live_x,
                                                                     const double* __restrict__ live_logL, const double* __restrict__ new_x,
                                                                     const double* __restrict__ new_logL, const double* __restrict__ mid_x,
                                                                     const double* __restrict__ mid_logL, const int* steps, const int* acc,
                                                                     const int* evals, double* __restrict__ part /*[D+2][nseg][kCorrMom]*/) {
    // grid (ceil((D + 1) / 32) + 1, nseg).  Block (ds, sg), ds below the last: the lanes of a warp are the 32 coordinates
    // d = 32 ds + lane (d == D: logL), the 32 warps stride over the segment's chains - a chain's rows are read with
    // consecutive lanes on consecutive doubles (one thread per chain and block per coordinate, as before, fetched a 32-byte
    // sector per 8 useful bytes, and every row D + 1 times: 32 us at K = 37632).  The last block row sums the batch counters.
    if (st->done) return;
    __shared__ double sh[32][33];
    const int sg = blockIdx.y, nseg = gridDim.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a0 = j0 + sg * kCorrSeg, a1 = min(j1, a0 + kCorrSeg);
    const bool mid = mid_x != nullptr && st->ctl[kind].twolag >= 2;       // this batch's chains left snapshots
    double a[kCorrMom] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool counters = (int)blockIdx.x == (int)gridDim.x - 1;
    const int d = counters ? D + 1 : (int)blockIdx.x * 32 + lane;
    if (counters) {
        // counters as doubles: exact up to 2^53
        for (int j = a0 + tid; j < a1; j += kCorrThreads) {
            a[0] += (double)steps[j];
            a[1] += (double)acc[j];
            a[2] += (double)(evals != nullptr ? evals[j] : 0);
            a[3] += (acc[j] == 0) ? 1.0 : 0.0;
        }
        // lanes hold different chains here: butterflies inside the warp first
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], m);
    } else if (d <= D) {
        // shift by the first chain's values to keep the sums well conditioned
        const int s0 = start_slot[j0];
        const double cs = (d < D) ? live_x[(size_t)s0 * Dp + d] : live_logL[s0];
        const double ce = (d < D) ? new_x[(size_t)j0 * Dp + d] : new_logL[j0];
#pragma unroll 4
        for (int j = a0 + warp; j < a1; j += 32) {
            const int sl = start_slot[j];
            const double s = ((d < D) ? live_x[(size_t)sl * Dp + d] : live_logL[sl]) - cs;
            const double e = ((d < D) ? new_x[(size_t)j * Dp + d] : new_logL[j]) - ce;
            a[0] += s; a[1] += e; a[2] += s * s; a[3] += e * e; a[4] += s * e;
            if (mid) {
                const double m = ((d < D) ? mid_x[(size_t)j * Dp + d] : mid_logL[j]) - ce;
                a[5] += m; a[6] += m * m; a[7] += s * m; a[8] += m * e;
            }
        }
    }
    // fixed-shape reduction over the 32 warps, one moment at a time: lane l of warp 0 adds the 32 warp values of coordinate
    // 32 ds + l in order (counters: every lane holds its warp's total, lane 0's column is used)
#pragma unroll
    for (int k = 0; k < kCorrMom; ++k) {
        sh[warp][lane] = a[k];
        __syncthreads();
        if (warp == 0) {
            double v = 0.0;
#pragma unroll 8
            for (int w = 0; w < 32; ++w) v += sh[w][lane];
            if (counters ? (lane == 0 && k < 4) : (d <= D)) part[((size_t)d * nseg + sg) * kCorrMom + k] = v;
        }
        __syncthreads();
    }
}
// one warp per coordinate (block d): lane l adds the segments l, l + 32, ... in order, the 32 lane sums meet in a butterfly -
// a fixed shape for a given number of segments, so the result is reproducible
__global__ void __launch_bounds__(32) k_chain_corr_final(NSState* st, int kind, int n, int D, int nseg, const double* __restrict__ part,
                                                         double* __restrict__ rho, double* __restrict__ rho_mid) {
    if (st->done) return;
    const int d = blockIdx.x, lane = threadIdx.x;
    if (d > D + 1) return;
    double a[kCorrMom] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int sg = lane; sg < nseg; sg += 32)
#pragma unroll
        for (int k = 0; k < kCorrMom; ++k) a[k] += part[((size_t)d * nseg + sg) * kCorrMom + k];
#pragma unroll
    for (int k = 0; k < kCorrMom; ++k)
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], m);
    if (lane != 0) return;
    if (d == D + 1) {
        for (int k = 0; k < 4; ++k) st->ctl[kind].batch[k] = (unsigned long long)a[k];
        return;
    }
    const double nn = (double)n;
    const double vs = a[2] - a[0] * a[0] / nn, ve = a[3] - a[1] * a[1] / nn;
    const double cv = a[4] - a[0] * a[1] / nn;
    rho[d] = cv / sqrt(vs * ve);
    if (rho_mid != nullptr) {
        // correlation of the increments d1 = mid - start, d2 = end - mid from the second moments of (s, m, e)
        const double vm = a[6] - a[5] * a[5] / nn, csm = a[7] - a[0] * a[5] / nn, cme = a[8] - a[5] * a[1] / nn;
        const double v1 = vm - 2.0 * csm + vs, v2 = ve - 2.0 * cme + vm;
        const double c12 = cme - vm - cv + csm;
        rho_mid[d] = (st->ctl[kind].twolag >= 2) ? c12 / sqrt(v1 * v2) : __longlong_as_double(0x7ff8000000000000LL);
    }
}

// Rank of each replacement among the K replacements (ties broken by chain index), and the sorted copy of their keys.
// The sorted survivors are the pivots: replacement q falls into bucket nle_q = #(survivors <= key_q) (bisect-right, the
// position `KeyOrderedList.search` returns, NestedSampling.py:42-46, which the merge needs anyway).  Replacements of
// different buckets are ordered by bucket; a bucket holds K / (N - K) replacements on average, so the order inside it is
// found by comparing its few members with each other:
//   k_rank_bucket  nle_q by binary search, histogram of the buckets (integer atomics: order independent)
//   k_rank_scan_*  exclusive prefix sum of the histogram (tiles, then tiles in front + scan of the tile)
//   k_rank_fill    the members of every bucket, listed in arbitrary order
//   k_rank_final   rank_q = start of the bucket + #(members with (key, chain) < (key_q, q)): does not depend on the
//                  order in which k_rank_fill listed them, so the result is deterministic
// Work O(K log N) instead of the K^2 comparisons of a counting rank; every step of it is replicated on all ranks of a
// multi-GPU run, where K is the whole batch.
__global__ void k_rank_bucket(const NSState* st, const double* __restrict__ new_logL, int K, const double* __restrict__ surv, int ns,
                              int* __restrict__ nle, int* __restrict__ hist) {
    if (st->done) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= K) return;
    const double key = new_logL[q];
    int lo = 0, hi = ns;                      // #(surv <= key)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (surv[mid] <= key) lo = mid + 1; else hi = mid;
    }
    nle[q] = lo;
    atomicAdd(hist + lo, 1);
}
// exclusive prefix sum of the bucket histogram in two launches with coalesced accesses: k_rank_scan_tiles sums tiles of
// kScanTile buckets; k_rank_scan_apply adds up the tiles in front of its own and scans its tile (hist[b] becomes the start
// of bucket b in the sorted replacements)
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;
__device__ __forceinline__ int block_sum_int(int v, int* sh) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((tid & 31) == 0) sh[tid >> 5] = v;
    __syncthreads();
    int t = (tid < kScanThreads / 32) ? sh[tid] : 0;
    if (tid < 32) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (tid == 0) sh[0] = t;
    }
    __syncthreads();
    const int r = sh[0];
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(kScanThreads) k_rank_scan_tiles(const NSState* st, const int* __restrict__ hist, int nb, int* __restrict__ tile_sum) {
    if (st->done) return;
    __shared__ int sh[32];
    const int base = blockIdx.x * kScanTile;
    int v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int b = base + i * kScanThreads + threadIdx.x;
        v += (b < nb) ? hist[b] : 0;
    }
    v = block_sum_int(v, sh);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = v;
}
__global__ void __launch_bounds__(kScanThreads) k_rank_scan_apply(const NSState* st, int* __restrict__ hist, int nb, const int* __restrict__ tile_sum) {
    if (st->done) return;
    __shared__ int sh[32];
    __shared__ int wsum[kScanThreads / 32];
    const int tid = threadIdx.x, base = blockIdx.x * kScanTile;
    // buckets in front of this tile
    int pre = 0;
    for (int t = tid; t < (int)blockIdx.x; t += kScanThreads) pre += tile_sum[t];
    pre = block_sum_int(pre, sh);
    // this thread owns kScanItems consecutive buckets of the tile
    int h[kScanItems], mine = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int b = base + tid * kScanItems + i;
        h[i] = (b < nb) ? hist[b] : 0;
        mine += h[i];
    }
    // exclusive scan of `mine` over the block: shuffles inside a warp, warp totals through shared memory
    int incl = mine;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, m);
        if ((tid & 31) >= m) incl += o;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = (tid < kScanThreads / 32) ? wsum[tid] : 0, wi = w;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, wi, m);
            if (tid >= m) wi += o;
        }
        if (tid < kScanThreads / 32) wsum[tid] = wi - w;          // exclusive
    }
    __syncthreads();
    int run = pre + wsum[tid >> 5] + incl - mine;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int b = base + tid * kScanItems + i;
        if (b < nb) hist[b] = run;
        run += h[i];
    }
}
__global__ void k_rank_fill(const NSState* st, int K, const int* __restrict__ nle, const int* __restrict__ start, int* __restrict__ cursor,
                            int* __restrict__ members) {
    if (st->done) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= K) return;
    const int b = nle[q];
    members[start[b] + atomicAdd(cursor + b, 1)] = q;
}
__global__ void k_rank_final(const NSState* st, const double* __restrict__ new_logL, int K, const int* __restrict__ nle,
                             const int* __restrict__ start, const int* __restrict__ cursor, const int* __restrict__ members,
                             int* __restrict__ rank, double* __restrict__ snew) {
    if (st->done) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= K) return;
    const int b = nle[q];
    const int s0 = start[b], n = cursor[b];
    const double key = new_logL[q];
    int r = s0;
    for (int t = 0; t < n; ++t) {
        const int o = members[s0 + t];
        const double ko = new_logL[o];
        r += (ko < key) || (ko == key && o < q);
    }
    rank[q] = r;
    snew[r] = key;
}

struct MergeParams {
    const NSState* st;
    NSState* st_w;
    int N, K, Dp, Dreal;
    int dshift;                // smallest shift with (1 << dshift) >= Dp
    const int* order_in;       // [N]
    const double* skeys_in;    // [N]
    int* order_out;
    double* skeys_out;
    const int* rank_new;       // [K]
    const int* nle;            // [K] #(survivors <= key) of every replacement (k_rank_bucket)
    const double* snew;        // [K] sorted replacement keys
    double* live_x;            // [N][Dp]
    double* live_logL;
    double* live_logP;
    const double* new_x;       // [K][Dp]
    const double* new_logL;
    const double* new_logP;
    const double* rho;         // [kinds][D+1] start/end correlations of this batch (k_chain_corr), per sampler kind
    double* rho2_ema;          // [kinds][D+1] smoothed, de-biased squares
    const double* rho_mid;     // [kinds][D+1] correlation of the two half-chain increments (two-lag control; NaN when not recorded)
    double* lag_ema;           // [kinds][D+1] (x2 reserved) its moving average
    int allow_twolag;          // the exchange carries the mid-chain snapshots (always on one GPU; fused exchange on several)
    int kind_n[kNumKinds];     // chains of each sampler kind in this batch
    double* dead_x;            // rows appended at dead0
    double* dead_logL;
    double* dead_logP;
    double* insertion;         // one entry per replacement, appended
};

__device__ __forceinline__ int lower_bound_d(const double* a, int n, double v) {   // #elements < v
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound_d(const double* a, int n, double v) {   // #elements <= v
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One thread per live-set position.  Threads [0,K) own a dead point + its replacement,
// threads [K,N) own a survivor.  bisect-right semantics: a replacement goes after equal keys.
// The rows themselves (nested_samples.extend(worst), then the slot is overwritten by the replacement) are moved
// by the threads beyond N, one coordinate each, so that consecutive threads touch consecutive doubles.
__global__ void k_merge_commit(const MergeParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = p.N, K = p.K;
    if (p.st->done) {              // run finished: keep the order valid in the other buffer
        if (i < N) {
            p.order_out[i] = p.order_in[i];
            p.skeys_out[i] = p.skeys_in[i];
        }
        return;
    }
    // k_ns_accumulate of this batch already advanced the counters
    const long long dead0 = p.st->iteration - K, ins0 = p.st->n_ins - K;
    if (i >= N) {
        // row r, coordinate d of the replacements: 2^dshift >= Dp threads per row (no division per element)
        const int w = i - N;
        const int r = w >> p.dshift, d = w & ((1 << p.dshift) - 1);
        if (r >= K || d >= p.Dp) return;
        const int slot = p.order_in[r];
        double* l = p.live_x + (size_t)slot * p.Dp + d;
        p.dead_x[(size_t)(dead0 + r) * p.Dp + d] = *l;
        *l = p.new_x[(size_t)r * p.Dp + d];
        return;
    }
    if (i >= K) {
        const double key = p.skeys_in[i];
        const int pos = (i - K) + lower_bound_d(p.snew, K, key);
        p.order_out[pos] = p.order_in[i];
        p.skeys_out[pos] = key;
    } else {
        const int slot = p.order_in[i];
        const double key = p.new_logL[i];
        const int nle = p.nle[i];
        const int pos = p.rank_new[i] + nle;
        p.order_out[pos] = slot;
        p.skeys_out[pos] = key;
        p.insertion[ins0 + i] = (double)nle / (double)(N - K);
        p.dead_logL[dead0 + i] = p.live_logL[slot];
        p.dead_logP[dead0 + i] = p.live_logP[slot];
        p.live_logL[slot] = key;
        p.live_logP[slot] = p.new_logP[i];
    }
}

// The chain-length controllers of the batch's sampler kinds (adapt_nmcmc: a pass over the D + 1 correlations of a kind).  Its own launch on the stream of the chain statistics, beside the ranking kernels: in the merge kernel it was one
// thread's serial work on the main stream's critical path.
__global__ void k_adapt_batch(const MergeParams p) {      // one warp
    if (p.st->done) return;
    const int nrho = p.Dreal + 1;
    for (int kind = 0; kind < kNumKinds; ++kind)
        if (p.kind_n[kind] > 0)
            adapt_nmcmc<true>(p.st_w, kind, p.kind_n[kind], p.rho != nullptr ? p.rho + kind * nrho : nullptr,
                        p.rho2_ema + kind * nrho, p.rho != nullptr ? nrho : 0,
                        (p.rho_mid != nullptr && kind == CPN_SAMPLER_MH) ? p.rho_mid + kind * nrho : nullptr,
                        p.lag_ema + 2 * kind * nrho, kind == CPN_SAMPLER_MH ? p.allow_twolag : 0);
}

// copy the (sorted) remaining live points into the dead log: the final sweep
__global__ void k_append_live_sorted(const int* order, int N, int Dp, const double* live_x, const double* live_logL,
                                     const double* live_logP, double* dead_x, double* dead_logL, double* dead_logP,
                                     long long dead0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int slot = order[i];
    for (int d = 0; d < Dp; ++d) dead_x[(size_t)(dead0 + i) * Dp + d] = live_x[(size_t)slot * Dp + d];
    dead_logL[dead0 + i] = live_logL[slot];
    dead_logP[dead0 + i] = live_logP[slot];
}

// ---- full sort of the live set (init only): bitonic network on (key, slot) pairs ------------------
__global__ void k_sort_fill(const double* logL, int N, int Npad, double* keys, int* idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    keys[i] = (i < N) ? logL[i] : INFINITY;
    idx[i] = (i < N) ? i : 0x7fffffff;
}
__global__ void k_bitonic_step(double* keys, int* idx, int Npad, int j, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    const int l = i ^ j;
    if (l > i) {
        const double a = keys[i], b = keys[l];
        const int ia = idx[i], ib = idx[l];
        const bool up = (i & k) == 0;
        // total order on (key, slot); NaN keys sort first so they are removed first
        const bool a_gt_b = (a > b) || (a == b && ia > ib) || (isnan(b) && !isnan(a));
        if (a_gt_b == up) {
            keys[i] = b; keys[l] = a;
            idx[i] = ib; idx[l] = ia;
        }
    }
}

// ---- log-trapezoid over the history (nest2pos.py:33-42), two-stage deterministic reduction --------
__global__ void k_trap_partial(const double* hist_logL, const double* hist_logvol, long long n, LSE* partial) {
    __shared__ LSE sh[256];
    LSE acc;
    acc.init();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // interval between history entries i-1 and i, entry -1 being the dummy (-inf, 0)
        const double f0 = (i == 0) ? -INFINITY : hist_logL[i - 1];
        const double f1 = hist_logL[i];
        const double x0 = (i == 0) ? 0.0 : hist_logvol[i - 1];
        const double x1 = hist_logvol[i];
        const double fs = d_logaddexp(f0, f1) - 0.69314718055994530942;
        const double dx = x0 + log1p(-exp(x1 - x0));          // logsubexp(x0, x1)
        acc.push(fs + dx, 0.0);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) { LSE a = sh[threadIdx.x]; a.merge(sh[threadIdx.x + s]); sh[threadIdx.x] = a; }
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void k_trap_final(const LSE* partial, int nb, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    LSE a;
    a.init();
    for (int i = 0; i < nb; ++i) a.merge(partial[i]);
    *out = (a.m > -INFINITY) ? a.m + log(a.s) : -INFINITY;
}

}  // namespace cpn
