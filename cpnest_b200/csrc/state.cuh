// Device-resident state of the nested-sampling driver: what NestedSampler, _NSintegralState and
// the samplers keep in Python attributes / mp.Value (cpnest/NestedSampling.py:88-134, 254-262,
// cpnest/cpnest.py:572-574, cpnest/sampler.py:86-105).
#pragma once
namespace cpn {

constexpr int kNumKinds = 3;      // cpn_sampler: MH, slice, HMC

// Chain control of one sampler kind.  The reference keeps these per sampler process (cpnest/sampler.py:79-105); a run
// that mixes kinds (CPNest(nthreads, nslice, nhamiltonian), cpnest/cpnest.py:197-261) has one controller per kind:
// an MH step, a slice and an HMC trajectory decorrelate at very different rates.
struct KindCtl {
    double nmcmc_exact;            // Sampler.Nmcmc_exact
    double rho_target;             // chains run until the start/end correlation across the batch falls to this
    double rho_max;                // smoothed max_d |corr(start_d, end_d)| of the last batches (d includes logL)
    double rho_decay;              // two-lag control: max_d of the decaying part of that correlation (-1: not estimated)
    double ne_ref, rho_ref;        // chain length and rho_max at the last reference batch (does rho fall when chains grow?)
    unsigned long long batch[4];   // steps, accepted, evals, failed chains of this kind in the last evolve
    int nmcmc;                     // Sampler.Nmcmc
    // Two-lag control (MH): when the start/end correlation stays high although the chains are long, part of it does not decay
    // with the chain length at all (chains stay in their mode of a multimodal live set).  `twolag` != 0 makes the chains
    // record a mid-chain snapshot; the two half-chain increments do not contain the mode centre, their correlation gives
    // the within-mode correlation over half a chain (ns_kernels.cuh adapt_nmcmc), and the target applies to that part only.
    short twolag;
    short stuck;                   // (unused)
};

struct NSState {
    double logZ;          // _NSintegralState.logZ
    double B;             // info + logZ, the linear part of the information recurrence
    double info;          // _NSintegralState.info
    double logw;          // _NSintegralState.logw
    double logLmin;       // manager.logLmin
    double logLmax;       // manager.logLmax
    double condition;     // NestedSampler.condition
    double tolerance;     // NestedSampler.tolerance (stopping)
    double nmcmc_safety;
    double nmcmc_tau;     // decay of the moving average, in chains
    double slice_mu;      // SliceSampler.mu: length scale of the slice directions (sampler.py:422, 429-437)
    double hmc_dt;        // HamiltonianProposal.dt: leapfrog time step (proposal.py:393, 539-552)
    long long iteration;  // NestedSampler.iteration == number of dead points
    long long n_hist;     // len(state.logLs) - 1
    long long n_ins;      // len(insertion_indices)
    long long batches;
    unsigned long long totals[4];   // steps, accepted, evals, failed chains (whole run)
    KindCtl ctl[kNumKinds];         // chain-length control + last-evolve counters, per sampler kind
    int maxmcmc;
    int adapt;
    int stop_next;        // condition <= tolerance seen: the following batch must not run
    int done;
    int sampler;          // primary cpn_sampler of the context (the kind cpn_get_state reports)
};

}  // namespace cpn
