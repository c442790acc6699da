// Constrained Hamiltonian Monte Carlo: the device form of
//   HamiltonianMonteCarloSampler.yield_sample      cpnest/sampler.py:365-402
//   LeapFrog.get_sample (energies, log_J)          cpnest/proposal.py:605-627
//   ConstrainedLeapFrog.evolve_trajectory          cpnest/proposal.py:793-835
//     .evolve_trajectory_one_step_position         cpnest/proposal.py:714-741  (bound reflection)
//     .evolve_trajectory_one_step_momentum         cpnest/proposal.py:743-773  (reflection off logL = logLmin)
//     .sample_trajectory                           cpnest/proposal.py:837-844  (point drawn with weight exp(-H))
//   HamiltonianProposal.kinetic_energy / update_mass   cpnest/proposal.py:510-522, 562-575
// A chain is owned by G lanes, E coordinates per lane (same layout as the MH kernel).  Momenta are drawn from
// N(0, M), M = (ensemble covariance)^-1, through the eigen-decomposition the ensemble statistics already
// provide: p0 = sum_i z_i |lambda_i|^-1/2 v_i.  The kinetic energy 0.5 p.C.p (D^2 work) is only re-evaluated
// after the momentum changed (bound reflection, likelihood reflection, non-zero force).
//
// Chain length.  The reference returns after ONE accepted trajectory whose length follows the estimated
// autocorrelation length, L = base_L + randint(Nmcmc, 5 Nmcmc) (sampler.py:375, proposal.py:553-559), and
// relies on its noisy spline normals to break the regularity of the reflections.  With exact normals the
// reflected motion inside a convex (e.g. spherical) constraint is an integrable billiard: a single trajectory,
// however long, keeps its impact parameter, so the new point stays correlated with its start (measured:
// start/end correlation 0.58 whatever L, logZ of the 10-D Gaussian +4 sigma).  The production kernel therefore
// runs Nmcmc trajectories per chain, each with a fresh momentum and L = base_L + randint(10, 50) leapfrog
// steps per direction, and stops with the MH rule (steps >= Nmcmc and accepted > 0, or steps >= maxmcmc,
// sampler.py:347); Nmcmc is set by the same start/end-correlation controller as for the other samplers.
// Replay reproduces the reference's rule (stop at the first accepted trajectory, L given).
//
// Reflection.  Production reflects off the likelihood boundary in the metric of the kinetic energy,
// p -= 2 (n.C.p / n.C.n) n, so that 0.5 p.C.p is conserved exactly (the reference's Euclidean reflection,
// p -= 2 (p.n) n, conserves it only for C proportional to the identity).  Measured on the 20-D Gaussian at nlive 4000:
// mean logZ deviation -0.85 sigma over 24 runs with the Euclidean rule, +0.24 +- 0.35 sigma with the metric rule.
//
// Production: the trajectory point is drawn in one pass by weighted reservoir sampling (same distribution as
// the reference's multinomial draw over the stored trajectory, no trajectory storage); the reflection normal is
// the analytic gradient of the functor's log_likelihood (the reference estimates it with per-dimension spline
// fits over the ensemble, proposal.py:417-486).
// Replay: momenta, reflection normals and uniforms come from a tape recorded from the reference; the
// trajectory is stored and the point picked from the normalised cumulative weights as np.random.choice does.
#pragma once
#include "chain_common.cuh"

namespace cpn {

struct HMCParams : ChainParams {        // nmcmc = trajectories per chain (1 in replay: the reference stops at the first accepted one)
    // ensemble statistics
    const double* cov;        // [D][D] ensemble covariance = inverse mass matrix (proposal.py:516)
    const double* eig_scale;  // [D] sqrt|lambda_i|
    const double* eig_vec;    // [D][Dp]
    double logdet;            // log det of the mass matrix (proposal.py:520); replay only (it cancels in every difference)
    double dt_override;
    int L_override;           // > 0: trajectory length given (replay)
    int base_L;
    int max_traj;             // trajectories per chain at most (the reference loops until one is accepted)
    // replay: per chain a flat tape, per trajectory { p0[D], normal[D] per reflection in order, u of the draw, u of the test }
    const double* tape;
    const long long* tape_off;
    double* traj;             // replay scratch [chains][2L+1][Dp + 3] = {q.., logL, logP, H}
    double* logJ_out;         // [chains] log_J of the last trajectory, or null
    // production, test entry (cpn_trace_hmc): what every trajectory of every chain consumed besides the chain state:
    // trace[c][trace_cap] = { n, then per trajectory: L, p0[D], the uniforms of the weighted draw for the points 1 .. 2L-1
    // in the order they are visited, the uniform of the acceptance test }, n = doubles written after the first (-1: too small)
    double* trace;
    int trace_cap;
};

template <int G, int E, class Model, bool REPLAY>
__global__ void __launch_bounds__(kMHThreads, (E <= 4) ? 4 : 1) k_hmc_chains(const HMCParams p) {
    extern __shared__ double smem[];
    Chain<G, E, Model> ch;
    if (!ch.begin(p, smem, REPLAY)) return;
    const Group<G>& g = ch.g;
    const ModelCtx& m = ch.m;
    const int c = ch.c, D = p.D, nmcmc = ch.nmcmc;
    const bool valid = ch.valid;
    const double logLmin = ch.logLmin;
    const double dt = p.use_override ? p.dt_override : p.state->hmc_dt;
    const uint32_t cid_lo = ch.cid_lo, cid_hi = ch.cid_hi;
    double (&x)[E] = ch.x;
    double (&blo)[E] = ch.blo;
    double (&bhi)[E] = ch.bhi;
    double& logL = ch.logL;
    double& logP = ch.logP;
    double imass[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int d = g.lane + e * G;
        imass[e] = (d < D) ? __ldg(p.cov + (size_t)d * D + d) : 0.0;      // inverse_mass = diag(covariance), :519
    }
    const int Lmax = p.L_override > 0 ? p.L_override : p.base_L + 50;

    // 0.5 p.C.p (+ the constants of proposal.py:575 in replay, where they enter exp(logw - norm) rounding)
    const double tconst = REPLAY ? (-p.logdet - 0.5 * (double)D * 1.8378770664093454836) : 0.0;
    // w = C pp (this lane's coordinates).  Production: the vector is exchanged through the chain's row of shared memory (one
    // broadcast load per column; the shuffle form - a select over the lane's slots and a 64-bit shuffle per column - was the
    // bulk of a reflecting step).  Same multiply-adds in the same order in both forms.
    double* hsm = smem + (Model::kScratch ? (kMHThreads / G) * kScratchRows * p.Dp : 0) + ((int)threadIdx.x / G) * p.Dp;
    // (this lane's column of the covariance as a pointer held in a register and advanced row by row: with the address rebuilt
    // from the launch parameters for every element the mat-vec was half of all instructions the kernel executed)
    const double* cov_lane = p.cov + g.lane;
    asm volatile("" : "+l"(cov_lane));
    bool has[E];
#pragma unroll
    for (int e = 0; e < E; ++e) has[e] = g.lane + e * G < D;
    auto matvec = [&](const double (&pp)[E], double (&w)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) w[e] = 0.0;
        if (!REPLAY && G > 1) {
            __syncwarp();
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (has[e]) hsm[g.lane + e * G] = pp[e];
            __syncwarp();
        }
        const double* crow = cov_lane;
        for (int k = 0; k < D; ++k) {
            const double pk = (!REPLAY && G > 1) ? hsm[k] : coord<G, E>(g, pp, k);
#pragma unroll
            for (int e = 0; e < E; ++e) w[e] = fma(has[e] ? __ldg(crow + e * G) : 0.0, pk, w[e]);
            crow += D;
        }
    };
    auto kinetic = [&](const double (&pp)[E]) -> double {
        double w[E];
        matvec(pp, w);
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) acc = fma(pp[e], w[e], acc);
        return 0.5 * g.sum(acc) + tconst;
    };

    // replay: one IEEE operation per reference (numpy) operation, no FMA contraction
    auto mul = [](double a, double b) -> double { return REPLAY ? __dmul_rn(a, b) : a * b; };
    auto add = [](double a, double b) -> double { return REPLAY ? __dadd_rn(a, b) : a + b; };

    int steps = 0, accepted = 0, evals = 0;
    bool active = true;
    double last_logJ = 0.0;
    const double* tp = REPLAY ? p.tape + p.tape_off[c] : nullptr;
    const int npts = 2 * Lmax + 1;
    double* traj = REPLAY ? p.traj + (size_t)c * npts * (p.Dp + 3) : nullptr;
    double* tr = (!REPLAY && p.trace != nullptr && valid) ? p.trace + (size_t)c * p.trace_cap : nullptr;
    int ti = 1;                   // next free double of the chain's trace area

    for (int attempt = 0; attempt < p.max_traj; ++attempt) {
        if (!__any_sync(FULL, active)) break;
        // trajectory length: given (replay), else base_L + randint(10, 50) (cf. proposal.py:553-559); the leapfrog
        // loops hold warp-wide shuffles, so the chains sharing a warp use the length drawn by the first of them
        int L = p.L_override;
        if (L <= 0) {
            const Philox4 r = philox4x32_10(cid_lo, cid_hi, (uint32_t)attempt, 2u, p.seed_lo, p.seed_hi);
            L = p.base_L + 10 + (int)u32_to_index(r.v[0], 40u);
            L = __shfl_sync(FULL, L, 0);
        }
        // ---- momentum draw: np.atleast_1d(momenta_distribution.rvs()), proposal.py:620 ------------------
        double p0[E];
        if (REPLAY) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                p0[e] = (active && d < D) ? tp[d] : 0.0;
            }
            if (active) tp += D;
        } else {
            double z[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                const Philox4 rz = philox4x32_10(cid_lo, cid_hi, (uint32_t)attempt, 0x80000000u | (uint32_t)d, p.seed_lo, p.seed_hi);
                double g0, g1;
                box_muller(rz.v[0], rz.v[1], g0, g1);
                const double sc = __ldg(p.eig_scale + min(d, D - 1));
                z[e] = (d < D && sc > 0.0) ? g0 / sc : 0.0;                            // z_i |lambda_i|^-1/2
                p0[e] = 0.0;
            }
            const double* vrow = p.eig_vec + g.lane;
            for (int i = 0; i < D; ++i) {
                const double zi = coord<G, E>(g, z, i);
#pragma unroll
                for (int e = 0; e < E; ++e) p0[e] = fma(zi, has[e] ? __ldg(vrow + e * G) : 0.0, p0[e]);
                vrow += p.Dp;
            }
        }
        bool rec = false;             // this trajectory of this chain is recorded
        int ti_u = 0;                 // where its uniforms start
        if (!REPLAY && tr != nullptr && active) {
            if (ti + D + 2 * L + 2 > p.trace_cap) {
                if (g.lane == 0) tr[0] = -1.0;
                tr = nullptr;
            } else {
                rec = true;
                if (g.lane == 0) tr[ti] = (double)L;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int d = g.lane + e * G;
                    if (d < D) tr[ti + 1 + d] = p0[e];
                }
                ti_u = ti + 1 + D;
                ti = ti_u + (2 * L - 1) + 1;
            }
        }
        const double E0 = kinetic(p0) - logP;                           // hamiltonian(p0, q0), :621

        // ---- the two half trajectories + the weighted draw ----------------------------------------------
        double Wsum = 0.0, Href = 0.0, Hprev = 0.0, wprev = 0.0;        // production: running sum of the weights exp(Href - H)
        double qc[E], logLc = CPN_NEG_INF, logPc = CPN_NEG_INF, Hc = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) qc[e] = x[e];
        int point = 0;                                                  // index in the reference's trajectory list
        Philox4 ru;
        ru.v[0] = ru.v[1] = ru.v[2] = ru.v[3] = 0u;
        for (int dirn = 0; dirn < 2; ++dirn) {
            double pm[E], q[E], gv[E];
#pragma unroll
            for (int e = 0; e < E; ++e) { pm[e] = dirn == 0 ? p0[e] : -p0[e]; q[e] = x[e]; }
            Model::template grad_potential<G, E>(g, q, m, gv);
#pragma unroll
            for (int e = 0; e < E; ++e) pm[e] = add(pm[e], mul(mul(-0.5, dt), gv[e]));     // half step, :757-759
            double T = 0.0;
            bool dirty = true;
            for (int i = 0; i < L; ++i) {
                ++point;
                // position step with reflection off the prior bounds, :729-741
                bool flipped = false, outside = false;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    q[e] = add(q[e], mul(mul(dt, pm[e]), imass[e]));
                    outside = outside | (q[e] < blo[e]) | (q[e] > bhi[e]);
                }
                // the reflection loop of :733-741 runs only when some coordinate of some chain of the warp left the box, and
                // as a loop: unrolled and predicated (64 rounds x E coordinates) it was 5 of every 6 instructions of a step
                if (__any_sync(FULL, outside)) {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
#pragma unroll 1
                        for (int guard = 0; guard < 64 && (q[e] < blo[e] || q[e] > bhi[e]); ++guard) {
                            if (q[e] > bhi[e]) { q[e] = bhi[e] - (q[e] - bhi[e]); pm[e] *= -1.0; flipped = true; }
                            if (q[e] < blo[e]) { q[e] = blo[e] + (blo[e] - q[e]); pm[e] *= -1.0; flipped = true; }
                        }
                    }
                }
                // momentum step, :760-773
                bool inb = true;
#pragma unroll
                for (int e = 0; e < E; ++e) inb = inb & (blo[e] < q[e]) & (q[e] < bhi[e]);
                inb = g.all(inb);
                double lp = Model::template log_prior<G, E>(g, q, m);
                lp = inb ? lp : CPN_NEG_INF;                            // q.logP = -V(q), :790
                const double ll = Model::template log_likelihood<G, E>(g, q, m);   // :791
                evals += active ? 1 : 0;
                Model::template grad_potential<G, E>(g, q, m, gv);
                const bool inside = ll - logLmin > 0.0;
                double nrm[E];
                if (REPLAY) {
                    const bool refl = active && !inside;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int d = g.lane + e * G;
                        nrm[e] = (refl && d < D) ? tp[d] : 0.0;
                    }
                    if (refl) tp += D;
                } else if (__any_sync(FULL, active && !inside)) {       // the normal is only needed by a reflecting chain
                    Model::template grad_log_likelihood<G, E>(g, q, m, nrm);
                    double n2 = 0.0;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        nrm[e] = (g.lane + e * G < D) ? nrm[e] : 0.0;
                        n2 = fma(nrm[e], nrm[e], n2);
                    }
                    n2 = g.sum(n2);
                    const double inv = (n2 > 0.0) ? 1.0 / sqrt(n2) : 0.0;
#pragma unroll
                    for (int e = 0; e < E; ++e) nrm[e] *= inv;           // unit_normal, :469-486
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) nrm[e] = 0.0;
                }
                // Reflection off the likelihood boundary.  The reference reflects in the Euclidean metric,
                // p -= 2 (p.n) n (:769-771; replay), which changes the kinetic energy 0.5 p.C.p unless C is a
                // multiple of the identity; every such change feeds the (inexact, see the header) trajectory draw +
                // energy test.  Production reflects in the metric of the kinetic energy instead,
                // p -= 2 (n.C.p / n.C.n) n, which conserves it exactly: under a flat prior H is then constant along
                // the trajectory, the draw is uniform and the energy test always passes.
                double pn = 0.0, scale = 2.0;
                if (REPLAY) {
#pragma unroll
                    for (int e = 0; e < E; ++e) pn = fma(pm[e], nrm[e], pn);
                    pn = g.sum(pn);
                } else if (__any_sync(FULL, active && !inside)) {
                    double cn[E];
                    matvec(nrm, cn);
                    double ncp = 0.0, ncn = 0.0;
#pragma unroll
                    for (int e = 0; e < E; ++e) { ncp = fma(cn[e], pm[e], ncp); ncn = fma(cn[e], nrm[e], ncn); }
                    ncp = g.sum(ncp);
                    ncn = g.sum(ncn);
                    pn = (ncn > 0.0) ? ncp / ncn : 0.0;
                }
#pragma unroll
                for (int e = 0; e < E; ++e) pm[e] = add(pm[e], inside ? mul(-dt, gv[e]) : mul(mul(-scale, pn), nrm[e]));
                dirty = dirty || (REPLAY && !inside) || !Model::kFlatPrior;
                if (g.all(!flipped) == false) dirty = true;
                if (__any_sync(FULL, dirty)) {
                    const double Tn = kinetic(pm);
                    T = dirty ? Tn : T;
                    dirty = false;
                }
                const double H = T - lp;                                 // hamiltonian(p, q) = T(p) + V(q)
                const bool last = (dirn == 1 && i == L - 1);             // trajectory[1:-1], :841
                if (REPLAY) {
                    if (valid && active) {
                        double* row = traj + (size_t)point * (p.Dp + 3);
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            const int d = g.lane + e * G;
                            if (d < D) row[d] = q[e];
                        }
                        if (g.lane == 0) { row[p.Dp] = ll; row[p.Dp + 1] = lp; row[p.Dp + 2] = H; }
                    }
                } else if (!last) {
                    // weighted reservoir: keep this point with probability w / (sum of weights so far)
                    // (a Philox block serves four points: it is drawn when the first of them is visited)
                    if (point == 1 || (point & 3) == 0)
                        ru = philox4x32_10(cid_lo, cid_hi, (uint32_t)attempt, 0x40000000u | (uint32_t)(point >> 2), p.seed_lo, p.seed_hi);
                    const int k = point & 3;
                    const uint32_t wv = (k & 2) ? ((k & 1) ? ru.v[3] : ru.v[2]) : ((k & 1) ? ru.v[1] : ru.v[0]);
                    const double uu = u32_to_unit_open(wv);
                    if (rec && g.lane == 0) tr[ti_u + point - 1] = uu;
                    // weights exp(-H) relative to a reference energy, summed in the linear domain: under a flat prior H is
                    // constant along the trajectory and a point costs an add, a multiply and a compare (the log-domain form
                    // took a log1p, an exp and a log in fp64 per point).  The reference energy follows H downwards.
                    bool take = false;
                    if (H < -CPN_NEG_INF) {                              // a point outside the box has weight 0
                        double wgt;
                        if (!(Wsum > 0.0)) { Href = H; wgt = 1.0; }
                        else if (H == Hprev) wgt = wprev;
                        else {
                            if (H < Href - 50.0) { Wsum *= exp(H - Href); Href = H; }
                            wgt = exp(Href - H);
                        }
                        Hprev = H; wprev = wgt;
                        Wsum += wgt;
                        take = uu * Wsum < wgt;
                    }
#pragma unroll
                    for (int e = 0; e < E; ++e) qc[e] = take ? q[e] : qc[e];
                    logLc = take ? ll : logLc;
                    logPc = take ? lp : logPc;
                    Hc = take ? H : Hc;
                }
            }
        }
        double u_acc;
        if (REPLAY) {
            // np.random.choice(range(1, n-1), p = exp(logw - logsumexp(logw))), :840-843: one uniform, searchsorted
            // on the normalised cumulative sum
            __syncwarp();
            const double u = active ? tp[0] : 0.5;
            u_acc = active ? tp[1] : 0.5;
            if (active) tp += 2;
            const int n = 2 * L - 1;                                     // points 1 .. 2L-1
            double mx = CPN_NEG_INF;
            for (int k = 1; k <= n; ++k) mx = fmax(mx, -traj[(size_t)k * (p.Dp + 3) + p.Dp + 2]);
            double ssum = 0.0;
            for (int k = 1; k <= n; ++k) ssum += exp(-traj[(size_t)k * (p.Dp + 3) + p.Dp + 2] - mx);
            const double norm = mx + log(ssum);
            double tot = 0.0;
            for (int k = 1; k <= n; ++k) tot += exp(-traj[(size_t)k * (p.Dp + 3) + p.Dp + 2] - norm);
            double cum = 0.0;
            int idx = n;
            for (int k = 1; k <= n; ++k) {
                cum += exp(-traj[(size_t)k * (p.Dp + 3) + p.Dp + 2] - norm);
                if (u < cum / tot) { idx = k; break; }                    // searchsorted(u, side='right')
            }
            const double* row = traj + (size_t)idx * (p.Dp + 3);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                qc[e] = (d < D) ? row[d] : 0.0;
            }
            logLc = row[p.Dp]; logPc = row[p.Dp + 1]; Hc = row[p.Dp + 2];
            __syncwarp();
        } else {
            const Philox4 ra = philox4x32_10(cid_lo, cid_hi, (uint32_t)attempt, 1u, p.seed_lo, p.seed_hi);
            u_acc = u32_to_unit_open(ra.v[0]);
            if (rec && g.lane == 0) tr[ti_u + 2 * L - 1] = u_acc;
        }
        const double log_J = fmin(0.0, E0 - Hc);                         // :626
        const bool acc = active && (log_J > log(u_acc)) && (logLc > logLmin);   // sampler.py:380-382
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = acc ? qc[e] : x[e];
        logL = acc ? logLc : logL;
        logP = acc ? logPc : logP;
        if (active) {
            ++steps;
            last_logJ = log_J;
            accepted += acc ? 1 : 0;
            // replay (nmcmc = 1): `while sub_accepted == 0`, sampler.py:375; production: the rule of sampler.py:347
            if ((steps >= nmcmc && accepted > 0) || steps >= p.max_traj) active = false;
        }
    }

    ch.finish(p, steps, accepted, evals, 0, 0);
    if (tr != nullptr && g.lane == 0) tr[0] = (double)(ti - 1);
    if (valid) {
        if (p.logJ_out != nullptr && g.lane == 0) p.logJ_out[c] = last_logJ;
    }
}

}  // namespace cpn
