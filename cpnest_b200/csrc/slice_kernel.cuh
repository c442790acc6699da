// Ensemble slice sampling of the likelihood-constrained prior: the device form of
//   SliceSampler.yield_sample                       cpnest/sampler.py:465-569
//   SliceSampler.reset_boundaries / increase_*      cpnest/sampler.py:439-463
//   EnsembleSliceProposalCycle.get_direction        cpnest/proposal.py:199-207
//   EnsembleSliceDifferential / Gaussian / CorrelatedGaussian   cpnest/proposal.py:121-187
// One launch evolves all chains of a nested-sampling batch, same lane layout as the MH kernel
// (a chain = G lanes, E coordinates per lane).  Per slice: a direction through the current point,
// stepping out in unit steps of the slice coordinate while the point stays in the box (and above the
// prior slice level), then shrinkage towards the current point until a point with log_prior above
// the slice level AND logL > logLmin is found.
#pragma once
#include "chain_common.cuh"

namespace cpn {

struct SliceParams : ChainParams {      // maxmcmc is also max_steps_out and max_slices (sampler.py:424-425)
    // direction cycle (cpn_slice_direction per slot) and the ensemble statistics behind the
    // correlated-Gaussian direction: N(mean, cov) = mean + sum_i z_i sqrt|lambda_i| v_i
    const uint8_t* cycle;
    int cycle_len, cycle_pos0;
    const double* mean;       // [D]
    const double* eig_scale;  // [D]
    const double* eig_vec;    // [D][Dp]
    double mu_override;       // slice length scale (SliceSampler.mu)
    // replay (parity) mode: per chain a flat tape  { uL, direction data, Exp(1), uJ, X'_1, X'_2, ... } per slice,
    // direction data = {i0, i1} (differential) or the D-vector numpy returned (Gaussian kinds)
    const double* tape;
    const long long* tape_off;    // [chains + 1]
    int compat;
    double* last;                 // [chains][4] = {Ne, Nc, L, R} of the chain's last slice, or null
    double* dir_out;              // [chains][D] direction of the chain's first slice, or null (tests)
    // production, test entry (cpn_trace_slice): what every slice of every chain consumed besides the chain state, in the order
    // of the replay tape: trace[c][trace_cap] = { n, then per slice: kind, nX, uL, direction data ({i0, i1} / the D-vector),
    // Exp(1), uJ, X'_1 .. X'_nX }, n = doubles written after the first (-1: the area was too small), or null
    double* trace;
    int trace_cap;
};

template <int G, int E, class Model, bool REPLAY>
__global__ void __launch_bounds__(kMHThreads) k_slice_chains(const SliceParams p) {
    extern __shared__ double smem[];
    Chain<G, E, Model> ch;
    if (!ch.begin(p, smem, REPLAY)) return;
    const Group<G>& g = ch.g;
    const ModelCtx& m = ch.m;
    const int c = ch.c, D = p.D, nmcmc = ch.nmcmc, maxmcmc = p.maxmcmc;
    const bool valid = ch.valid;
    const double logLmin = ch.logLmin;
    const double mu = p.use_override ? p.mu_override : p.state->slice_mu;
    const uint32_t cid_lo = ch.cid_lo, cid_hi = ch.cid_hi;
    double (&x)[E] = ch.x;
    double (&blo)[E] = ch.blo;
    double (&bhi)[E] = ch.bhi;
    double& logL = ch.logL;
    double& logP = ch.logP;

    int steps = 0, accepted = 0, evals = 0, ne_sum = 0, nc_sum = 0;
    double Ne = 0.0, Nc = 0.0, L = 0.0, R = 0.0;
    bool active = true;
    const int clen = p.cycle_len;
    int pos = p.cycle_pos0;
    const double* tp = REPLAY ? p.tape + p.tape_off[c] : nullptr;
    const uint32_t ens_n = (uint32_t)p.ens_n;
    const double mu_diff = (REPLAY && p.compat) ? (double)(float)mu : mu;   // LivePoint * float, parameter.pyx:96

    double* tr = (!REPLAY && p.trace != nullptr && valid) ? p.trace + (size_t)c * p.trace_cap : nullptr;
    int ti = 1, ti_nx = 0;        // next free double of the chain's trace area; where the current slice's nX goes

    // point on the slice: direction * t + old (LivePoint.__mul__ then __add__; t rounded to fp32 in the reference)
    auto along = [&](const double (&dir)[E], double t, double (&out)[E]) {
        const double ts = (REPLAY && p.compat) ? (double)(float)t : t;
#pragma unroll
        for (int e = 0; e < E; ++e) out[e] = REPLAY ? __dadd_rn(__dmul_rn(dir[e], ts), x[e]) : fma(dir[e], ts, x[e]);
    };
    auto in_box = [&](const double (&v)[E]) -> bool { return ch.in_box(v); };

    for (int s = 0;; ++s) {
        if (!__any_sync(FULL, active)) break;
        pos = (pos + 1 == clen) ? 0 : pos + 1;                       // pre-increment, proposal.py:203
        const int kind = p.cycle[pos];

        // ---- draws of this slice + direction ----------------------------------------------------------
        double uL, eexp, uJ;
        double dir[E];
        double trace_a = 0.0, trace_b = 0.0, trace_v[E];     // direction data as the replay tape holds it (trace entry only)
#pragma unroll
        for (int e = 0; e < E; ++e) trace_v[e] = 0.0;
        if (REPLAY) {
            uL = active ? tp[0] : 0.5;
            const double* q = tp + 1;
            if (kind == CPN_SLICE_DIFF) {
                const uint32_t i0 = active ? (uint32_t)q[0] : 0u, i1 = active ? (uint32_t)q[1] : 1u;
                double ra[E], rb[E];
                load_row<G, E>(g, p.ens_x + (size_t)i0 * p.Dp, D, ra);
                load_row<G, E>(g, p.ens_x + (size_t)i1 * p.Dp, D, rb);
#pragma unroll
                for (int e = 0; e < E; ++e) dir[e] = __dmul_rn(__dsub_rn(ra[e], rb[e]), mu_diff);   // proposal.py:131-133
                q += 2;
            } else {
                double z[E], n2 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int d = g.lane + e * G;
                    z[e] = (active && d < D) ? q[d] : 0.0;
                    n2 += z[e] * z[e];
                }
                if (kind == CPN_SLICE_GAUSS) {
                    const double nrm = sqrt(g.sum(n2));               // np.linalg.norm, proposal.py:185
#pragma unroll
                    for (int e = 0; e < E; ++e) dir[e] = __dmul_rn(__ddiv_rn(z[e], nrm), mu);    // :186-187
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) dir[e] = __dmul_rn(mu, z[e]);                     // :172
                }
                q += D;
            }
            eexp = active ? q[0] : 1.0;
            uJ = active ? q[1] : 0.5;
            if (active) tp = q + 2;
        } else {
            const Philox4 r = philox4x32_10(cid_lo, cid_hi, (uint32_t)s, 0u, p.seed_lo, p.seed_hi);
            uL = u32_to_unit_open(r.v[0]);                           // np.random.uniform(0,1), sampler.py:444
            uJ = u32_to_unit_open(r.v[1]);                           // :491
            eexp = -log(u32_to_unit_open(r.v[2]));                   // np.random.exponential(), :490
            if (kind == CPN_SLICE_DIFF) {
                const Philox4 r2 = philox4x32_10(cid_lo, cid_hi, (uint32_t)s, 1u, p.seed_lo, p.seed_hi);
                const uint32_t i0 = u32_to_index(r.v[3], ens_n);     // sample(list(ensemble), 2), proposal.py:131
                uint32_t i1 = u32_to_index(r2.v[0], ens_n - 1);
                i1 += (i1 >= i0);
                double ra[E], rb[E];
                load_row<G, E>(g, p.ens_x + (size_t)i0 * p.Dp, D, ra);
                load_row<G, E>(g, p.ens_x + (size_t)i1 * p.Dp, D, rb);
#pragma unroll
                for (int e = 0; e < E; ++e) dir[e] = (ra[e] - rb[e]) * mu;
                trace_a = (double)i0; trace_b = (double)i1;
            } else {
                // one N(0,1) per coordinate: block (chain, slice, 0x80000000 | d)
                double z[E], n2 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int d = g.lane + e * G;
                    const Philox4 rz = philox4x32_10(cid_lo, cid_hi, (uint32_t)s, 0x80000000u | (uint32_t)d, p.seed_lo, p.seed_hi);
                    double g0, g1;
                    box_muller(rz.v[0], rz.v[1], g0, g1);
                    z[e] = (d < D) ? g0 : 0.0;
                    n2 += z[e] * z[e];
                }
                if (kind == CPN_SLICE_GAUSS) {
                    const double sc = mu / sqrt(g.sum(n2));           // unit vector * mu, proposal.py:184-187
#pragma unroll
                    for (int e = 0; e < E; ++e) { dir[e] = z[e] * sc; trace_v[e] = z[e]; }
                } else {
                    // mu * N(mean, cov) (the mean is not subtracted, proposal.py:172)
                    double acc[E];
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int d = g.lane + e * G;
                        acc[e] = (d < D) ? __ldg(p.mean + d) : 0.0;
                    }
                    // (this lane's column of the eigenvector table as a running pointer: rebuilt from the launch parameters
                    // per element the address arithmetic outweighed the multiply-adds, as in the HMC kernel's mat-vec)
                    const double* vrow = p.eig_vec + g.lane;
                    const double* esc = p.eig_scale;
                    for (int i = 0; i < D; ++i) {
                        const double zi = coord<G, E>(g, z, i) * __ldg(esc + i);
#pragma unroll
                        for (int e = 0; e < E; ++e) acc[e] = fma(zi, (g.lane + e * G < D) ? __ldg(vrow + e * G) : 0.0, acc[e]);
                        vrow += p.Dp;
                    }
#pragma unroll
                    for (int e = 0; e < E; ++e) { dir[e] = mu * acc[e]; trace_v[e] = acc[e]; }
                }
            }
        }
        bool rec = false;             // this slice of this chain is recorded
        if (!REPLAY && tr != nullptr && active) {
            if (ti + D + 8 > p.trace_cap) {
                if (g.lane == 0) tr[0] = -1.0;
                tr = nullptr;
            } else {
                rec = true;
                if (g.lane == 0) { tr[ti] = (double)kind; tr[ti + 2] = uL; }
                ti_nx = ti + 1;
                ti += 3;
                if (kind == CPN_SLICE_DIFF) {
                    if (g.lane == 0) { tr[ti] = trace_a; tr[ti + 1] = trace_b; }
                    ti += 2;
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int d = g.lane + e * G;
                        if (d < D) tr[ti + d] = trace_v[e];
                    }
                    ti += D;
                }
                if (g.lane == 0) { tr[ti] = eexp; tr[ti + 1] = uJ; tr[ti_nx] = 0.0; }
                ti += 2;
            }
        }
        int nX = 0;
        if (p.dir_out != nullptr && s == 0 && valid) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int d = g.lane + e * G;
                if (d < D) p.dir_out[(size_t)c * D + d] = dir[e];
            }
        }

        // ---- initial interval and stepping out (sampler.py:439-447, 489-524) ----------------------------
        L = -uL;
        R = L + 1.0;
        Ne = 0.0;
        Nc = 0.0;
        const double Yp = logP - eexp;                               // :490
        int J = (int)floor((double)maxmcmc * uJ);                    // :491
        int K = (maxmcmc - 1) - J;                                   // :492
        if (!REPLAY && Model::kFlatPrior) {
            // Under the flat prior of Model.log_prior (model.py:66-82) stepping out only ends at the box or
            // when its budget is spent: the number of unit steps follows from the interval of the slice
            // coordinate that lies inside the box.
            double tlo = CPN_NEG_INF, thi = -CPN_NEG_INF;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double a = (blo[e] - x[e]) / dir[e], b = (bhi[e] - x[e]) / dir[e];
                const bool use = dir[e] != 0.0 && (blo[e] > CPN_NEG_INF);
                tlo = use ? fmax(tlo, fmin(a, b)) : tlo;
                thi = use ? fmin(thi, fmax(a, b)) : thi;
            }
            tlo = g.max(tlo);
            thi = g.min(thi);
            const double kl = fmin((double)J, fmax(0.0, ceil(L - tlo)));
            const double kr = fmin((double)K, fmax(0.0, ceil(thi - R)));
            L -= kl;
            R += kr;
            Ne = kl + kr;
        } else {
            bool go = active && J > 0;
            while (__any_sync(FULL, go)) {                           // :495-508
                double xt[E];
                along(dir, L, xt);
                const bool inb = in_box(xt);
                const double lp = Model::template log_prior<G, E>(g, xt, m);
                const bool grow = go && inb && !(Yp > lp);
                L = grow ? L - 1.0 : L;
                Ne = grow ? Ne + 1.0 : Ne;
                J = grow ? J - 1 : J;
                go = grow && J > 0;
            }
            go = active && K > 0;
            while (__any_sync(FULL, go)) {                           // :512-524
                double xt[E];
                along(dir, R, xt);
                const bool inb = in_box(xt);
                const double lp = Model::template log_prior<G, E>(g, xt, m);
                const bool grow = go && inb && !(Yp > lp);
                R = grow ? R + 1.0 : R;
                Ne = grow ? Ne + 1.0 : Ne;
                K = grow ? K - 1 : K;
                go = grow && K > 0;
            }
        }

        // ---- shrinkage (sampler.py:529-551) ---------------------------------------------------------------
        bool searching = active;
        Philox4 ru;
        ru.v[0] = ru.v[1] = ru.v[2] = ru.v[3] = 0u;
        for (int sl = 0; sl < maxmcmc; ++sl) {
            if (!__any_sync(FULL, searching)) break;
            double Xp;
            if (REPLAY) {
                Xp = searching ? tp[0] : 0.0;                        // the value np.random.uniform(L, R) returned
                if (searching) ++tp;
            } else {
                // (a Philox block serves four candidates: it is drawn when the first of them is asked for)
                if ((sl & 3) == 0) ru = philox4x32_10(cid_lo, cid_hi, (uint32_t)s, 0x40000000u | (uint32_t)(sl >> 2), p.seed_lo, p.seed_hi);
                const uint32_t w = (sl & 2) ? ((sl & 1) ? ru.v[3] : ru.v[2]) : ((sl & 1) ? ru.v[1] : ru.v[0]);
                Xp = fma(R - L, u32_to_unit_open(w), L);
                if (rec && searching) {
                    if (ti >= p.trace_cap) {
                        if (g.lane == 0) tr[0] = -1.0;
                        tr = nullptr; rec = false;
                    } else {
                        if (g.lane == 0) tr[ti] = Xp;
                        ++ti; ++nX;
                    }
                }
            }
            double xn[E];
            along(dir, Xp, xn);
            const bool inb = in_box(xn);
            double logP_new = Model::template log_prior<G, E>(g, xn, m);
            logP_new = inb ? logP_new : CPN_NEG_INF;
            const bool pass = searching && (logP_new > Yp);          // :535
            bool acc = false;
            if (Model::kCheap || __any_sync(FULL, pass)) {
                const double logL_new = Model::template log_likelihood<G, E>(g, xn, m);   // :537
                evals += pass ? 1 : 0;
                acc = pass && (logL_new > logLmin);                  // :538
#pragma unroll
                for (int e = 0; e < E; ++e) x[e] = acc ? xn[e] : x[e];
                logL = acc ? logL_new : logL;
                logP = acc ? logP_new : logP;
                accepted += acc ? 1 : 0;
            }
            const bool shrink = searching && !acc;
            if (shrink && Xp < 0.0) { L = Xp; Nc += 1.0; }           // :545-550
            else if (shrink && Xp > 0.0) { R = Xp; Nc += 1.0; }
            searching = shrink;
        }
        if (rec && g.lane == 0) tr[ti_nx] = (double)nX;
        if (active) {
            ++steps;
            ne_sum += (int)Ne;
            nc_sum += (int)Nc;
            if ((steps > nmcmc && accepted > 0) || steps > maxmcmc) active = false;   // :553-557
        }
    }

    ch.finish(p, steps, accepted, evals, ne_sum, nc_sum);
    if (tr != nullptr && g.lane == 0) tr[0] = (double)(ti - 1);
    if (valid) {
        if (p.last != nullptr && g.lane == 0) {
            double* q = p.last + (size_t)c * 4;
            q[0] = Ne; q[1] = Nc; q[2] = L; q[3] = R;
        }
    }
}

}  // namespace cpn
