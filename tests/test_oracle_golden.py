"""Pin oracle/cpnest_oracle.py against the fixtures generated from the unmodified
reference (tests/golden/gen_golden.py).  CPU only."""
import math
import os

import numpy as np
import pytest

from oracle import cpnest_oracle as O

RTOL = 1e-12


def close(a, b, rtol=RTOL, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_inf | (np.abs(a - b) <= atol + rtol * np.abs(b))
    return bool(np.all(ok))


def test_integrator_golden(golden):
    for case in golden("integrator.json"):
        st = O.IntegralState(case["nlive"])
        for (logL, nl, nrep), (logZ, info, logw) in zip(case["steps"], case["trace"]):
            st.increment(logL, nlive=nl, nreplace=nrep)
            assert st.logZ == logZ and st.info == info and st.logw == logw, case["tag"]
        if "finalised_logZ" in case:
            assert st.finalise() == case["finalised_logZ"]


def test_gv1_gv2_values(golden):
    g = {c["tag"]: c for c in golden("integrator.json")}
    assert g["GV1"]["trace"][-1] == [-2.4325473298423219, 1.3084925082869918, -0.40000000000000002]
    assert g["GV2"]["trace"][-1][0] == -1.8260744632029229
    assert g["GV2"]["finalised_logZ"] == -2.1688217513916106


def test_nest2pos_golden(golden):
    g = golden("nest2pos.json")
    c = g["logsubexp"]
    assert close(O.logsubexp(np.array(c["x"]), np.array(c["y"])), c["z"], 0.0)
    c = g["log_trap"]
    assert O.log_integrate_log_trap(c["log_func"], c["log_support"]) == c["value"]
    for c in g["compute_weights"]:
        ev, w = O.compute_weights(np.array(c["data"]), c["nlive"])
        assert ev == c["log_ev"]
        assert close(w, c["log_wts"], 0.0)
    for c in g["acl"]:
        assert close(O.autocorrelation(np.array(c["x"])), c["acf"], 0.0)
        assert O.acl(np.array(c["x"])) == c["acl"]
    gv4 = [c for c in g["acl"] if c["tag"] == "GV4"][0]
    assert abs(gv4["acl"] - 8.7330809567711487) < 1e-13


def test_cycle_golden(golden):
    for c in golden("cycle.json"):
        rs = np.random.RandomState(c["seed"])
        cyc = O.make_cycle(rs)
        assert "".join("WSDE"[i] for i in cyc) == c["cycle"]
    assert golden("cycle.json")[0]["cycle"].startswith("SDDEESSEEEDDEEDDDWEEDDWDEDDESDEDEWEESEDEWSWEDDWDSDW")


def _replay_proposal(call, case):
    ens = np.array(case["ensemble"])
    tp = O.TapeReader(call["tape"])
    old = np.array(call["old"])
    k = call["kind"]
    if k == "walk":
        idx = tp.take("sample")
        g = [tp.take("gauss") for _ in range(3)]
        return O.ensemble_walk(old, ens, idx, g)
    if k == "stretch":
        return O.ensemble_stretch(old, ens, tp.take("choice"), tp.take("uniform"))
    if k == "de":
        ia, ib = tp.take("sample")
        return O.differential_evolution(old, ens, ia, ib, tp.take("gauss"))
    i = tp.take("randrange")
    return O.ensemble_eigenvector(old, case["eigen_values"], case["eigen_vectors"], i, tp.take("gauss"))


def test_proposals_golden(golden):
    n = 0
    for case in golden("proposals.json"):
        for call in case["calls"]:
            new, log_J = _replay_proposal(call, case)
            # bit-exact: same fp64 operations in the same order, incl. the fp32 scalar rounding
            assert np.array_equal(new, np.array(call["new"])), (case["D"], call["kind"])
            assert log_J == call["log_J"]
            n += 1
    assert n > 80


def test_fp32_scalar_truncation_matters(golden):
    """SURVEY hard part 2: without (double)(float)s the reference's numbers are NOT reproduced."""
    case = golden("proposals.json")[3]
    call = [c for c in case["calls"] if c["kind"] == "de"][0]
    tp = O.TapeReader(call["tape"])
    ia, ib = tp.take("sample")
    new, _ = O.differential_evolution(np.array(call["old"]), np.array(case["ensemble"]), ia, ib,
                                      tp.take("gauss"), compat=False)
    assert not np.array_equal(new, np.array(call["new"]))
    assert close(new, call["new"], 1e-6, 1e-6)


def test_covariance_eigen_golden(golden):
    for case in golden("proposals.json"):
        mean, cov, w, v = O.ensemble_covariance(np.array(case["ensemble"]))
        assert close(cov, case["covariance"], 1e-13, 1e-15)
        assert close(w, case["eigen_values"], 1e-12, 1e-15)
        V = np.array(case["eigen_vectors"])
        for i in range(len(w)):
            s = np.sign(np.dot(V[:, i], v[:, i]))
            assert close(s * v[:, i], V[:, i], 1e-9, 1e-10)
    gv = [c for c in golden("proposals.json") if c["D"] == 3][0]
    assert close(gv["covariance"][0], [0.3895397209973588, -0.1645720041869938, 0.06439469792129816], 1e-14)
    assert close(gv["eigen_values"], [0.05931699065096262, 0.34275156822160785, 0.665821244936251], 1e-13)


def _model_for(name, case):
    if name == "gaussian_mixture":
        return O.builtin_model(name, data=case["data"])
    return O.builtin_model(name)


def test_models_golden(golden):
    for name, case in golden("models.json").items():
        m = _model_for(name, case)
        assert m.names == case["names"]
        assert m.bounds == case["bounds"]
        for x, inb, lp, ll in zip(case["X"], case["in_bounds"], case["logP"], case["logL"]):
            x = np.array(x)
            assert m.in_bounds(x) == inb
            assert close(m.log_prior(x), lp, 1e-14), name
            if math.isnan(ll):
                assert math.isnan(m.log_likelihood(x))
            else:
                assert close(m.log_likelihood(x), ll, 2e-13 if name == "gaussian_mixture" else 1e-13), name


def test_mh_chain_golden(golden):
    mg = golden("models.json")
    nchains = 0
    for case in golden("mh_chain.json"):
        m = _model_for(case["model"], mg[case["model"]])
        pool = [np.array(p) for p in case["pool"]]
        pool_logL = list(case["pool_logL"])
        pool_logP = list(case["pool_logP"])
        idx = 0
        for ch in case["chains"]:
            assert idx == ch["cycle_idx_before"]
            r = O.mh_chain(m, pool, pool_logL, pool_logP, case["cycle"], idx, case["eigen_values"],
                           case["eigen_vectors"], ch["tape"], case["Nmcmc"], case["maxmcmc"],
                           case["logLmin"], record=True)
            idx = r["cycle_idx"]
            assert r["steps"] == ch["steps"]
            assert r["accepted"] == round(ch["sub_acceptance"] * ch["steps"])
            assert np.array_equal(r["out"], np.array(ch["out"])), case["model"]
            assert close(r["logL"], ch["out_logL"], 1e-13)
            assert close(r["logP"], ch["out_logP"], 1e-13)
            assert np.array_equal(np.array(r["states"]), np.array(ch["states"]))
            nchains += 1
    assert nchains >= 20


def test_slice_chain_golden(golden):
    """SliceSampler.yield_sample (sampler.py:465-569): chain outcome, interval, expansion/contraction
    counts and the adapted length scale, bit for bit, from the reference's recorded random numbers."""
    mg = golden("models.json")
    nchains = 0
    for case in golden("slice_chain.json"):
        m = _model_for(case["model"], mg[case["model"]])
        pool = [np.array(p) for p in case["pool"]]
        pool_logL = list(case["pool_logL"])
        pool_logP = list(case["pool_logP"])
        idx, mu, counter = 0, case["mu0"], 0
        for ch in case["chains"]:
            assert idx == ch["cycle_idx_before"] and mu == ch["mu_before"]
            tp = O.TapeReader(ch["tape"])
            r = O.slice_chain(m, pool, pool_logL, pool_logP, case["cycle"], idx, mu, tp, case["Nmcmc"],
                              case["maxmcmc"], case["logLmin"], mcmc_counter=counter, tuning_steps=10 * case["P"])
            assert tp.exhausted()
            idx, mu, counter = r["cycle_idx"], r["mu"], r["mcmc_counter"]
            assert r["steps"] == ch["steps"]
            assert r["accepted"] == round(ch["sub_acceptance"] * ch["steps"])
            assert (r["Ne"], r["Nc"], r["L"], r["R"]) == (ch["Ne"], ch["Nc"], ch["L"], ch["R"]), case["model"]
            assert close(r["out"], ch["out"], 1e-15), case["model"]      # direction norm: numpy dot vs nrm2
            assert close(r["logL"], ch["out_logL"], 1e-13)
            assert close(r["logP"], ch["out_logP"], 1e-13)
            assert r["mu"] == ch["mu_after"]
            nchains += 1
    assert nchains >= 20


def _grad_V(name):
    """Model.force of the fixture models: the gradient of the potential -log_prior."""
    if name == "gaussian_mean_sigma":
        return lambda x: np.array([0.0, 1.0 / x[1]])
    return lambda x: np.zeros(len(x))


def test_hmc_chain_golden(golden):
    """HamiltonianMonteCarloSampler.yield_sample + ConstrainedLeapFrog (sampler.py:365-402, proposal.py:793-844)
    from the reference's recorded momenta, reflection normals and uniforms."""
    mg = golden("models.json")
    nchains = 0
    for case in golden("hmc_chain.json"):
        m = _model_for(case["model"], mg[case["model"]])
        assert O.hmc_set_integration_parameters(m.bounds) == (case["base_L"], case["base_dt"])
        C = np.array(case["covariance"])
        cov = O.ensemble_covariance(np.array(case["pool"]))[1]
        assert close(C, cov, 1e-13, 1e-15)
        assert close(-np.linalg.slogdet(C)[1], case["logdeterminant"], 1e-11)
        accepted = counter = 0
        for ch in case["chains"]:
            tp = O.TapeReader(ch["tape"])
            r = O.hmc_chain(m, _grad_V(case["model"]), ch["start"], ch["start_logL"], ch["start_logP"], C,
                            np.array(case["inverse_mass"]), case["logdeterminant"], ch["dt"], ch["L"], case["logLmin"], tp)
            assert r["steps"] == ch["steps"], case["model"]
            assert close(r["out"], ch["out"], 1e-12), case["model"]
            assert close(r["logL"], ch["out_logL"], 1e-12) and close(r["logP"], ch["out_logP"], 1e-12)
            assert close(r["log_J"], ch["log_J"], 1e-9, 1e-12)
            accepted += 1
            counter += r["steps"]
            assert close(accepted / counter, ch["acceptance"], 1e-15)
            scale, dt = O.hmc_update_time_step(ch["scale"], ch["acceptance"], case["base_dt"])
            assert close(scale, ch["scale_after"], 1e-15) and close(dt, ch["dt_after"], 1e-15)
            assert tp.take("randint") + case["base_L"] == ch["L_after"] and tp.exhausted()
            nchains += 1
    assert nchains >= 17


def test_ns_loop_golden(golden):
    for case in golden("ns_loop.json"):
        ns = O.NSLoopOracle(case["L0"], mode="reference")
        logLmax = max(case["L0"])
        K = case["nthreads"]
        for b in case["batches"]:
            acc = [r[2] for r in b["replacements"] if r[2] > b["logLmin"]]
            logLmax = max([logLmax] + [r[2] for r in b["replacements"]])
            # the reference evaluates `condition` with the logLmax known before the batch's recv()s
            pre = max([max(case["L0"])] + [l for bb in case["batches"][:case["batches"].index(b)] for (_, _, l) in bb["replacements"]])
            thr = ns.batch(K, acc, pre)
            assert thr == b["logLmin"]
            assert ns.state.logZ == b["logZ"] and ns.state.info == b["info"] and ns.state.logw == b["logw"]
            assert ns.iteration == b["iteration"]
            assert ns.insertion_indices[-K:] == b["insertion_indices"]
            if b["live_logL"] is not None:
                assert ns.keys == b["live_logL"]
        rect, info, fin = ns.finish()
        assert rect == case["final_logZ_rect"] and info == case["final_info"] and fin == case["final_logZ"]
        assert ns.nested_logL == case["nested_logL"]


def test_survivor_rank_equals_reference_index_for_k1(golden):
    case = golden("ns_loop.json")[0]
    assert case["nthreads"] == 1
    ns = O.NSLoopOracle(case["L0"], mode="reference")
    for b in case["batches"]:
        acc = [r[2] for r in b["replacements"] if r[2] > b["logLmin"]]
        ranks = ns.survivor_ranks(1, acc)
        ns.batch(1, acc, 0.0)
        assert ranks == b["insertion_indices"]


def test_sequential_mode_matches_final_sweep_rule():
    """mode='sequential' over a whole run == the reference integrator fed one dead point at
    a time with the shrinking live count (NestedSampling.py:474-476)."""
    rs = np.random.RandomState(0)
    L = np.sort(rs.normal(-10, 3, 64))
    a = O.NSLoopOracle(L, mode="sequential")
    a.batch(16, list(L[-1] + 1 + np.arange(16.0)), 0.0)
    st = O.IntegralState(64)
    for i in range(16):
        st.increment(L[i], nlive=64 - i)
    assert a.state.logZ == st.logZ and a.state.info == st.info and a.state.logw == st.logw


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_recorded_reference_runs_are_sane():
    """tests/golden/ref_runs.json (complete runs of the unmodified reference, tests/golden/gen_ref_runs.py) against what is
    known analytically about its models: the fixture the GPU runs are compared with is itself checked."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_runs.json")) as f:
        runs = {r["name"]: r for r in json.load(f)}
    assert set(runs) == {"gaussian2d_mh", "ring_mh", "gaussian_mean_sigma_mh", "gaussian2d_slice",
                         "gaussian10d_mh", "gaussian50d_mh", "gaussian2d_hmc"}
    anchors = {"gaussian2d_mh": -2 * math.log(20.0), "gaussian2d_slice": -2 * math.log(20.0), "ring_mh": -2.318358,
               # normalised 1/sigma prior on [0.05, 1] x uniform mean on [-5, 5]: Z = 1/10 up to the tails beyond |mean| = 5
               "gaussian_mean_sigma_mh": -math.log(10.0),
               "gaussian10d_mh": -10 * math.log(20.0), "gaussian50d_mh": -50 * math.log(20.0),
               "gaussian2d_hmc": -2 * math.log(20.0)}
    # how far the reference's own run may sit from the analytic value, in its own sigma = sqrt(H/nlive): 3 everywhere except
    # the HMC run (spline-estimated reflection normals; recorded at 5.7 sigma) — the GPU test of that
    # fixture asserts the same fact and compares the device run with the analytic value instead
    allowed = {"gaussian2d_hmc": 7.0}
    for name, r in runs.items():
        sigma = math.sqrt(r["info"] / r["kwargs"]["nlive"])
        assert abs(r["logZ"] - anchors[name]) < allowed.get(name, 3.0) * sigma, (name, r["logZ"], anchors[name], sigma)
        assert r["n_posterior"] > 1500 and all(len(v) == 1500 for v in r["posterior"].values())
    # information of an iid unit Gaussian in [-10, 10]^D: H = D (ln 20 - (1 + ln 2 pi) / 2)
    for name, d in (("gaussian10d_mh", 10), ("gaussian50d_mh", 50)):
        h = d * (math.log(20.0) - 0.5 * (1.0 + math.log(2.0 * math.pi)))
        assert abs(runs[name]["info"] - h) < 0.1 * h, (name, runs[name]["info"], h)
        assert all(abs(runs[name]["posterior_var"][n] - 1.0) < 0.25 for n in runs[name]["names"][:8])
    g = runs["gaussian2d_mh"]
    assert all(abs(g["posterior_mean"][n]) < 0.1 and abs(g["posterior_var"][n] - 1.0) < 0.1 for n in ("x", "y"))
    ms = runs["gaussian_mean_sigma_mh"]
    # p(sigma) ~ 1/sigma: E[sigma] = 0.95 / ln 20, Var(mean) = E[sigma^2] = (1 - 0.05^2) / (2 ln 20)
    assert abs(ms["posterior_mean"]["sigma"] - 0.95 / math.log(20.0)) < 0.02
    assert abs(ms["posterior_var"]["mean"] - (1 - 0.0025) / (2 * math.log(20.0))) < 0.02


def test_production_draw_excludes_the_chains_own_slot():
    """oracle.production_draw with self_slot (mh_kernel.cuh DrawFeed::refill): the ensemble members of a step are drawn from the
    slots other than the one the chain started from (the reference pops a chain's point from the pool, sampler.py:325); distinct
    where the reference samples without replacement; the mapping from the compact index is monotone, so without a self slot the
    same Philox words give the un-shifted indices."""
    n, D, seed = 37, 5, 99
    scale = np.ones(D)
    for kind in (O.WALK, O.STRETCH, O.DE, O.EIGEN):
        seen = set()
        for step in range(400):
            self_slot = step % n
            a = O.production_draw(seed, 12, step, kind, n, D, scale, self_slot=self_slot)
            b = O.production_draw(seed, 12, step, kind, n - 1, D, scale)           # the same words over n - 1 slots, no exclusion
            used = {O.WALK: 3, O.STRETCH: 1, O.DE: 2, O.EIGEN: 1}[kind]
            idx = [int(v) for v in a[1:1 + used]]
            if kind == O.EIGEN:
                assert idx == [int(b[1])] and 0 <= idx[0] < D
                continue
            assert all(0 <= i < n and i != self_slot for i in idx), (kind, step, idx, self_slot)
            assert len(set(idx)) == used
            assert idx == [int(v) + (int(v) >= self_slot) for v in b[1:1 + used]]
            assert a[4:] == b[4:]                                                    # scalars and threshold do not depend on it
            seen.update(idx)
        if kind != O.EIGEN:
            assert len(seen) == n                                                    # every slot is reachable


def test_hmc_production_chain_restatement_is_self_consistent():
    """oracle.hmc_production_chain (the restatement the production HMC kernel is stepped against on the GPU): under a flat prior
    the metric reflection conserves the kinetic energy, so log_J = 0 and a trajectory is accepted exactly when the drawn point lies
    above the threshold; the counts follow the rule of sampler.py:347; every likelihood call is counted."""
    m = O.builtin_model("gaussian2d")
    rs = np.random.RandomState(3)
    C = np.array([[1.3, 0.4], [0.4, 0.7]])
    start = np.array([0.3, -0.2])
    logLmin = m.log_likelihood(np.array([1.2, 0.0]))                                  # the disc |x| < 1.2
    grad_logL = lambda q: -np.asarray(q)
    grad_V = lambda q: np.zeros(2)
    trajs = []
    for a in range(9):
        L = 20 + a
        trajs.append(dict(L=L, p0=np.linalg.solve(np.linalg.cholesky(C).T, rs.normal(size=2)), u=rs.uniform(size=2 * L - 1), u_acc=0.999999))
    r = O.hmc_production_chain(m, grad_logL, grad_V, True, start, m.log_likelihood(start), 0.0, C, 0.15, logLmin, trajs, nmcmc=4, max_traj=9)
    assert r["steps"] >= 4 and r["accepted"] >= 1 and r["steps"] <= 9
    assert r["evals"] == sum(2 * t["L"] for t in trajs[:r["steps"]])
    assert r["reflections"] > 0                                                       # dt * L spans the disc several times
    assert r["logL"] > logLmin and abs(r["logL"] - m.log_likelihood(r["out"])) < 1e-12
    # energy conservation, seen from outside: with u_acc ~ 1 a trajectory is rejected only through the likelihood constraint,
    # i.e. accepted + (draws that fell outside the disc) = steps; rerun with the threshold removed: every trajectory is accepted
    r2 = O.hmc_production_chain(m, grad_logL, grad_V, True, start, m.log_likelihood(start), 0.0, C, 0.15, -np.inf, trajs, nmcmc=9, max_traj=9)
    assert r2["accepted"] == r2["steps"] == 9 and r2["reflections"] == 0


def test_production_slice_draw_follows_the_philox_blocks():
    """oracle.production_slice_draw / production_slice_uniform (slice_kernel.cuh): which Philox block and word feeds which draw."""
    seed, cid, s, n, D = 7, 3, 11, 50, 4
    key = (seed, 0)
    r0 = O.philox4x32_10((cid, 0, s, 0), key)
    uL, (i0, i1), eexp, uJ = O.production_slice_draw(seed, cid, s, O.SLICE_DIFF, n, D)
    assert uL == O.u32_to_unit_open(r0[0]) and uJ == O.u32_to_unit_open(r0[1])
    assert eexp == -math.log(O.u32_to_unit_open(r0[2])) and i0 == O.u32_to_index(r0[3], n)
    assert i1 != i0 and 0 <= i1 < n
    _, z, _, _ = O.production_slice_draw(seed, cid, s, O.SLICE_GAUSS, n, D)
    assert z.shape == (D,) and z[2] == O.box_muller_f32(*O.philox4x32_10((cid, 0, s, 0x80000000 | 2), key)[:2])[0]
    mean, scale, V = np.arange(D) * 0.1, np.array([0.5, 1.0, 2.0, 0.0]), np.eye(D)
    _, v, _, _ = O.production_slice_draw(seed, cid, s, O.SLICE_CORR, n, D, mean, scale, V)
    np.testing.assert_allclose(v, mean + z * scale, rtol=1e-15)
    for k in (0, 3, 4, 9):
        assert O.production_slice_uniform(seed, cid, s, k) == O.u32_to_unit_open(O.philox4x32_10((cid, 0, s, 0x40000000 | (k >> 2)), key)[k & 3])
