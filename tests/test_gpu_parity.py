"""GPU parity tests: the CUDA path (through the C-ABI, via ctypes) against the CPU oracle and
the golden fixtures generated from the unmodified reference.  Run on the B200 box: -m gpu.

Tolerances (BASELINE.json north_star): deterministic pieces agree with the reference on
identical inputs and random numbers within 1e-12 relative in fp64; integer/index results
(steps, accepted counts, ranks, orders) bit-exactly.
"""
import math

import numpy as np
import pytest

from oracle import cpnest_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def rel_close(a, b, rtol=RTOL, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    both_nan = np.isnan(a) & np.isnan(b)
    with np.errstate(invalid="ignore"):
        ok = same_inf | both_nan | (np.abs(a - b) <= atol + rtol * np.abs(b))
    return bool(np.all(ok))


def engine_for(name, case, nlive=64, **kw):
    from cpnest_b200.engine import Engine
    if name == "gaussian_mixture":
        m = O.builtin_model(name, data=case["data"])
    else:
        m = O.builtin_model(name)
    params = {}
    if name == "gaussian_mixture":
        params = dict(data=case["data"])
    return m, Engine(m.functor, m.bounds, nlive=nlive, params=params, **kw)


def test_library_exports_and_device():
    from cpnest_b200 import _lib
    lib = _lib.load()
    assert lib.cpn_device_count() >= 1


def test_philox_known_answers_on_device():
    from cpnest_b200.engine import Engine
    with Engine("gaussian_iid", [[-1, 1]] * 2, nlive=8) as e:
        assert e.philox([0, 0, 0, 0], [0, 0]).tolist() == [[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]]
        assert e.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [[0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]]
        assert e.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
            [[0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]]
        rs = np.random.RandomState(0)
        ctr = rs.randint(0, 2 ** 32, size=(257, 4), dtype=np.uint64).astype(np.uint32)
        key = rs.randint(0, 2 ** 32, size=2, dtype=np.uint64).astype(np.uint32)
        out = e.philox(ctr, key)
        for i in range(0, 257, 16):
            assert out[i].tolist() == O.philox4x32_10(ctr[i], key)


LANES = {"gaussian2d": [1], "gaussian1d": [1], "half_gaussian": [1], "eggbox": [1, 4, 32], "ring": [1, 8],
         "gaussian_mean_sigma": [1, 8], "rosenbrock2d": [1, 4], "rosenbrock10d": [4, 8, 16, 32],
         "gaussian10d": [4, 8, 16, 32], "gaussian50d": [8, 16, 32], "gaussian_mixture": [32, 8, 1]}


def test_device_functors_match_reference_models(golden):
    """Model.log_prior / log_likelihood: device functor vs the values the reference's own Python
    callables returned (tests/golden/models.json), 1e-12 relative."""
    mg = golden("models.json")
    for name, case in mg.items():
        for lanes in LANES[name]:
            m, e = engine_for(name, case, lanes=lanes)
            with e:
                X = np.array(case["X"])
                logL, logP = e.eval_model(X)
                assert rel_close(logP, case["logP"], 1e-14, 1e-15), (name, lanes)
                ref = np.array(case["logL"])
                inb = np.array(case["in_bounds"])
                tol = 3e-12 if name == "gaussian_mixture" else RTOL
                assert rel_close(logL[inb], ref[inb], tol, 1e-13), (name, lanes, np.max(np.abs(logL[inb] - ref[inb])))


def test_device_functors_random_points_vs_oracle():
    """10^4 random in-bounds points per model."""
    from cpnest_b200.engine import Engine
    rs = np.random.RandomState(3)
    specs = [("gaussian50d", {}), ("rosenbrock10d", {}), ("eggbox", {}), ("ackley10d", {}), ("ring", {})]
    for name, params in specs:
        m = O.builtin_model(name)
        with Engine(m.functor, m.bounds, nlive=16, params=params) as e:
            X = m.lo + (m.hi - m.lo) * rs.uniform(size=(10000, m.D))
            logL, logP = e.eval_model(X)
            ref = np.array([m.log_likelihood(x) for x in X[:2000]])
            assert rel_close(logL[:2000], ref, RTOL, 1e-12), name
            assert np.all(logP == 0.0)
    m = O.make_correlated_gaussian(50)
    with Engine("gaussian_corr", m.bounds, nlive=16, params=m.params) as e:
        X = rs.normal(size=(500, 50)) * 0.5
        logL, _ = e.eval_model(X)
        ref = np.array([m.log_likelihood(x) for x in X])
        assert rel_close(logL, ref, RTOL), np.max(np.abs(logL - ref))


def build_tape(cycle, cycle_idx, events, maxmcmc):
    """Event list of the fixture -> [maxmcmc][8] tape rows {i0,i1,i2,g0,g1,g2,u_stretch,u_mh}."""
    tp = O.TapeReader(events)
    tape = np.zeros((maxmcmc, 8))
    s = 0
    idx = cycle_idx
    while not tp.exhausted():
        idx = (idx + 1) % len(cycle)
        k = cycle[idx]
        row = tape[s]
        if k == O.WALK:
            row[0:3] = tp.take("sample")
            row[3:6] = [tp.take("gauss") for _ in range(3)]
        elif k == O.STRETCH:
            row[0] = tp.take("choice")
            row[6] = tp.take("uniform")
        elif k == O.DE:
            row[0:2] = tp.take("sample")
            row[3] = tp.take("gauss")
        else:
            row[0] = tp.take("randrange")
            row[3] = tp.take("gauss")
        row[7] = tp.take("random")
        s += 1
    return tape, s


REPLAY_LANES = {2: [1], 10: [4, 8, 16, 32], 50: [8, 16, 32]}


def test_mh_chain_replay_matches_reference(golden):
    """MetropolisHastingsSampler.yield_sample + the four proposals, driven by the random tape the
    reference consumed: final point, every intermediate chain state, step and acceptance counts."""
    mg = golden("models.json")
    nchains = 0
    for case in golden("mh_chain.json"):
        name = case["model"]
        D = case["D"]
        for lanes in REPLAY_LANES[D]:
            m, e = engine_for(name, mg[name], lanes=lanes, maxmcmc=case["maxmcmc"])
            with e:
                e.set_cycle(case["cycle"])
                pool = [np.array(p) for p in case["pool"]]
                pool_logL = list(case["pool_logL"])
                pool_logP = list(case["pool_logP"])
                idx = 0
                for ch in case["chains"]:
                    start = np.concatenate((pool[0], [pool_logL[0], m.log_prior(pool[0])]))
                    ens = np.array(pool[1:])                      # the deque after popleft (sampler.py:328)
                    tape, nsteps = build_tape(case["cycle"], idx, ch["tape"], case["maxmcmc"])
                    out, steps, acc, states = e.replay_mh(ens, start, case["eigen_values"], case["eigen_vectors"], idx,
                                                          tape, case["Nmcmc"], case["maxmcmc"], case["logLmin"],
                                                          compat=True, record=True)
                    # advance the pool with the oracle (already pinned to the same fixture)
                    r = O.mh_chain(m, pool, pool_logL, pool_logP, case["cycle"], idx, case["eigen_values"],
                                   case["eigen_vectors"], ch["tape"], case["Nmcmc"], case["maxmcmc"], case["logLmin"])
                    idx = r["cycle_idx"]
                    assert steps[0] == ch["steps"] == nsteps, (name, lanes)
                    assert acc[0] == round(ch["sub_acceptance"] * ch["steps"]), (name, lanes)
                    assert rel_close(out[0, :D], ch["out"], RTOL), (name, lanes)
                    assert rel_close(out[0, D], ch["out_logL"], RTOL)
                    assert rel_close(out[0, D + 1], ch["out_logP"], RTOL)
                    assert rel_close(states[0, :steps[0]], np.array(ch["states"]), RTOL), (name, lanes)
                    nchains += 1
    assert nchains >= 30


def test_replay_without_fp32_rounding_differs_from_reference(golden):
    """The reference rounds LivePoint scalars to fp32 (parameter.pyx:96-120); compat=0 is the
    production arithmetic and must NOT reproduce those numbers bit-for-bit but stay close."""
    mg = golden("models.json")
    case = golden("mh_chain.json")[0]
    m, e = engine_for(case["model"], mg[case["model"]], maxmcmc=case["maxmcmc"])
    with e:
        e.set_cycle(case["cycle"])
        pool = [np.array(p) for p in case["pool"]]
        ch = case["chains"][0]
        start = np.concatenate((pool[0], [case["pool_logL"][0], 0.0]))
        tape, _ = build_tape(case["cycle"], 0, ch["tape"], case["maxmcmc"])
        a = e.replay_mh(np.array(pool[1:]), start, case["eigen_values"], case["eigen_vectors"], 0, tape,
                        case["Nmcmc"], case["maxmcmc"], case["logLmin"], compat=True)[0]
        b = e.replay_mh(np.array(pool[1:]), start, case["eigen_values"], case["eigen_vectors"], 0, tape,
                        case["Nmcmc"], case["maxmcmc"], case["logLmin"], compat=False)[0]
        assert not np.array_equal(a, b)
        assert rel_close(a, b, 1e-5, 1e-5)


def test_integrator_matches_reference(golden):
    """_NSintegralState.increment / finalise on the device (both accumulation rules)."""
    from cpnest_b200.engine import Engine
    with Engine("gaussian_iid", [[-1, 1]] * 2, nlive=1000) as e:
        for case in golden("integrator.json"):
            e.ns_reset()
            st = O.IntegralState(case["nlive"])
            for (logL, nl, nrep), (logZ, info, logw) in zip(case["steps"], case["trace"]):
                # the reference's call increment(logL, nlive=nl, nreplace=nrep): logt = -nrep/nl
                e.ns_increment([logL] * 1 if nrep == 1 else [-np.inf] * (nrep - 1) + [logL],
                               nl or case["nlive"], mode="reference" if nrep > 1 else "sequential")
            s = e.state()
            logZ, info, logw = case["trace"][-1]
            assert rel_close(s["logZ"], logZ, RTOL), case["tag"]
            assert rel_close(s["info"], info, 1e-11, 1e-12), case["tag"]
            assert rel_close(s["logw"], logw, RTOL), case["tag"]
            if "finalised_logZ" in case:
                assert rel_close(e.ns_trapezoid(), case["finalised_logZ"], RTOL), case["tag"]


def test_integrator_batched_sequential_equals_pointwise(golden):
    """A batch of K dead points accumulated in one launch (parallel scan + log-sum-exp) equals the
    reference integrator fed one point at a time with the shrinking live count
    (NestedSampling.py:474-476)."""
    from cpnest_b200.engine import Engine
    rs = np.random.RandomState(1)
    with Engine("gaussian_iid", [[-1, 1]] * 2, nlive=5000) as e:
        for N, K, nb in [(64, 16, 3), (5000, 2048, 4), (5000, 3000, 2)]:
            e.ns_reset()
            st = O.IntegralState(N)
            base = -200.0
            for b in range(nb):
                L = np.sort(rs.normal(base, 5.0, K))
                base += 12.0
                e.ns_increment(L, N, mode="sequential")
                for i, l in enumerate(L):
                    st.increment(l, nlive=N - i)
                s = e.state()
                assert rel_close(s["logZ"], st.logZ, RTOL)
                assert rel_close(s["info"], st.info, 1e-11)
                assert rel_close(s["logw"], st.logw, RTOL)
            assert rel_close(e.ns_trapezoid(), st.finalise(), RTOL)


def test_ensemble_covariance_and_eigen(golden):
    """EnsembleEigenVector.update_eigenvectors: np.cov + np.linalg.eigh on the device."""
    from cpnest_b200.engine import Engine
    for case in golden("proposals.json"):
        X = np.array(case["ensemble"])
        P, D = X.shape
        with Engine("gaussian_iid", [[-100, 100]] * D, nlive=P) as e:
            rows = np.concatenate((X, np.arange(P)[:, None] * 1.0, np.zeros((P, 1))), axis=1)
            e.set_live(rows)
            mean, cov, w, v = e.ensemble_stats()
            assert rel_close(mean, X.mean(0), 1e-13, 1e-15)
            assert rel_close(cov, np.atleast_2d(case["covariance"]), 1e-12, 1e-14), D
            assert rel_close(w, case["eigen_values"], 1e-10, 1e-13), D
            V = np.atleast_2d(case["eigen_vectors"])
            for i in range(D):
                s = np.sign(np.dot(V[:, i], v[:, i]))
                assert rel_close(s * v[:, i], V[:, i], 1e-7, 1e-8), (D, i)
            # decomposition residual independent of LAPACK's conventions
            assert np.max(np.abs(cov @ v - v * w)) < 1e-12 * max(1.0, np.max(np.abs(w)))
    # more than 32 slices of the ensemble (api.cu stat_blocks: ~320 points per slice beyond nlive = 1e4)
    rs = np.random.RandomState(6)
    X = rs.normal(size=(40123, 20)) @ rs.normal(size=(20, 20))
    with Engine("gaussian_iid", [[-1e3, 1e3]] * 20, nlive=40123) as e:
        e.set_live(np.concatenate((X, np.arange(40123.0)[:, None], np.zeros((40123, 1))), axis=1))
        mean, cov, w, v = e.ensemble_stats()
        assert rel_close(mean, X.mean(0), 1e-12, 1e-13)
        assert rel_close(cov, np.cov(X.T), 1e-11, 1e-12)
    rs = np.random.RandomState(5)
    X = rs.normal(size=(4000, 50)) @ rs.normal(size=(50, 50))
    with Engine("gaussian_iid", [[-1e3, 1e3]] * 50, nlive=4000) as e:
        e.set_live(np.concatenate((X, np.arange(4000.0)[:, None], np.zeros((4000, 1))), axis=1))
        mean, cov, w, v = e.ensemble_stats()
        assert rel_close(cov, np.cov(X.T), 1e-11, 1e-12)
        assert rel_close(w, np.linalg.eigvalsh(np.cov(X.T)), 1e-9, 1e-10)
        assert np.max(np.abs(v.T @ v - np.eye(50))) < 1e-12
        # warm-started refreshes (previous eigenvectors as the starting basis) on drifting ensembles
        for k in range(20):
            X = X * (1.0 + 0.05 * rs.normal(size=(1, 50))) + 0.1 * rs.normal(size=X.shape)
            e.set_live(np.concatenate((X, np.arange(4000.0)[:, None], np.zeros((4000, 1))), axis=1))
            mean, cov, w, v = e.ensemble_stats()
            ref = np.cov(X.T)
            assert rel_close(cov, ref, 1e-11, 1e-12)
            assert rel_close(w, np.linalg.eigvalsh(ref), 1e-9, 1e-10), k
            assert np.max(np.abs(v.T @ v - np.eye(50))) < 1e-11
            assert np.max(np.abs(ref @ v - v * w)) < 1e-10 * np.max(np.abs(w))


def test_ns_batch_loop_matches_reference(golden):
    """consume_sample with scripted sampler replies: threshold, evidence, information, sorted live
    keys after every batch (mode='reference'), and the survivor-rank insertion statistic."""
    from cpnest_b200.engine import Engine
    for case in golden("ns_loop.json"):
        N, K = case["nlive"], case["nthreads"]
        X0, L0 = np.array(case["X0"]), np.array(case["L0"])
        with Engine("gaussian_iid", [[-10, 10]] * 2, nlive=N, ns_mode="reference") as e:
            e.set_live(np.concatenate((X0, L0[:, None], np.zeros((N, 1))), axis=1))
            ns = O.NSLoopOracle(L0, mode="reference")
            all_ranks = []
            for b in case["batches"]:
                acc = [r for r in b["replacements"] if r[2] > b["logLmin"]]
                e.select_and_accumulate(K)
                s = e.state()
                assert s["logLmin"] == b["logLmin"]
                assert rel_close(s["logZ"], b["logZ"], RTOL)
                assert rel_close(s["info"], b["info"], 1e-11, 1e-12)
                assert rel_close(s["logw"], b["logw"], RTOL)
                rows = np.array([r[1] + [r[2], 0.0] for r in acc])
                e.stage_replacements(rows)
                all_ranks += ns.survivor_ranks(K, [r[2] for r in acc])
                e.insert_batch(K)
                ns.batch(K, [r[2] for r in acc], 0.0)
                if b["live_logL"] is not None:
                    live = e.live()
                    assert live[:, 2].tolist() == b["live_logL"]
                    assert np.all(-0.5 * (live[:, :2] ** 2).sum(1) == live[:, 2])     # rows travel with their keys
            assert e.state()["iteration"] == case["batches"][-1]["iteration"]
            assert e.insertion_indices().tolist() == all_ranks
            dead = e.dead()
            assert dead[:, 2].tolist() == case["nested_logL"][:len(dead)]
            rect, trap = e.finalise()
            assert rel_close(rect, case["final_logZ_rect"], RTOL)
            assert rel_close(trap, case["final_logZ"], RTOL)
            assert rel_close(e.state()["info"], case["final_info"], 1e-11)
            assert e.dead()[:, 2].tolist() == case["nested_logL"]


def test_ns_batch_loop_sequential_mode_vs_oracle():
    from cpnest_b200.engine import Engine
    rs = np.random.RandomState(8)
    N, K = 512, 128
    X0 = rs.uniform(-10, 10, (N, 2))
    L0 = -0.5 * (X0 ** 2).sum(1)
    with Engine("gaussian_iid", [[-10, 10]] * 2, nlive=N, ns_mode="sequential") as e:
        e.set_live(np.concatenate((X0, L0[:, None], np.zeros((N, 1))), axis=1))
        ns = O.NSLoopOracle(L0, mode="sequential")
        for b in range(12):
            thr = ns.keys[K - 1]
            r = np.sqrt(-2 * thr) * np.sqrt(rs.uniform(size=K))
            th = rs.uniform(0, 2 * np.pi, K)
            X = np.stack((r * np.cos(th), r * np.sin(th)), 1)
            Ln = -0.5 * (X ** 2).sum(1)
            e.select_and_accumulate(K)
            e.stage_replacements(np.concatenate((X, Ln[:, None], np.zeros((K, 1))), axis=1))
            ranks = ns.survivor_ranks(K, Ln)
            e.insert_batch(K)
            ns.batch(K, list(Ln), 0.0)
            s = e.state()
            assert rel_close(s["logZ"], ns.state.logZ, RTOL) and rel_close(s["info"], ns.state.info, 1e-11)
            assert e.live()[:, 2].tolist() == ns.keys
            assert e.insertion_indices()[-K:].tolist() == ranks
        rect, trap = e.finalise()
        orect, oinfo, otrap = ns.finish()
        assert rel_close(rect, orect, RTOL) and rel_close(trap, otrap, RTOL)
        rows, lv = e.dead(log_vols=True)
        assert rows[:, 2].tolist() == ns.nested_logL
        assert rel_close(lv, ns.state.log_vols[1:], RTOL)


def test_device_chain_length_rule_equals_estimate_nmcmc_on_the_fly():
    """Sampler.estimate_nmcmc_on_the_fly (cpnest/sampler.py:156-176) - the moving average of safety * (2/acc - 1) the reference
    applies chain after chain - against the device controller of the host protocol (k_adapt_nmcmc_on_the_fly: the K sequential
    updates of a batch as one ordered reduction of clamped affine maps), including chains that made no move (acc = 0)."""
    from cpnest_b200.engine import Engine
    m = O.builtin_model("gaussian10d")
    N, K, maxmcmc = 256, 200, 60
    rs = np.random.RandomState(8)
    with Engine(m.functor, m.bounds, nlive=N, maxmcmc=maxmcmc, nmcmc_init=20, nmcmc_safety=5.0, adapt_nmcmc=True, seed=5) as e:
        e.set_cycle(O.make_cycle(np.random.RandomState(5)))
        e.init_prior()
        live = e.live()
        ne = e.state()["nmcmc_exact"]
        saw_immobile = False
        for q in (0.3, 0.6, None, 0.5):              # None: a threshold no proposal can beat - every chain returns unmoved
            logLmin = float(np.quantile(live[:, 10], q)) if q is not None else 1e9
            good = np.where(live[:, 10] > (logLmin if q is not None else -np.inf))[0]
            req = np.ascontiguousarray(live[good[rs.randint(0, len(good), size=K)]])
            n_chain = e.state()["nmcmc"]
            _, steps, acc = e.evolve_host(req, logLmin, n_chain)
            for j in range(K):
                ne, n = O.estimate_nmcmc_on_the_fly(ne, acc[j] / steps[j], maxmcmc, poolsize=N, safety=5.0)
            saw_immobile |= bool((acc == 0).any())
            s = e.state()
            assert rel_close(s["nmcmc_exact"], ne, rtol=1e-12), (s["nmcmc_exact"], ne)
            assert s["nmcmc"] == n
        assert saw_immobile, "no chain without a move: the acc = 0 branch was not exercised"


def test_chain_statistics_equal_the_correlations_of_the_recorded_chains():
    """k_chain_corr_partial / _final + the controller's de-biased maximum (ns_kernels.cuh) against numpy: the chain history
    records where every chain of a batch started, cpn_get_last_batch where it ended; corr(start, end) per coordinate and for
    logL across the K chains, rho_max = sqrt(max_d(rho_d^2 - 1/K)) on the first batch (no moving average yet).  K = 1300 is
    two full 512-chain segments and a ragged one; D + 1 = 41 coordinates are a full and a ragged group of 32 lanes."""
    import ctypes as C
    from cpnest_b200.engine import Engine
    from cpnest_b200 import _lib as L
    D, N, K = 40, 3000, 1300
    m = O.builtin_model("gaussian40d")
    # (maxmcmc far above the chain length: no chain is sent again from another start, which the log could not tell from a move)
    with Engine(m.functor, m.bounds, nlive=N, maxmcmc=400, nmcmc_init=20, adapt_nmcmc=True, seed=21) as e:
        e.set_cycle(O.make_cycle(np.random.RandomState(5)))
        e.init_prior()
        cap = 160
        e.enable_chain_log(K, cap)
        e.ns_step(K)
        rows, steps, acc = e.last_batch(K)
        start = np.empty((K, D + 1))
        ev = np.empty((cap, D + 4))
        n = C.c_int32(0)
        for j in range(K):
            L.check(e.lib.cpn_get_chain_log(e.h, j, L.dptr(ev), cap, C.byref(n)))
            first = ev[0]
            assert first[0] == 0 and first[1] == -1 and n.value >= 2          # (tag 0, step -1): the chain's start point
            start[j, :D] = first[4:4 + D]
            start[j, D] = first[2]
        end = rows[:, :D + 1]
        rho = np.array([np.corrcoef(start[:, d], end[:, d])[0, 1] for d in range(D + 1)])
        expect = math.sqrt(max(np.max(rho ** 2 - 1.0 / K), 0.0))
        s = e.state()
        assert s["failed_chains"] == 0 and steps.max() < 400 and acc.min() > 0
        assert rel_close(s["rho_max"], expect, 1e-9), (s["rho_max"], expect)
        assert 0.0 < expect < 1.0
