"""GPU parity and end-to-end tests of the slice sampler kernel (SliceSampler.yield_sample,
cpnest/sampler.py:414-569, with EnsembleSliceProposalCycle, cpnest/proposal.py:114-207), through the C-ABI.

Replay: the chains of tests/golden/slice_chain.json (recorded from the unmodified reference) are
re-run on the device from the same random numbers: integer results (slices, accepted slices,
likelihood calls, expansion/contraction counts) exactly, floating point within 1e-12 relative.
"""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import cpnest_oracle as O
from test_gpu_parity import engine_for, rel_close, RTOL

pytestmark = pytest.mark.gpu

REPLAY_LANES = {2: [1], 10: [4, 8, 16, 32], 50: [8, 16, 32]}


def flatten_tape(events):
    """Fixture event list -> the flat sequence cpn_replay_slice consumes (include/cpnest_b200.h)."""
    out = []
    for kind, v in events:
        if kind in ("np_uniform", "np_exponential"):
            out.append(v)
        elif kind in ("sample", "np_normal", "np_mvn"):
            out.extend(v)
        else:
            raise AssertionError(kind)
    return out


def pop_start(pool, pool_logL, pool_logP, logLmin):
    """sampler.py:471-478: rotate the pool until a member above the threshold comes first."""
    j = 0
    while j < len(pool) and not (pool_logL[0] > logLmin):
        pool.append(pool.pop(0)); pool_logL.append(pool_logL.pop(0)); pool_logP.append(pool_logP.pop(0))
        j += 1


def test_slice_chain_replay_matches_reference(golden):
    mg = golden("models.json")
    nchains = 0
    for case in golden("slice_chain.json"):
        name, D = case["model"], case["D"]
        for lanes in REPLAY_LANES[D]:
            m, e = engine_for(name, mg[name], lanes=lanes, maxmcmc=case["maxmcmc"], sampler="slice")
            with e:
                e.set_slice_cycle(case["cycle"])
                pool = [np.array(p) for p in case["pool"]]
                pool_logL = list(case["pool_logL"])
                pool_logP = list(case["pool_logP"])
                idx, mu, counter = 0, case["mu0"], 0
                for ch in case["chains"]:
                    pop_start(pool, pool_logL, pool_logP, case["logLmin"])
                    start = np.concatenate((pool[0], [pool_logL[0], pool_logP[0]]))
                    ens = np.array(pool[1:])                     # the deque the directions are drawn from
                    out, steps, acc, evals, last = e.replay_slice(ens, start, idx, [flatten_tape(ch["tape"])], mu,
                                                                  case["Nmcmc"], case["maxmcmc"], case["logLmin"], compat=True)
                    r = O.slice_chain(m, pool, pool_logL, pool_logP, case["cycle"], idx, mu, ch["tape"], case["Nmcmc"],
                                      case["maxmcmc"], case["logLmin"], mcmc_counter=counter, tuning_steps=10 * case["P"])
                    idx, mu, counter = r["cycle_idx"], r["mu"], r["mcmc_counter"]
                    assert steps[0] == ch["steps"], (name, lanes)
                    assert acc[0] == round(ch["sub_acceptance"] * ch["steps"]), (name, lanes)
                    assert evals[0] == r["evals"], (name, lanes)
                    assert (last[0, 0], last[0, 1]) == (ch["Ne"], ch["Nc"]), (name, lanes)
                    assert rel_close(last[0, 2:], [ch["L"], ch["R"]], RTOL), (name, lanes)
                    assert rel_close(out[0, :D], ch["out"], RTOL), (name, lanes, out[0, :D], ch["out"])
                    assert rel_close(out[0, D], ch["out_logL"], RTOL)
                    assert rel_close(out[0, D + 1], ch["out_logP"], RTOL)
                    nchains += 1
    assert nchains >= 36


def test_slice_replay_many_chains_in_one_launch(golden):
    """All chains of a fixture case at once (chains of different length share warps)."""
    mg = golden("models.json")
    case = [c for c in golden("slice_chain.json") if c["model"] == "gaussian10d"][0]
    m, e = engine_for("gaussian10d", mg["gaussian10d"], lanes=8, maxmcmc=case["maxmcmc"], sampler="slice")
    with e:
        e.set_slice_cycle(case["cycle"])
        # every chain replayed from the SAME pool state is only valid for the first; so replay chain 0 five times
        pool = [np.array(p) for p in case["pool"]]
        pool_logL, pool_logP = list(case["pool_logL"]), list(case["pool_logP"])
        pop_start(pool, pool_logL, pool_logP, case["logLmin"])
        start = np.concatenate((pool[0], [pool_logL[0], pool_logP[0]]))
        ch = case["chains"][0]
        n = 5
        out, steps, acc, evals, last = e.replay_slice(np.array(pool[1:]), np.tile(start, (n, 1)), 0,
                                                      [flatten_tape(ch["tape"])] * n, case["mu0"], case["Nmcmc"],
                                                      case["maxmcmc"], case["logLmin"])
        assert np.all(steps == ch["steps"])
        for i in range(n):
            assert rel_close(out[i, :10], ch["out"], RTOL)


def test_production_slice_directions():
    """The three direction proposals of the production kernel (Philox draws): differential directions are
    differences of two distinct live points times mu, Gaussian directions have norm mu, correlated-Gaussian
    directions follow mu * N(mean, cov) of the live ensemble (proposal.py:121-187)."""
    from cpnest_b200.engine import Engine
    from cpnest_b200 import _lib as L
    D, N, mu = 6, 4096, 0.7
    rs = np.random.RandomState(5)
    A = rs.normal(size=(D, D))
    X = rs.normal(size=(N, D)) @ A.T * 0.3 + rs.uniform(-1, 1, D)
    X = np.clip(X, -9.9, 9.9)
    rows = np.concatenate((X, np.zeros((N, 2))), axis=1)
    with Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=50, seed=11, sampler="slice") as e:
        e.set_live(rows)
        g = e.slice_directions(L.SLICE_GAUSS, N, mu)
        assert np.allclose(np.linalg.norm(g, axis=1), mu, rtol=1e-12)
        assert np.all(np.abs(g.mean(0)) < 5 * mu / math.sqrt(D * N))
        from scipy.spatial import cKDTree
        tree = cKDTree(X)
        d = e.slice_directions(L.SLICE_DIFF, 64, mu) / mu
        for v in d:
            # v = X[i] - X[j] for some pair i != j: X[j] + v must be a live point for some j
            dist, _ = tree.query(X + v)
            assert dist.min() < 1e-12 and np.linalg.norm(v) > 0
        c = e.slice_directions(L.SLICE_CORR, N, mu) / mu
        mean, cov = X.mean(0), np.cov(X.T)
        se = np.sqrt(np.diag(cov) / N)
        assert np.all(np.abs(c.mean(0) - mean) < 5 * se)
        assert np.allclose(np.cov(c.T), cov, rtol=0.2, atol=0.05 * np.max(np.diag(cov)))


def run_slice(functor, bounds, nlive, K, maxmcmc=200, seed=1234, params=None, **kw):
    from cpnest_b200.engine import Engine
    e = Engine(functor, bounds, nlive=nlive, maxmcmc=maxmcmc, seed=seed, params=params, sampler="slice", **kw)
    e.init_prior()
    nb = e.run(K, tolerance=0.1)
    ins = e.insertion_indices()
    rect, trap = e.finalise()
    return e, e.state(), trap, ins, nb


def check_logz(s, logz, analytic, nlive, nsigma=3.0):
    sigma = math.sqrt(s["info"] / nlive)
    assert abs(logz - analytic) < nsigma * sigma, "logZ %.4f vs analytic %.4f (sigma %.4f, H %.2f)" % (
        logz, analytic, sigma, s["info"])


def test_slice_gaussian2d_logz():
    # tests/test_2d_slice.py: 2-D unit Gaussian with nslice=1
    e, s, logz, ins, nb = run_slice("gaussian_iid", [[-10, 10]] * 2, nlive=1000, K=64)
    with e:
        check_logz(s, logz, -2 * math.log(20.0), 1000)
        assert stats.kstest(ins, "uniform").pvalue > 1e-3
        assert s["slice_mu"] > 0 and s["total_evals"] > 0


def test_slice_ring_logz():
    # tests/test_ring_slice.py
    e, s, logz, ins, nb = run_slice("ring", [[-2, 2]] * 2, nlive=1000, K=64, params=dict(radius=1.0, thickness=0.1))
    with e:
        check_logz(s, logz, -2.318358, 1000)
        assert stats.kstest(ins, "uniform").pvalue > 1e-3


def test_slice_nonflat_prior_logz():
    # (mean, sigma) model with the 1/sigma prior: stepping out is bounded by the prior slice level
    from scipy import integrate
    f = lambda m, s: math.exp(-0.5 * m * m / (s * s)) / (s * math.sqrt(2 * math.pi)) / s
    Z, _ = integrate.dblquad(f, 0.05, 1.0, lambda s: -5, lambda s: 5, epsabs=1e-10)
    Z /= 10.0 * math.log(1.0 / 0.05)
    e, s, logz, ins, nb = run_slice("gaussian_mean_sigma", [[-5, 5], [0.05, 1]], nlive=1000, K=64)
    with e:
        check_logz(s, logz, math.log(Z), 1000)


def test_slice_rosenbrock10d_runs_and_gaussian10d_logz():
    # config 3 (BASELINE.json): 10-D, slice sampler.  The 10-D Gaussian has the analytic anchor.
    e, s, logz, ins, nb = run_slice("gaussian_iid", [[-10, 10]] * 10, nlive=2000, K=128, maxmcmc=500)
    with e:
        check_logz(s, logz, -10 * math.log(20.0), 2000)
        assert stats.kstest(ins[-10000:], "uniform").pvalue > 1e-3
    e, s, logz, ins, nb = run_slice("rosenbrock", [[-5, 5]] * 10, nlive=2000, K=128, maxmcmc=500)
    with e:
        live_dead = e.dead()
        assert np.isfinite(logz) and s["info"] > 10
        assert np.all(np.diff(live_dead[:, 10]) >= 0)
        assert stats.kstest(ins[-10000:], "uniform").pvalue > 1e-4


def test_slice_is_deterministic_per_seed_and_lane_independent():
    a = run_slice("gaussian_iid", [[-10, 10]] * 4, nlive=500, K=32, seed=7, lanes=1)
    b = run_slice("gaussian_iid", [[-10, 10]] * 4, nlive=500, K=32, seed=7, lanes=4)
    c = run_slice("gaussian_iid", [[-10, 10]] * 4, nlive=500, K=32, seed=8, lanes=4)
    # Same Philox streams whatever the lane layout; the trajectories still separate after a while: lane-order
    # dependent rounding (direction norm, likelihood sum) perturbs the live set in the last bit, and the
    # eigenvectors behind the correlated-Gaussian direction are ill-conditioned functions of it (nearly
    # degenerate eigenvalues for an isotropic target).  The evidence must agree statistically.
    sigma = math.sqrt(a[1]["info"] / 500)
    assert abs(a[2] - b[2]) < 3 * sigma and abs(a[2] + 4 * math.log(20.0)) < 3 * sigma
    a2 = run_slice("gaussian_iid", [[-10, 10]] * 4, nlive=500, K=32, seed=7, lanes=1)
    assert np.array_equal(a[0].dead(), a2[0].dead()) and a[2] == a2[2]      # bit-reproducible per seed and layout
    a2[0].close()
    assert a[2] != c[2]
    for r in (a, b, c):
        r[0].close()
