"""Mixed sampler populations: CPNest(nthreads, nslice, nhamiltonian) with more than one kind > 0
(cpnest/cpnest.py:197-261 starts nthreads-nslice-nhamiltonian MH, nhamiltonian HMC and nslice slice sampler processes and
every iteration hands each of them one of the worst points, cpnest/NestedSampling.py:324-342).  On the device the K chains
of a batch are cut into one contiguous range per kind; each kind keeps its own chain-length controller."""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import cpnest_oracle as O

pytestmark = pytest.mark.gpu


def make(functor="gaussian_iid", dim=2, nlive=1000, mix=None, sampler="mh", seed=1234, maxmcmc=400, **kw):
    from cpnest_b200.engine import Engine
    e = Engine(functor, [[-10, 10]] * dim, nlive=nlive, maxmcmc=maxmcmc, seed=seed, mix=mix, sampler=sampler, **kw)
    e.set_cycle(O.make_cycle(np.random.RandomState(seed)))
    e.init_prior()
    return e


def finish(e, K):
    e.run(K, tolerance=0.1)
    ins = e.insertion_indices()
    rect, trap = e.finalise()
    return e.state(), trap, ins


@pytest.mark.parametrize("mix,K,expect", [
    (dict(mh=2, slice=1, hmc=1), 100, [50, 25, 25]),
    (dict(mh=1, slice=1), 7, [4, 3, 0]),            # largest remainder, ties to the lower kind
    (dict(slice=1, hmc=3), 64, [0, 16, 48]),
    (dict(mh=1000, slice=1, hmc=1), 10, [8, 1, 1]),  # every kind present gets at least one chain
    ((3, 3, 3), 3, [1, 1, 1]),
])
def test_batch_is_split_in_the_requested_proportions(mix, K, expect):
    with make(mix=mix, nlive=200) as e:
        e.ns_step(K)
        s = e.state()
        assert s["chains_kind"] == expect
        assert s["iteration"] == K


def test_each_kind_runs_its_own_kernel_and_controller():
    K = 96
    with make(dim=4, mix=dict(mh=1, slice=1, hmc=1), nlive=600) as e:
        for _ in range(30):
            e.ns_step(K)
        s = e.state()
        assert s["chains_kind"] == [32, 32, 32]
        rows, steps, acc = e.last_batch(K)
        n_mh, n_sl, n_hmc = s["nmcmc_kind"]
        # the controllers have moved apart: MH needs many more steps than slices or trajectories
        assert n_mh > 1.5 * n_sl and n_mh > 2 * n_hmc, s["nmcmc_kind"]
        assert all(r >= 0 for r in s["rho_kind"])
        # every point returned satisfies the constraint and lies in the box, whatever its kind
        assert np.all(rows[:, 4] > s["logLmin"]) or s["done"]
        assert np.all(np.abs(rows[:, :4]) < 10)
        assert np.allclose(rows[:, 4], -0.5 * (rows[:, :4] ** 2).sum(1) - 2 * math.log(2 * math.pi), rtol=1e-12)
        # chain lengths by range: MH chains stop after >= nmcmc steps (sampler.py:347), slice chains after > nmcmc slices
        # (sampler.py:553-557), HMC chains after >= nmcmc trajectories; the ranges differ visibly
        assert np.all(acc > 0)
        assert steps[:32].min() > steps[32:64].max() or steps[:32].mean() > 2 * steps[32:64].mean()
        assert s["slice_mu"] != 1.0                   # the slice range adapted the length scale
        assert s["total_steps"] > 0 and s["total_evals"] > 0


def test_pure_population_through_mix_equals_sampler_argument():
    out = []
    for kw in (dict(sampler="slice"), dict(mix=dict(slice=5)), dict(sampler="mh", mix=(0, 2, 0))):
        with make(dim=3, nlive=400, **kw) as e:
            for _ in range(10):
                e.ns_step(50)
            s = e.state()
            out.append((s["logZ"], s["total_steps"], s["total_evals"], s["nmcmc"], e.live().tobytes()))
    assert out[0] == out[1] == out[2]


@pytest.mark.parametrize("dim,nlive,K,mix", [
    (2, 1000, 64, dict(mh=2, slice=1, hmc=1)),      # CPNest(nthreads=4, nslice=1, nhamiltonian=1)
    (10, 2000, 192, dict(mh=1, slice=1, hmc=1)),
    (10, 2000, 128, dict(mh=1, hmc=1)),
])
def test_mixed_population_logz_within_3_sigma(dim, nlive, K, mix):
    with make(dim=dim, nlive=nlive, mix=mix, maxmcmc=1000) as e:
        s, logz, ins = finish(e, K)
        analytic = -dim * math.log(20.0)
        sigma = math.sqrt(s["info"] / nlive)
        assert abs(logz - analytic) < 3.0 * sigma, (logz, analytic, sigma, s["nmcmc_kind"], s["rho_kind"])
        assert stats.kstest(ins, "uniform").pvalue > 1e-3
        # the replacements of every kind are distributed like the others': insertion ranks by range of the batch
        ins = np.asarray(ins)
        ins = ins[:len(ins) // K * K].reshape(-1, K)
        kb = np.concatenate(([0], np.cumsum(s["chains_kind"])))
        for k in range(3):
            if kb[k + 1] > kb[k]:
                assert stats.kstest(ins[:, kb[k]:kb[k + 1]].ravel(), "uniform").pvalue > 1e-4, ("kind", k)


def test_mixed_run_is_deterministic_and_resumes_bit_identically():
    mix = dict(mh=2, slice=1, hmc=1)
    K = 64

    def sig(e):
        s = e.state()
        return (s["logZ"], s["total_steps"], s["total_evals"], tuple(s["nmcmc_kind"]), s["slice_mu"], s["hmc_dt"],
                e.live().tobytes(), e.dead().tobytes())
    with make(dim=3, nlive=500, mix=mix, seed=11) as a, make(dim=3, nlive=500, mix=mix, seed=11) as b:
        for _ in range(12):
            a.ns_step(K)
            b.ns_step(K)
        assert sig(a) == sig(b)
        blob = a.save_checkpoint()
        for _ in range(8):
            a.ns_step(K)
        from cpnest_b200.engine import Engine
        c = Engine("gaussian_iid", [[-10, 10]] * 3, nlive=500, maxmcmc=400, seed=999, mix=mix)
        with c:
            c.load_checkpoint(blob)
            for _ in range(8):
                c.ns_step(K)
            assert sig(c) == sig(a)
        # a context with another population refuses the blob
        d = Engine("gaussian_iid", [[-10, 10]] * 3, nlive=500, maxmcmc=400, seed=11, mix=dict(mh=1, slice=1))
        with d:
            from cpnest_b200._lib import CpnError
            with pytest.raises(CpnError, match="different run"):
                d.load_checkpoint(blob)


def test_argument_errors():
    from cpnest_b200._lib import CpnError
    from cpnest_b200.engine import Engine
    with pytest.raises(ValueError, match="mix"):
        Engine("gaussian_iid", [[-1, 1]] * 2, nlive=100, mix=(0, 0, 0))
    with pytest.raises(ValueError, match="unknown sampler"):
        Engine("gaussian_iid", [[-1, 1]] * 2, nlive=100, mix=dict(nuts=1))
    with make(mix=dict(mh=1, slice=1, hmc=1), nlive=100) as e:
        with pytest.raises(CpnError, match="sampler kinds"):
            e.ns_step(2)


def test_host_protocol_serves_a_mixed_population():
    # cpn_evolve_host: requests [0, K) are served by the kinds' kernels in the same ranges
    K, D = 60, 3
    with make(dim=D, nlive=400, mix=dict(mh=1, slice=1, hmc=1)) as e:
        live = e.live()
        e.set_live(live)
        logLmin = live[K - 1, D]
        req = np.ascontiguousarray(live[K:2 * K])
        rep, steps, acc = e.evolve_host(req, logLmin, 40)
        assert e.state()["chains_kind"] == [20, 20, 20]
        assert np.all(rep[:, D] > logLmin)
        assert np.allclose(rep[:, D], -0.5 * (rep[:, :D] ** 2).sum(1) - 0.5 * D * math.log(2 * math.pi), rtol=1e-12)
        assert np.all(acc > 0)
        assert np.all(steps[:20] >= 40) and np.all(steps[20:40] > 40)      # MH: >= Nmcmc steps; slice: > Nmcmc slices
        assert not np.allclose(rep[:, :D], req[:, :D])
