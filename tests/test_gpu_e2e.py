"""End-to-end nested-sampling runs on the device (engine level): log-evidence within 3 sigma of
the analytic value, sigma = sqrt(H/nlive) (BASELINE.json north_star; the reference's own tests use
2 sigma, tests/test_2d.py:37-38), insertion indices uniform, and size-independent invariants."""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import cpnest_oracle as O

pytestmark = pytest.mark.gpu


def run_ns(functor, bounds, nlive, K, maxmcmc=200, seed=1234, params=None, tolerance=0.1, **kw):
    from cpnest_b200.engine import Engine
    e = Engine(functor, bounds, nlive=nlive, maxmcmc=maxmcmc, seed=seed, params=params, **kw)
    e.set_cycle(O.make_cycle(np.random.RandomState(seed)))
    e.init_prior()
    nb = e.run(K, tolerance=tolerance)
    ins = e.insertion_indices()
    rect, trap = e.finalise()
    s = e.state()
    return e, s, trap, ins, nb


def check_logz(s, logz, analytic, nlive, nsigma=3.0):
    sigma = math.sqrt(s["info"] / nlive)
    assert abs(logz - analytic) < nsigma * sigma, "logZ %.4f vs analytic %.4f (sigma %.4f, H %.2f)" % (
        logz, analytic, sigma, s["info"])


def test_gaussian2d_logz_and_invariants():
    # config 1: tests/test_2d.py model, nlive=1000
    e, s, logz, ins, nb = run_ns("gaussian_iid", [[-10, 10]] * 2, nlive=1000, K=64)
    with e:
        check_logz(s, logz, -2 * math.log(20.0), 1000)
        assert 2.5 < s["info"] < 4.2                     # docs/index.rst:92-93 quotes H = 3.37
        assert stats.kstest(ins, "uniform").pvalue > 1e-3
        dead = e.dead()
        assert len(dead) == s["iteration"] == nb * 64 + 1000
        assert np.all(np.diff(dead[:, 2]) >= 0)          # nested samples come out ordered by logL
        assert np.all(np.abs(dead[:, :2]) < 10)
        assert np.allclose(dead[:, 2], -0.5 * (dead[:, :2] ** 2).sum(1) - math.log(2 * math.pi), rtol=1e-12)
        # posterior moments from the weighted dead points
        _, lw = O.compute_weights(dead[:, 2], 1000)
        w = np.exp(lw - lw.max())
        w /= w.sum()
        mean = (w[:, None] * dead[:, :2]).sum(0)
        var = (w[:, None] * (dead[:, :2] - mean) ** 2).sum(0)
        assert np.all(np.abs(mean) < 0.15) and np.all(np.abs(var - 1.0) < 0.2)


def test_gaussian2d_is_deterministic_per_seed():
    a = run_ns("gaussian_iid", [[-10, 10]] * 2, nlive=500, K=32, seed=7)
    b = run_ns("gaussian_iid", [[-10, 10]] * 2, nlive=500, K=32, seed=7)
    c = run_ns("gaussian_iid", [[-10, 10]] * 2, nlive=500, K=32, seed=8)
    assert a[2] == b[2] and a[1]["iteration"] == b[1]["iteration"]
    assert np.array_equal(a[0].dead(), b[0].dead())
    assert a[2] != c[2]
    for r in (a, b, c):
        r[0].close()


def test_ring_logz():
    # tests/test_ring.py: analytic -2.31836
    e, s, logz, ins, nb = run_ns("ring", [[-2, 2]] * 2, nlive=1000, K=64, params=dict(radius=1.0, thickness=0.1))
    with e:
        check_logz(s, logz, -2.318358, 1000)
        assert stats.kstest(ins, "uniform").pvalue > 1e-3


def test_half_gaussian_and_1d_logz():
    e, s, logz, ins, nb = run_ns("gaussian_iid", [[-10, 10]], nlive=1000, K=64)
    with e:
        check_logz(s, logz, -math.log(20.0), 1000)
    e, s, logz, ins, nb = run_ns("gaussian_iid", [[0, 20]], nlive=500, K=32)
    with e:
        check_logz(s, logz, math.log(stats.norm.cdf(20) - 0.5) - math.log(20.0), 500)


def test_gaussian_mean_sigma_nonflat_prior():
    # tests/test_gaussian.py: logZ = log int L(mean,sigma) p(mean,sigma); evaluate by quadrature
    from scipy import integrate
    # the sampler only sees prior ratios, so Z is taken w.r.t. the NORMALISED 1/sigma prior on the box
    # (the reference's log_prior, -log(sigma) - log(10) - log(0.95), integrates to 10 ln(20) / 9.5, not 1)
    f = lambda m, s: math.exp(-0.5 * m * m / (s * s)) / (s * math.sqrt(2 * math.pi)) / s
    Z, _ = integrate.dblquad(f, 0.05, 1.0, lambda s: -5, lambda s: 5, epsabs=1e-10)
    Z /= 10.0 * math.log(1.0 / 0.05)
    e, s, logz, ins, nb = run_ns("gaussian_mean_sigma", [[-5, 5], [0.05, 1]], nlive=1000, K=64)
    with e:
        check_logz(s, logz, math.log(Z), 1000)


def test_eggbox_logz():
    # config 2: anchor 235.85594 by quadrature; 18 equal modes
    e, s, logz, ins, nb = run_ns("eggbox", [[0, 10 * math.pi]] * 2, nlive=2000, K=128)
    with e:
        check_logz(s, logz, 235.85594, 2000)


def test_rosenbrock2d_logz():
    e, s, logz, ins, nb = run_ns("rosenbrock", [[-5, 5]] * 2, nlive=2000, K=128, maxmcmc=400)
    with e:
        check_logz(s, logz, -5.804132, 2000)


def test_gaussian50d_logz():
    # config 4 model at a test-sized live set: analytic -149.786614 (examples/test_50d.py:16)
    e, s, logz, ins, nb = run_ns("gaussian_iid", [[-10, 10]] * 50, nlive=2000, K=256, maxmcmc=5000)
    with e:
        check_logz(s, logz, -50 * math.log(20.0), 2000)
        assert stats.kstest(ins[-20000:], "uniform").pvalue > 1e-3
        assert 60 < s["info"] < 100


def test_full_size_invariants_50d_nlive_1e4():
    """BASELINE size (50-D Gaussian, nlive=10^4): size-independent properties of the batch step."""
    from cpnest_b200.engine import Engine
    N, K = 10000, 2048
    with Engine("gaussian_iid", [[-10, 10]] * 50, nlive=N, maxmcmc=100, seed=3) as e:
        e.set_cycle(O.make_cycle(np.random.RandomState(3)))
        e.init_prior()
        live0 = e.live()
        assert np.all(np.diff(live0[:, 50]) >= 0)
        for b in range(6):
            e.ns_step(K)
        s = e.state()
        live = e.live()
        dead = e.dead()
        assert s["iteration"] == 6 * K == len(dead)
        assert np.all(np.diff(live[:, 50]) >= 0) and np.all(np.diff(dead[:, 50]) >= 0)
        assert live[0, 50] > s["logLmin"] == dead[-1, 50]
        assert np.all(np.abs(live[:, :50]) < 10)
        logL, logP = e.eval_model(live[:, :50])
        assert np.allclose(logL, live[:, 50], rtol=1e-13, atol=0) and np.all(logP == 0)
        ins = e.insertion_indices()
        assert len(ins) == 6 * K and ins.min() >= 0 and ins.max() <= 1
        assert s["total_evals"] <= s["total_steps"] and s["total_accepted"] <= s["total_evals"]
        # the union of live and dead points holds every point exactly once (no slot lost in the merge)
        allp = np.concatenate((live, dead))
        assert s["failed_chains"] == 0
        assert len(np.unique(allp[:, :3], axis=0)) == len(allp)


def test_chain_that_cannot_move_is_sent_again_instead_of_returning_a_survivor():
    """`consume_sample` asks again until the reply beats the threshold (cpnest/NestedSampling.py:354-357).  With maxmcmc = 3
    a good part of the chains makes no accepted step in one attempt; each of them starts again from another survivor, so no
    reply is the copy of a live point and no chain is reported as failed."""
    from cpnest_b200.engine import Engine
    N, K = 2000, 512
    with Engine("gaussian_iid", [[-10, 10]] * 10, nlive=N, maxmcmc=3, nmcmc_init=3, adapt_nmcmc=False, seed=11) as e:
        e.set_cycle(O.make_cycle(np.random.RandomState(5)))
        e.init_prior()
        longest = 0
        for b in range(12):
            e.ns_step(K)
            rows, steps, acc = e.last_batch(K)
            assert np.all(acc > 0)
            assert np.all(rows[:, 10] > e.state()["logLmin"])
            longest = max(longest, int(steps.max()))
        s = e.state()
        assert longest > 3, "no chain needed a second attempt: the test does not exercise the rule"
        assert s["failed_chains"] == 0
        allp = np.concatenate((e.live(), e.dead()))
        assert len(np.unique(allp[:, :3], axis=0)) == len(allp)
        # without the rule about a fifth of the chains of this set-up would have handed back a copy of their start point
        assert s["total_steps"] > 12 * K * 3


def test_checkpoint_resume_continues_bit_identically():
    """NestedSampler.checkpoint / resume (cpnest/NestedSampling.py:495-537): a run restored from its state blob in a NEW
    context continues exactly like the uninterrupted one (also across a pending ensemble-statistics refresh)."""
    import hashlib
    from cpnest_b200.engine import Engine

    def make(sampler):
        e = Engine("gaussian_iid", [[-10, 10]] * 6, nlive=1000, maxmcmc=100, seed=11, sampler=sampler)
        if sampler == "mh":
            e.set_cycle(O.make_cycle(np.random.RandomState(11)))
        return e

    def digest(e):
        s = e.state()
        return (hashlib.sha1(e.dead().tobytes() + e.live().tobytes() + e.insertion_indices().tobytes()).hexdigest(),
                s["logZ"], s["info"], s["nmcmc"], s["total_evals"], s["slice_mu"], s["hmc_dt"])

    for sampler in ("mh", "slice", "hmc"):
        for cut in (7, 8, 9, 10):                       # the refresh cadence makes some of these cuts fall mid-refresh
            a = make(sampler)
            a.init_prior()
            for _ in range(cut):
                a.ns_step(250)
            blob = a.save_checkpoint()
            for _ in range(9):
                a.ns_step(250)
            ref = digest(a)
            a.close()
            b = make(sampler)
            b.load_checkpoint(blob)
            for _ in range(9):
                b.ns_step(250)
            assert digest(b) == ref, (sampler, cut)
            b.close()
    with pytest.raises(Exception, match="different run"):
        c = Engine("gaussian_iid", [[-10, 10]] * 6, nlive=500, maxmcmc=100, seed=11)
        c.load_checkpoint(blob)


@pytest.mark.parametrize("dim,lanes", [(50, 0), (50, 32), (10, 0), (10, 8), (6, 0)])
def test_zero_padded_rows_give_the_same_run_bit_for_bit(dim, lanes, monkeypatch):
    """For the iid Gaussian centred at 0 the live rows are zero-padded to G*E doubles and the MH kernel runs its loop without
    the masks of the slots beyond D (mh_kernel.cuh, Model::pad_ok).  CPN_NO_ROW_PAD=1 keeps rows of roundup(D, 4) doubles and
    the masked loop: both must produce the same chains, hence the same run, bit for bit."""
    from cpnest_b200.engine import Engine

    def run():
        e = Engine("gaussian_iid", [[-10, 10]] * dim, nlive=1500, maxmcmc=300, seed=42, lanes=lanes)
        e.set_cycle(O.make_cycle(np.random.RandomState(42)))
        e.init_prior()
        for _ in range(12):
            e.ns_step(300)
        s = e.state()
        out = (s["logZ"], s["total_steps"], s["total_accepted"], s["total_evals"], s["nmcmc"], e.live().tobytes(), e.dead().tobytes())
        e.close()
        return out
    padded = run()
    monkeypatch.setenv("CPN_NO_ROW_PAD", "1")
    plain = run()
    assert padded == plain


def test_correlated_gaussian_50d_logz():
    """BASELINE config 4, correlated flavour: N(0, Sigma), Sigma = A A^T with condition number 1e2 well inside the
    box, so logZ stays -50 log 20 = -149.786614.  The one analytic config whose likelihood (D^2 + 3D = 2650 flops) can be
    FP64 bound; the D x D group mat-vec of the functor is checked against numpy on the way."""
    m = O.make_correlated_gaussian(50)
    e, s, logz, ins, nb = run_ns(m.functor, m.bounds, nlive=2000, K=500, maxmcmc=5000, params=m.params, seed=21)
    with e:
        check_logz(s, logz, -50 * math.log(20.0), 2000)
        assert 110 < s["info"] < 165, s["info"]        # H = 79 + sum_i -log(sigma_i) = 79 + 57.6 nats
        assert stats.kstest(ins[-8000:], "uniform").pvalue > 1e-3
        dead = e.dead()
        ref = np.array([m.log_likelihood(x) for x in dead[-200:, :50]])
        assert np.allclose(dead[-200:, 50], ref, rtol=1e-11, atol=0)
        assert s["failed_chains"] == 0


def test_two_lag_control_on_a_bimodal_live_set():
    """20-D mixture of two separated Gaussians (BASELINE config 5's stand-in): chains stay in their mode, so the start/end
    correlation has a part that no chain length removes.  The single-lag control reads it as "not decorrelated" and ends with
    the chain length pinned at maxmcmc; the two-lag control (mid-chain snapshot: the two half-chain increments do not contain
    the mode centre, q = 1 + 2 corr(d1, d2)) applies the target to the part that decays and needs fewer steps for the same
    evidence (measured: 2.8e8 -> 2.2e8 steps, final chain length 2000 -> ~900)."""
    mix = dict(w=0.3, m1=-2.5, s1=1.0, m2=2.5, s2=1.0)
    res = {}
    for two in (False, True):
        e, s, logz, ins, nb = run_ns("gaussian_mixture_nd", [[-10, 10]] * 20, nlive=4000, K=1000, maxmcmc=2000, params=mix, seed=31,
                                     two_lag=two)
        with e:
            check_logz(s, logz, -20 * math.log(20.0), 4000)
            res[two] = (s["total_steps"], s["nmcmc"], s["rho_max"], s["two_lag_kind"][0], s["rho_decay_kind"][0])
    assert res[False][1] > 1800 and res[False][3] == 0, res      # pinned at maxmcmc, two-lag never on
    assert res[True][3] > 0 and 0 <= res[True][4] < 0.2, res       # switched on; the decaying part is small at the end
    assert res[True][2] > 0.25, res                                # while the raw start/end correlation stays large
    assert res[True][1] < 0.75 * res[False][1], res
    assert res[True][0] < 0.9 * res[False][0], res
