"""Edge cases of the batch step through the C-ABI: smallest and ragged sizes, one dimension, more dimensions than a
warp has lanes, a single chain per batch (the reference's nthreads=1), chains that cannot move, bad arguments."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(functor, D, nlive, K, sampler, maxmcmc=50, nbatches=6, bound=10.0, **kw):
    from cpnest_b200.engine import Engine
    e = Engine(functor, [[-bound, bound]] * D, nlive=nlive, maxmcmc=maxmcmc, seed=5, sampler=sampler, **kw)
    e.init_prior()
    for _ in range(nbatches):
        e.ns_step(K)
    s = e.state()
    live, dead = e.live(), e.dead()
    assert s["iteration"] == nbatches * K == len(dead)
    assert np.all(np.diff(live[:, D]) >= 0) and np.all(np.diff(dead[:, D]) >= 0)
    assert live[0, D] > dead[-1, D] or s["failed_chains"] > 0
    assert np.all(np.abs(live[:, :D]) < bound) and np.all(np.isfinite(live))
    logL, _ = e.eval_model(live[:, :D])
    assert np.allclose(logL, live[:, D], rtol=1e-12, atol=1e-12)
    e.close()
    return s


@pytest.mark.parametrize("sampler", ["mh", "slice", "hmc"])
def test_sizes_and_dimensions(sampler):
    _run("gaussian_iid", 1, 8, 1, sampler)                 # one dimension, one chain per batch, tiny live set
    _run("gaussian_iid", 3, 37, 5, sampler)                # ragged: nothing divides anything
    _run("gaussian_iid", 2, 64, 61, sampler)               # K = nlive - 3, the largest batch allowed
    _run("gaussian_iid", 33, 200, 50, sampler)             # one more dimension than a warp has lanes
    _run("gaussian_iid", 100, 400, 100, sampler, nbatches=3)
    _run("rosenbrock", 7, 100, 25, sampler, bound=5.0)     # neighbour shuffles across an odd dimension


def test_chains_that_cannot_move_are_sent_again():
    """maxmcmc = 1 with a sharp constraint: most chains make no move in one attempt.  The reference resends such points
    (NestedSampling.py:354-357); the device chain starts again from another survivor, up to 15 times, so a batch of 50
    costs many more than 50 steps; the few chains that fail every attempt hand back a copy of a survivor and are counted."""
    s = _run("gaussian_iid", 20, 200, 50, "mh", maxmcmc=1, nbatches=10)
    assert 10 * 50 < s["total_steps"] <= 16 * 10 * 50
    assert s["failed_chains"] < 0.2 * 10 * 50


def test_bad_arguments_fail_loudly():
    from cpnest_b200.engine import Engine
    from cpnest_b200._lib import CpnError
    with pytest.raises(CpnError, match="nlive must be >= 4"):
        Engine("gaussian_iid", [[-1, 1]], nlive=3)
    with pytest.raises(CpnError, match="is empty"):
        Engine("gaussian_iid", [[1, 1]], nlive=16)
    with pytest.raises(CpnError, match="ring needs dim=2"):
        Engine("ring", [[-2, 2]] * 3, nlive=16, params=dict(radius=1.0, thickness=0.1))
    with Engine("gaussian_iid", [[-1, 1]] * 2, nlive=16) as e:
        with pytest.raises(CpnError, match="live set not initialised"):
            e.ns_step(4)
        e.init_prior()
        with pytest.raises(CpnError, match=r"must be in \[1, nlive-3\]"):
            e.ns_step(14)
        with pytest.raises(CpnError, match="must be in"):
            e.set_nmcmc(10 ** 6)


def test_large_batches_are_ranked_in_segments_with_ties():
    """K > 8192 replacements are ranked by segment-wise counting + merge (csrc/ns_kernels.cuh k_rank_*_seg).  The replies
    are scripted (cpn_stage_replacements) as copies of survivors, each used several times, so equal keys occur inside and
    across segments: the merged order must still be a sorted permutation of the live set, and no point may be lost."""
    from cpnest_b200.engine import Engine
    N, K, D = 20000, 10000, 3
    rs = np.random.RandomState(4)
    with Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=1, seed=2) as e:
        e.init_prior()
        for b in range(4):
            live = e.live()
            surv = live[K:]
            reps = surv[rs.randint(0, 300, size=K) * 7 % len(surv)]          # 300 distinct rows for 10000 replies
            e.select_and_accumulate(K)
            e.stage_replacements(reps)
            e.insert_batch(K)
            new_live, dead = e.live(), e.dead()
            assert np.all(np.diff(new_live[:, D]) >= 0) and np.all(np.diff(dead[:, D]) >= 0)
            assert len(dead) == (b + 1) * K and dead[-1, D] <= new_live[0, D]
            # the new live set is exactly survivors + replies, as a multiset
            want = np.sort(np.concatenate((surv[:, D], reps[:, D])))
            assert np.array_equal(new_live[:, D], want)
            assert np.array_equal(dead[b * K:], live[:K])
            logL, _ = e.eval_model(new_live[:, :D])
            assert np.allclose(logL, new_live[:, D], rtol=1e-12, atol=1e-12)
        ins = e.insertion_indices()
        assert len(ins) == 4 * K and ins.min() >= 0 and ins.max() <= 1


def test_host_table_protocol_equals_the_buffer_protocol():
    """cpn_evolve_host_table (the device reads the requests from / writes the replies into the caller's pinned host table)
    against cpn_evolve_host (request / reply buffers moved by the caller): same chains, same replies, same table."""
    import torch
    from cpnest_b200._lib import CpnError
    from cpnest_b200.engine import Engine
    D, N, K = 7, 600, 200
    rs = np.random.RandomState(0)

    def make():
        e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=200, seed=21)
        e.init_prior()
        live = e.live()
        e.set_live(live)
        return e, live
    a, live_a = make()
    b, live_b = make()
    assert np.array_equal(live_a, live_b)
    table_t = torch.from_numpy(live_b.copy()).pin_memory()
    table = table_t.numpy()
    with a, b:
        for it in range(4):
            logL = live_a[:, D]
            part = np.argpartition(logL, K - 1)
            worst, surv = part[:K], part[K:]
            logLmin = logL[worst].max()
            starts = surv[rs.randint(0, N - K, size=K)]
            req = np.ascontiguousarray(live_a[starts])
            rep, st_a, ac_a = a.evolve_host(req, logLmin, 30, slots=worst.astype(np.int32))
            live_a[worst] = rep
            st_b, ac_b = b.evolve_host_table(table, starts, worst, logLmin, 30, rows_async=(it % 2 == 1))
            # CPN_TABLE_ROWS_ASYNC: keys and counters are final at return, the coordinates after a synchronize
            assert np.array_equal(table[:, D:], live_a[:, D:]), it
            b.synchronize()
            assert np.array_equal(table, live_a), it
            assert np.array_equal(st_a, st_b) and np.array_equal(ac_a, ac_b)
            assert np.all(table[worst, D] > logLmin)
        assert a.state()["total_evals"] == b.state()["total_evals"]
        with pytest.raises(CpnError, match="pinned"):
            b.evolve_host_table(live_a, starts, worst, logLmin, 30)          # pageable numpy memory
        bad = starts.copy()
        bad[5] = N
        with pytest.raises(CpnError, match="out of range"):
            b.evolve_host_table(table, bad, worst, logLmin, 30)


def test_team_mode_chains_agree_with_the_plain_functor():
    """The mixture functor (1e4 data terms per call) runs MH chains in team mode: the four warps of a block evolve one chain
    in lock step and split the data sum.  The logL every chain hands back must be the functor's value at the point it hands
    back (evaluated here by the eval kernel, which has no teams), the run must be reproducible, and chains must move."""
    from cpnest_b200.engine import Engine
    rs = np.random.RandomState(1234)
    n = 2000
    data = np.concatenate((rs.normal(0.5, 0.5, size=n // 10), rs.normal(-1.5, 0.03, size=n - n // 10)))
    b = [[-3, 3], [0.01, 1], [-3, 3], [0.01, 1], [0.0, 1.0]]

    def run(lanes):
        e = Engine("gaussian_mixture", b, nlive=300, maxmcmc=60, seed=9, params=dict(data=data), lanes=lanes)
        e.init_prior()
        for _ in range(6):
            e.ns_step(37)                      # ragged: not a multiple of anything
        s = e.state()
        live = e.live()
        rows, steps, acc = e.last_batch(37)
        logL, logP = e.eval_model(live[:, :5])
        e.close()
        return s, live, rows, steps, acc, logL, logP
    s, live, rows, steps, acc, logL, logP = run(0)          # automatic: one warp per chain -> team mode
    assert np.allclose(live[:, 5], logL, rtol=1e-12, atol=0)
    assert np.allclose(live[:, 6], logP, rtol=1e-12, atol=0)
    assert np.all(rows[:, 5] > s["logLmin"]) and np.all(acc > 0) and np.all(steps >= 1)
    assert s["iteration"] == 6 * 37
    s2, live2 = run(0)[:2]
    assert s2["logZ"] == s["logZ"] and np.array_equal(live, live2)
    # 8 lanes per chain: no teams; a different but equally valid run
    s3, live3, rows3, steps3, acc3, logL3, _ = run(8)
    assert np.allclose(live3[:, 5], logL3, rtol=1e-12, atol=0) and np.all(acc3 > 0)


@pytest.mark.parametrize("D,P", [(1, 5), (2, 4), (7, 33), (8, 64), (9, 1000), (17, 129), (50, 10000), (63, 257), (64, 4097), (100, 300)])
def test_tensor_core_covariance_for_ragged_shapes(D, P):
    """np.cov of the ensemble (EnsembleEigenVector.update_eigenvectors, cpnest/proposal.py:311-323) is accumulated with fp64
    tensor-core MMAs in 8x8 tiles over groups of four points: dimensions that are not multiples of 8, ensembles that are
    not multiples of 4 or of the 32 slices, and the D == 1 variance branch (np.var, ddof 0, proposal.py:314-318)."""
    from cpnest_b200.engine import Engine
    rs = np.random.RandomState(D * 1000 + P)
    X = rs.normal(size=(P, D)) @ rs.normal(size=(D, D)) + 3.0 * rs.normal(size=(1, D))
    with Engine("gaussian_iid", [[-1e3, 1e3]] * D, nlive=P) as e:
        e.set_live(np.concatenate((X, np.arange(float(P))[:, None], np.zeros((P, 1))), axis=1))
        mean, cov, w, v = e.ensemble_stats()
    ref = np.atleast_2d(np.var(X[:, 0])) if D == 1 else np.cov(X.T)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert np.max(np.abs(mean - X.mean(0))) < 1e-13 * (1 + np.max(np.abs(X)))
    assert np.max(np.abs(cov - ref) / scale) < 1e-12
    assert np.array_equal(cov, cov.T)
    assert np.max(np.abs(ref @ v - v * w)) < 1e-10 * max(1.0, np.max(np.abs(w)))
