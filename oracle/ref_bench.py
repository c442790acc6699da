"""CPU baseline runner: times the reference's own hot path on the host cores.

*** TEST / BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline and --impl reference). ***

kind == "reference": the UNMODIFIED reference installed in oracle/_ref by oracle/build_ref.sh
is imported and its `MetropolisHastingsSampler.yield_sample` (cpnest/sampler.py:322-358) is
driven with its own `DefaultProposalCycle` (cpnest/proposal.py:345-362), one sampler per host
core in forked processes, exactly the objects `CPNest(..., nthreads=cores)` would create
(cpnest/cpnest.py:197-217), minus the pipes: each process evolves its pool under a fixed
likelihood threshold for a bounded wall-clock budget and counts likelihood calls.
kind == "port": if oracle/_ref is absent, the pure-Python restatement in cpnest_oracle.py is
timed the same way.
"""
from __future__ import annotations

import math
import multiprocessing as mp
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def reference_available():
    return os.path.isdir(os.path.join(REF, "cpnest"))


def _worker_reference(args):
    dim, poolsize, nmcmc, maxmcmc, q, budget_s, seed = args
    sys.path.insert(0, REF)
    import cpnest.model
    import cpnest.proposal as rprop
    import cpnest.sampler as rsamp
    from cpnest.parameter import LivePoint
    from array import array

    class Counter:
        n = 0

    class GaussianModel(cpnest.model.Model):
        # the model of examples/test_50d.py:10-19, restated (the GPU box has no /root/reference)
        def __init__(self, dim):
            self.names = ['{0}'.format(i) for i in range(dim)]
            self.bounds = [[-10, 10] for _ in range(dim)]

        def log_likelihood(self, p):
            Counter.n += 1
            return np.sum([-0.5 * p[n] ** 2 - 0.5 * np.log(2.0 * np.pi) for n in p.names])

    class Val:
        def __init__(self, v):
            self.value = v

    class Manager:
        logLmin = Val(-np.inf)
        logLmax = Val(-np.inf)
        checkpoint_flag = Val(0)
        periodic_checkpoint_interval = np.inf
        nthreads = 1

        def connect_producer(self):
            return None, 0

    random.seed(seed)
    np.random.seed(seed)
    model = GaussianModel(dim)
    smp = rsamp.MetropolisHastingsSampler(model, maxmcmc, seed=seed, poolsize=poolsize,
                                          proposal=rprop.DefaultProposalCycle(), manager=Manager())
    smp.Nmcmc = nmcmc
    X = np.random.normal(size=(poolsize, dim))
    for row in X:
        p = LivePoint(list(model.names), d=array('d', row.tolist()))
        p.logP = model.log_prior(p)
        p.logL = model.log_likelihood(p)
        smp.evolution_points.append(p)
    logLmin = float(np.quantile([p.logL for p in smp.evolution_points], q))
    smp.proposal.set_ensemble(smp.evolution_points)
    gen = smp.yield_sample(logLmin)
    state = dict(chains=0)

    def run(budget):
        """evolve the pool for `budget` seconds; (likelihood calls, MCMC steps, accepted, chains, seconds) of this leg"""
        n0, s0, a0, c0 = Counter.n, smp.mcmc_counter, smp.mcmc_accepted, state["chains"]
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < budget:
            next(gen)
            state["chains"] += 1
            if state["chains"] % (poolsize // 4) == 0:
                smp.proposal.set_ensemble(smp.evolution_points)      # sampler.py:252-254
        dt = time.perf_counter() - t0
        return Counter.n - n0, smp.mcmc_counter - s0, smp.mcmc_accepted - a0, state["chains"] - c0, dt

    if budget_s is None:
        return run                      # persistent worker: the caller drives the legs
    return run(budget_s)


def _persistent_worker(conn, args):
    run = _worker_reference(args)
    conn.send("ready")
    while True:
        budget = conn.recv()
        if budget is None:
            break
        conn.send(run(budget))


class ReferencePool(object):
    """`cores` forked processes, each holding one reference MH sampler with its pool; `leg(budget_s)` lets all of them
    evolve for budget_s seconds and returns the summed counts.  Set-up (fork, pool of 1000 points) is paid once."""

    def __init__(self, dim=50, cores=None, poolsize=1000, nmcmc=100, maxmcmc=100, quantile=0.2, seed=1234):
        if not reference_available():
            raise RuntimeError("oracle/_ref is absent")
        self.cores = cores or os.cpu_count() or 1
        self.desc = (dim, poolsize, nmcmc, quantile)
        ctx = mp.get_context("fork")
        self.procs, self.conns = [], []
        for i in range(self.cores):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_persistent_worker, args=(b, (dim, poolsize, nmcmc, maxmcmc, quantile, None, seed + i)), daemon=True)
            pr.start()
            self.procs.append(pr)
            self.conns.append(a)
        for c in self.conns:
            assert c.recv() == "ready"

    def leg(self, budget_s):
        for c in self.conns:
            c.send(float(budget_s))
        res = [c.recv() for c in self.conns]
        tmax = max(r[4] for r in res)
        evals, steps, acc, chains = (sum(r[k] for r in res) for k in range(4))
        dim, poolsize, nmcmc, quantile = self.desc
        return dict(value=evals / tmax, unit="likelihood evals/s", cores=self.cores, kind="reference", steps_per_s=steps / tmax,
                    acceptance=acc / max(1, steps), chains=chains, wall_s=tmax,
                    sample=("unmodified reference (oracle/_ref) MH sampler (DefaultProposalCycle) on the %d-D Gaussian of "
                            "examples/test_50d.py: %d forked processes x %.1f s, poolsize %d, Nmcmc %d, threshold at the %.0f%% "
                            "pool quantile; %d chains, %d MCMC steps, %d likelihood calls" % (
                                dim, self.cores, budget_s, poolsize, nmcmc, 100 * quantile, chains, steps, evals)))

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for pr in self.procs:
            pr.join(timeout=5)


def _worker_port(args):
    dim, poolsize, nmcmc, maxmcmc, q, budget_s, seed = args
    from oracle import cpnest_oracle as O
    rng = random.Random(seed)
    rs = np.random.RandomState(seed)
    m = O.builtin_model("gaussian%dd" % dim)
    evals = [0]
    ll = m.log_likelihood

    def counted(x):
        evals[0] += 1
        return ll(x)
    m.log_likelihood = counted
    pool = [rs.normal(size=dim) for _ in range(poolsize)]
    pool_logL = [ll(x) for x in pool]
    pool_logP = [0.0] * poolsize
    logLmin = float(np.quantile(pool_logL, q))
    cycle = O.make_cycle(rs)
    _, _, w, v = O.ensemble_covariance(np.array(pool))
    idx = 0
    steps = acc = chains = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        tape = []
        ci = idx
        n = poolsize - 1
        for _ in range(maxmcmc):
            ci = (ci + 1) % len(cycle)
            k = cycle[ci]
            if k == O.WALK:
                tape += [["sample", rng.sample(range(n), 3)]] + [["gauss", rng.gauss(0, 1)] for _ in range(3)]
            elif k == O.STRETCH:
                tape += [["choice", rng.randrange(n)], ["uniform", rng.uniform(-1, 1)]]
            elif k == O.DE:
                tape += [["sample", rng.sample(range(n), 2)], ["gauss", rng.gauss(1.0, 1e-4)]]
            else:
                tape += [["randrange", rng.randrange(dim)], ["gauss", rng.gauss(0, 1)]]
            tape.append(["random", rng.random()])
        r = O.mh_chain(m, pool, pool_logL, pool_logP, cycle, idx, w, v, tape, nmcmc, maxmcmc, logLmin)
        idx = r["cycle_idx"]
        steps += r["steps"]
        acc += r["accepted"]
        chains += 1
    dt = time.perf_counter() - t0
    return evals[0], steps, acc, chains, dt


def time_reference(dim=50, cores=None, budget_s=15.0, poolsize=1000, nmcmc=100, maxmcmc=100, quantile=0.2,
                   seed=1234, force_port=False):
    """Likelihood evaluations per second of the reference MH sampler over `cores` processes.
    Returns a dict ready for bench.py's cpu_baseline."""
    cores = cores or os.cpu_count() or 1
    use_ref = reference_available() and not force_port
    worker = _worker_reference if use_ref else _worker_port
    args = [(dim, poolsize, nmcmc, maxmcmc, quantile, budget_s, seed + i) for i in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(worker, args)
    wall = time.perf_counter() - t0
    evals = sum(r[0] for r in res)
    steps = sum(r[1] for r in res)
    acc = sum(r[2] for r in res)
    chains = sum(r[3] for r in res)
    tmax = max(r[4] for r in res)
    return dict(
        value=evals / tmax, unit="likelihood evals/s", cores=cores, kind="reference" if use_ref else "port",
        steps_per_s=steps / tmax, acceptance=acc / max(1, steps), chains=chains, wall_s=wall,
        sample=("%s MH sampler (DefaultProposalCycle) on the %d-D Gaussian of examples/test_50d.py: %d forked "
                "processes x %.0f s, poolsize %d, Nmcmc %d, threshold at the %.0f%% pool quantile; %d chains, "
                "%d MCMC steps, %d likelihood calls" % (
                    "unmodified reference (oracle/_ref)" if use_ref else "oracle port (cpnest_oracle.mh_chain)",
                    dim, cores, budget_s, poolsize, nmcmc, 100 * quantile, chains, steps, evals)))


if __name__ == "__main__":
    import json
    sys.path.insert(0, os.path.dirname(HERE))
    print(json.dumps(time_reference(budget_s=float(sys.argv[1]) if len(sys.argv) > 1 else 5.0,
                                    force_port="--port" in sys.argv), indent=1))


# ----------------------------------------------------------------------------------------------------------------
# The stock path: the UNMODIFIED reference run through its own public API, `cpnest.CPNest(model, nlive, nthreads=cores,
# ...).run()` (cpnest/cpnest.py:103-316), in a child process group, observed from outside for a bounded time
# (SURVEY.md 8d "CPU reference timing": wall clock around run(), likelihood evaluations counted by a wrapper model that
# bumps a multiprocessing.Value every 1000 calls).  MCMC steps are counted the same way through log_prior, which the MH
# sampler calls once per step (cpnest/sampler.py:335).
# ----------------------------------------------------------------------------------------------------------------
def _stock_child(dim, nlive, nthreads, maxmcmc, poolsize, seed, n_evals, n_steps, outdir):
    os.setsid()                                     # own process group: the parent stops exactly this group
    devnull = open(os.devnull, "w")
    os.dup2(devnull.fileno(), 1)
    os.dup2(devnull.fileno(), 2)
    sys.path.insert(0, REF)
    import cpnest
    import cpnest.model

    class GaussianModel(cpnest.model.Model):
        # the model of examples/test_50d.py:5-28, restated (the GPU box has no /root/reference)
        _e = 0
        _s = 0

        def __init__(self, dim):
            self.names = ['{0}'.format(i) for i in range(dim)]
            self.bounds = [[-10, 10] for _ in range(dim)]

        def log_likelihood(self, p):
            GaussianModel._e += 1
            if GaussianModel._e == 1000:
                GaussianModel._e = 0
                with n_evals.get_lock():
                    n_evals.value += 1000
            return np.sum([-0.5 * p[n] ** 2 - 0.5 * np.log(2.0 * np.pi) for n in p.names])

        def log_prior(self, p):
            GaussianModel._s += 1
            if GaussianModel._s == 1000:
                GaussianModel._s = 0
                with n_steps.get_lock():
                    n_steps.value += 1000
            return super(GaussianModel, self).log_prior(p)

    work = cpnest.CPNest(GaussianModel(dim), verbose=2, nthreads=nthreads, nlive=nlive, maxmcmc=maxmcmc, poolsize=poolsize,
                         seed=seed, output=outdir)
    work.run()


class StockReferenceRun(object):
    """cpnest.CPNest(...).run() of the unmodified reference in a child process group; progress() reads the counters and the
    last iteration line of its cpnest.log (NestedSampling.py:360)."""

    def __init__(self, dim=50, nlive=10000, cores=None, maxmcmc=5000, poolsize=1000, seed=1234):
        import tempfile
        if not reference_available():
            raise RuntimeError("oracle/_ref is absent")
        self.cores = cores or os.cpu_count() or 1
        self.args = dict(dim=dim, nlive=nlive, nthreads=self.cores, maxmcmc=maxmcmc, poolsize=poolsize, seed=seed)
        ctx = mp.get_context("fork")
        self.n_evals = ctx.Value("q", 0)
        self.n_steps = ctx.Value("q", 0)
        self.outdir = tempfile.mkdtemp(prefix="cpnest_ref_")
        self.proc = ctx.Process(target=_stock_child, args=(dim, nlive, self.cores, maxmcmc, poolsize, seed, self.n_evals,
                                                           self.n_steps, self.outdir))
        self.t0 = time.perf_counter()
        self.proc.start()

    def counts(self):
        return self.n_evals.value, self.n_steps.value, time.perf_counter()

    def progress(self):
        """(NS iteration, Nmcmc, logZ) from the last progress line of cpnest.log, or None before the first one."""
        import re
        try:
            with open(os.path.join(self.outdir, "cpnest.log"), "rb") as f:
                f.seek(0, 2)
                f.seek(max(0, f.tell() - 4096))
                tail = f.read().decode(errors="replace")
        except OSError:
            return None
        m = re.findall(r"(\d+): n:\s*(\d+) NS_acc:([\d.]+) .*? logZ: (\S+) logLmax", tail)
        if not m:
            return None
        it, n, acc, z = m[-1]
        return dict(iteration=int(it), nmcmc=int(n), ns_acceptance=float(acc), logZ=float(z))

    def alive(self):
        return self.proc.is_alive()

    def wait_ready(self, timeout_s=240.0):
        """until the nested-sampling loop is running (first progress line); False if the run ended or timed out first"""
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < timeout_s:
            if self.progress() is not None:
                return True
            if not self.alive():
                return False
            time.sleep(0.5)
        return False

    def stop(self):
        import shutil
        import signal
        if self.proc.pid is not None and self.proc.is_alive():
            try:
                os.killpg(self.proc.pid, signal.SIGKILL)     # the group created by os.setsid() in the child
            except (ProcessLookupError, PermissionError):
                pass
        self.proc.join(timeout=10)
        shutil.rmtree(self.outdir, ignore_errors=True)
