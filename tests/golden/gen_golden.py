#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

The reference (johnveitch/cpnest) has no golden vectors of its own (SURVEY.md section 8c:
"parity unpinned" at the step level), so the oracle in oracle/cpnest_oracle.py is pinned
against outputs of the reference itself, produced here.  The reference is imported from
oracle/_ref (built by oracle/build_ref.sh from /root/reference); it cannot be assumed to
exist on the GPU box, therefore the vectors are committed as JSON and this script is the
committed recipe that made them.

Every random number the reference consumes on the hot path is *recorded* by patching the
names the reference imported (`cpnest/proposal.py:6-7`, `cpnest/sampler.py:9`), so a
fixture is (inputs, random tape, outputs): the oracle and the CUDA replay kernels consume
the same tape and must reproduce the outputs to 1e-12.

Run:  bash oracle/build_ref.sh && python tests/golden/gen_golden.py
"""
import json
import math
import os
import random as pyrandom
import sys
from array import array
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import cpnest  # noqa: E402  (the reference)
import cpnest.model  # noqa: E402
import cpnest.nest2pos as n2p  # noqa: E402
import cpnest.proposal as rprop  # noqa: E402
import cpnest.sampler as rsamp  # noqa: E402
from cpnest.NestedSampling import NestedSampler, _NSintegralState  # noqa: E402
from cpnest.parameter import LivePoint  # noqa: E402


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


def fl(x):
    return float(x)


# --------------------------------------------------------------------------------------
# A. evidence integrator  (cpnest/NestedSampling.py:88-144)
# --------------------------------------------------------------------------------------
def gen_integrator():
    cases = []

    def run(nlive, steps, finalise=True, tag=""):
        st = _NSintegralState(nlive)
        trace = []
        for (logL, nl, nrep) in steps:
            st.increment(logL, nlive=nl, nreplace=nrep)
            trace.append([fl(st.logZ), fl(st.info), fl(st.logw)])
        out = dict(tag=tag, nlive=nlive, steps=[[fl(a), b, c] for a, b, c in steps], trace=trace)
        if finalise:
            out["finalised_logZ"] = fl(st.finalise())
        cases.append(out)

    # GV1 / GV2 of SURVEY.md section 8c
    run(10, [(-3.0, None, 1), (-2.0, None, 1), (-1.0, None, 1), (-0.5, None, 1)], finalise=False, tag="GV1")
    run(10, [(-3.0, None, 4), (-2.0, None, 4), (-1.0, None, 4), (-0.5, None, 4)], tag="GV2")
    rng = np.random.RandomState(11)
    # a full synthetic run: main loop with batch nreplace, then the final sweep nlive=N-i
    for nlive, nrep, niter in [(50, 1, 400), (64, 8, 60), (1000, 16, 300)]:
        logLs = np.sort(rng.normal(-30.0, 8.0, size=niter + nlive))
        steps = [(float(l), None, nrep) for l in logLs[:niter]]
        steps += [(float(l), nlive - i, 1) for i, l in enumerate(logLs[niter:])]
        run(nlive, steps, tag="synthetic_n%d_k%d" % (nlive, nrep))
    # very negative logL (50-D like) so that logZ underflows in linear space
    logLs = np.sort(-0.5 * rng.chisquare(50, size=700) - 46.0 - 300.0)
    steps = [(float(l), None, 2) for l in logLs[:500]] + [(float(l), 200 - i, 1) for i, l in enumerate(logLs[500:])]
    run(200, steps, tag="deep")
    dump("integrator.json", cases)


# --------------------------------------------------------------------------------------
# B. nest2pos helpers (cpnest/nest2pos.py:17-71, 175-199)
# --------------------------------------------------------------------------------------
def gen_nest2pos():
    out = {}
    rng = np.random.RandomState(5)
    x = rng.normal(0, 3, 20)
    y = x - np.abs(rng.normal(0, 2, 20))
    out["logsubexp"] = dict(x=x.tolist(), y=y.tolist(), z=n2p.logsubexp(x, y).tolist())
    f = rng.normal(-5, 2, 40)
    s = -np.cumsum(rng.uniform(0.01, 0.2, 40))
    out["log_trap"] = dict(log_func=f.tolist(), log_support=s.tolist(),
                           value=fl(n2p.log_integrate_log_trap(f, s)))
    cw = []
    data = np.array([-5, -4, -3, -2.5, -2, -1.5, -1, -.5])
    ev, w = n2p.compute_weights(data, 4)
    cw.append(dict(tag="GV3", data=data.tolist(), nlive=4, log_ev=fl(ev), log_wts=w.tolist()))
    data = np.sort(rng.normal(-20, 5, 300))
    ev, w = n2p.compute_weights(data, 50)
    cw.append(dict(tag="rand", data=data.tolist(), nlive=50, log_ev=fl(ev), log_wts=w.tolist()))
    out["compute_weights"] = cw
    acls = []
    i = np.arange(64)
    xs = np.sin(0.3 * i) + 0.01 * i
    acls.append(dict(tag="GV4", x=xs.tolist(), acl=fl(n2p.acl(xs)), acf=n2p.autocorrelation(xs).tolist()))
    ar = np.zeros(500)
    e = rng.normal(size=500)
    for k in range(1, 500):
        ar[k] = 0.9 * ar[k - 1] + e[k]
    acls.append(dict(tag="ar1", x=ar.tolist(), acl=fl(n2p.acl(ar)), acf=n2p.autocorrelation(ar).tolist()))
    out["acl"] = acls
    dump("nest2pos.json", out)


# --------------------------------------------------------------------------------------
# random-tape recorder: patches the names the reference imported
# --------------------------------------------------------------------------------------
class Tape:
    """Replacement for the stdlib-random entry points the reference calls; records draws."""

    def __init__(self, seed):
        self.rng = pyrandom.Random(seed)
        self.events = []  # list of [kind, value(s)]

    # proposal.py:7  `from random import sample,gauss,randrange,uniform`
    def sample(self, population, k):
        idx = self.rng.sample(range(len(population)), k)
        self.events.append(["sample", idx])
        return [population[i] for i in idx]

    def gauss(self, mu, sigma):
        g = self.rng.gauss(0.0, 1.0)
        v = mu + sigma * g
        self.events.append(["gauss", v])
        return v

    def randrange(self, n):
        v = self.rng.randrange(n)
        self.events.append(["randrange", v])
        return v

    def uniform(self, a, b):
        v = self.rng.uniform(a, b)
        self.events.append(["uniform", v])
        return v

    # proposal.py:253 `random.choice(self.ensemble)`
    def choice(self, seq):
        i = self.rng.randrange(len(seq))
        self.events.append(["choice", i])
        return seq[i]

    # sampler.py:9 `from random import random`
    def random(self):
        v = self.rng.random()
        self.events.append(["random", v])
        return v


class _RandomModuleShim:
    def __init__(self, tape):
        self.choice = tape.choice


def install_tape(tape):
    rprop.sample = tape.sample
    rprop.gauss = tape.gauss
    rprop.randrange = tape.randrange
    rprop.uniform = tape.uniform
    rprop.random = _RandomModuleShim(tape)
    rsamp.random = tape.random


def make_points(names, X, logL=None, logP=None):
    pts = []
    for i, row in enumerate(X):
        p = LivePoint(list(names), d=array("d", [float(v) for v in row]))
        if logL is not None:
            p.logL = float(logL[i])
        if logP is not None:
            p.logP = float(logP[i])
        pts.append(p)
    return pts


# --------------------------------------------------------------------------------------
# C. the four ensemble proposals (cpnest/proposal.py:209-342)
# --------------------------------------------------------------------------------------
def gen_proposals():
    out = []
    for D, P, seed in [(1, 12, 1), (2, 16, 2), (3, 8, 7), (10, 40, 3), (50, 120, 4)]:
        rs = np.random.RandomState(seed)
        if seed == 7:
            # the GV5/GV6 ensemble of SURVEY.md section 8c
            X = np.array([np.random.RandomState(7).uniform(-1, 1, 3)])
            r7 = np.random.RandomState(7)
            X = np.array([r7.uniform(-1, 1, 3) for _ in range(8)])
        else:
            A = rs.normal(size=(D, D)) / math.sqrt(D)
            X = rs.normal(size=(P, D)) @ A.T + rs.uniform(-1, 1, D)
        names = ["p%d" % i for i in range(D)]
        ens = deque(make_points(names, X))
        case = dict(D=D, P=len(X), ensemble=X.tolist(), calls=[])
        eig = rprop.EnsembleEigenVector()
        eig.set_ensemble(ens)
        case["covariance"] = np.atleast_2d(eig.covariance).tolist()
        case["eigen_values"] = np.atleast_1d(eig.eigen_values).tolist()
        case["eigen_vectors"] = np.atleast_2d(eig.eigen_vectors).tolist()
        props = dict(walk=rprop.EnsembleWalk(), stretch=rprop.EnsembleStretch(),
                     de=rprop.DifferentialEvolution(), eigen=eig)
        for p in props.values():
            p.set_ensemble(ens)
        tape = Tape(1000 + seed)
        install_tape(tape)
        for rep in range(6):
            old = rs.normal(size=D)
            for kind, prop in props.items():
                if kind == "walk" and len(X) < 3:
                    continue
                tape.events = []
                oldp = make_points(names, [old])[0]
                new = prop.get_sample(oldp.copy())
                case["calls"].append(dict(kind=kind, old=old.tolist(), tape=tape.events,
                                          new=list(new.values), log_J=fl(prop.log_J)))
        out.append(case)
    dump("proposals.json", out)


# --------------------------------------------------------------------------------------
# D. proposal cycle (cpnest/proposal.py:71-75, 345-362)
# --------------------------------------------------------------------------------------
def gen_cycle():
    out = []
    letter = {rprop.EnsembleWalk: "W", rprop.EnsembleStretch: "S",
              rprop.DifferentialEvolution: "D", rprop.EnsembleEigenVector: "E"}
    for seed in [1234, 1, 42]:
        np.random.seed(seed)
        c = rprop.DefaultProposalCycle()
        s = "".join(letter[type(p)] for p in c.cycle)
        out.append(dict(seed=seed, cycle=s, weights=[fl(w) for w in c.weights]))
    dump("cycle.json", out)


# --------------------------------------------------------------------------------------
# F. the reference's own example / test models
# --------------------------------------------------------------------------------------
def _load_module(path, name):
    import importlib.util
    import types
    # examples import matplotlib in some files; stub it (absent here, unused by the models)
    for m in ["matplotlib", "matplotlib.pyplot"]:
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
            sys.modules[m].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


REFROOT = os.environ.get("REF", "/root/reference")


def reference_models():
    """name -> (model instance, functor description consumed by cpnest_b200)."""
    m = {}
    t2d = _load_module(os.path.join(REFROOT, "tests/test_2d.py"), "ref_test_2d")
    m["gaussian2d"] = t2d.GaussianModel()
    t50 = _load_module(os.path.join(REFROOT, "examples/test_50d.py"), "ref_test_50d")
    m["gaussian50d"] = t50.GaussianModel(dim=50)
    m["gaussian10d"] = t50.GaussianModel(dim=10)
    egg = _load_module(os.path.join(REFROOT, "examples/eggbox.py"), "ref_eggbox")
    m["eggbox"] = egg.EggboxModel()
    ros = _load_module(os.path.join(REFROOT, "examples/rosenbrock.py"), "ref_rosenbrock")
    m["rosenbrock2d"] = ros.RosenbrockModel(2)
    m["rosenbrock10d"] = ros.RosenbrockModel(10)
    ring = _load_module(os.path.join(REFROOT, "tests/test_ring.py"), "ref_ring")
    m["ring"] = ring.RingModel()
    np.random.seed(1234)
    tg = _load_module(os.path.join(REFROOT, "tests/test_gaussian.py"), "ref_test_gaussian")
    m["gaussian_mean_sigma"] = tg.GaussianModel()
    t1 = _load_module(os.path.join(REFROOT, "tests/test_1d.py"), "ref_test_1d")
    m["gaussian1d"] = t1.GaussianModel()
    th = _load_module(os.path.join(REFROOT, "tests/test_half_gaussian.py"), "ref_half")
    m["half_gaussian"] = th.HalfGaussianModel(xmax=20)
    np.random.seed(1234)  # the reference draws its data unseeded; SURVEY 8d fixes seed 1234
    gm = _load_module(os.path.join(REFROOT, "examples/gaussianmixture.py"), "ref_gmm")
    m["gaussian_mixture"] = gm.GaussianMixtureModel()
    return m


def gen_models():
    out = {}
    rs = np.random.RandomState(99)
    for name, model in reference_models().items():
        D = len(model.names)
        lo = np.array([b[0] for b in model.bounds], dtype=float)
        hi = np.array([b[1] for b in model.bounds], dtype=float)
        n = 24 if name != "gaussian_mixture" else 8
        X = lo + (hi - lo) * rs.uniform(size=(n, D))
        # a few points outside / on the boundary to pin in_bounds strictness (model.py:33)
        X[0, 0] = lo[0]
        X[1, -1] = hi[-1]
        X[2, 0] = hi[0] + 0.5
        if name in ("gaussian50d", "gaussian10d", "gaussian2d", "rosenbrock10d", "rosenbrock2d"):
            X[3:12] = rs.normal(size=(9, D))  # near the mode
        pts = make_points(model.names, X)
        logP, logL, inb = [], [], []
        for p in pts:
            inb.append(bool(model.in_bounds(p)))
            lp = model.log_prior(p)
            logP.append(fl(lp))
            with np.errstate(all="ignore"):
                ll = model.log_likelihood(p)
            logL.append(fl(np.asarray(ll).reshape(-1)[0]))
        case = dict(names=list(model.names), bounds=[[fl(a), fl(b)] for a, b in model.bounds],
                    X=X.tolist(), in_bounds=inb, logP=logP, logL=logL)
        if name == "gaussian_mixture":
            case["data"] = model.data.tolist()
        out[name] = case
    dump("models.json", out)


# --------------------------------------------------------------------------------------
# E. MetropolisHastingsSampler.yield_sample traces (cpnest/sampler.py:322-358)
# --------------------------------------------------------------------------------------
class _Val:
    def __init__(self, v):
        self.value = v


class _FakeManager:
    """Just enough of RunManager (cpnest/cpnest.py:553-584) for a Sampler to be built."""

    def __init__(self):
        self.logLmin = _Val(-np.inf)
        self.logLmax = _Val(-np.inf)
        self.checkpoint_flag = _Val(0)
        self.periodic_checkpoint_interval = np.inf
        self.nthreads = 1

    def connect_producer(self):
        return None, 0


def gen_mh_chain():
    models = reference_models()
    out = []
    letter = {rprop.EnsembleWalk: 0, rprop.EnsembleStretch: 1,
              rprop.DifferentialEvolution: 2, rprop.EnsembleEigenVector: 3}
    specs = [
        # (model, poolsize, Nmcmc, maxmcmc, nchains, logLmin quantile, scale of pool around mode)
        ("gaussian2d", 16, 12, 40, 4, 0.3, 3.0),
        ("gaussian_mean_sigma", 20, 10, 30, 4, 0.3, None),
        ("gaussian10d", 32, 15, 60, 3, 0.2, 1.5),
        ("rosenbrock2d", 24, 10, 50, 3, 0.4, 1.0),
        ("eggbox", 24, 10, 30, 3, 0.2, None),
        ("gaussian50d", 96, 30, 100, 2, 0.05, 1.2),
        ("gaussian2d", 16, 5, 8, 4, 0.97, 3.0),  # nearly impossible constraint: maxmcmc exit, no acceptance
    ]
    for si, (mname, P, Nmcmc, maxmcmc, nchains, q, scale) in enumerate(specs):
        model = models[mname]
        D = len(model.names)
        rs = np.random.RandomState(100 + si)
        lo = np.array([b[0] for b in model.bounds], dtype=float)
        hi = np.array([b[1] for b in model.bounds], dtype=float)
        if scale is None:
            X = lo + (hi - lo) * rs.uniform(size=(P, D))
        else:
            X = np.clip(rs.normal(size=(P, D)) * scale, lo + 1e-3, hi - 1e-3)
        pts = make_points(model.names, X)
        for p in pts:
            p.logP = float(model.log_prior(p))
            p.logL = float(np.asarray(model.log_likelihood(p)).reshape(-1)[0])
        logLs = np.array([p.logL for p in pts])
        logLmin = float(np.quantile(logLs, q))
        np.random.seed(1234 + si)
        cycle = rprop.DefaultProposalCycle()
        tape = Tape(5000 + si)
        install_tape(tape)
        mgr = _FakeManager()
        smp = rsamp.MetropolisHastingsSampler(model, maxmcmc, seed=1, poolsize=P, proposal=cycle, manager=mgr)
        smp.Nmcmc = Nmcmc
        for p in pts:
            smp.evolution_points.append(p)
        cycle.set_ensemble(smp.evolution_points)
        eig = [p for p in cycle.proposals if isinstance(p, rprop.EnsembleEigenVector)][0]
        case = dict(model=mname, D=D, P=P, Nmcmc=Nmcmc, maxmcmc=maxmcmc, logLmin=logLmin,
                    pool=X.tolist(), pool_logL=logLs.tolist(), pool_logP=[p.logP for p in pts],
                    cycle=[letter[type(p)] for p in cycle.cycle],
                    eigen_values=np.atleast_1d(eig.eigen_values).tolist(),
                    eigen_vectors=np.atleast_2d(eig.eigen_vectors).tolist(),
                    chains=[])
        gen = smp.yield_sample(logLmin)
        for c in range(nchains):
            tape.events = []
            idx0 = cycle.idx
            nsamp0 = len(smp.samples)
            steps, outp = next(gen)
            states = [list(s.values) for s in list(smp.samples)[nsamp0:]] if len(smp.samples) - nsamp0 == steps else None
            case["chains"].append(dict(cycle_idx_before=idx0, tape=tape.events, steps=steps,
                                       out=list(outp.values), out_logL=fl(outp.logL), out_logP=fl(outp.logP),
                                       sub_acceptance=fl(smp.sub_acceptance), states=states,
                                       logLmax=fl(mgr.logLmax.value)))
        out.append(case)
    dump("mh_chain.json", out)


# --------------------------------------------------------------------------------------
# H. SliceSampler.yield_sample traces (cpnest/sampler.py:414-569, cpnest/proposal.py:114-207)
# --------------------------------------------------------------------------------------
class _NpRandomTape:
    """Stands in for `np.random` inside cpnest.sampler / cpnest.proposal: draws from a private
    RandomState and records every value the reference receives."""

    def __init__(self, tape, seed):
        self.tape = tape
        self.rs = np.random.RandomState(seed)

    def uniform(self, low=0.0, high=1.0, size=None):
        v = self.rs.uniform(low, high, size)
        self.tape.events.append(["np_uniform", float(v)])
        return v

    def exponential(self, scale=1.0, size=None):
        v = self.rs.exponential(scale, size)
        self.tape.events.append(["np_exponential", float(v)])
        return v

    def normal(self, loc=0.0, scale=1.0, size=None):
        v = self.rs.normal(loc, scale, size)
        self.tape.events.append(["np_normal", np.atleast_1d(v).tolist()])
        return v

    def multivariate_normal(self, mean, cov):
        v = self.rs.multivariate_normal(mean, cov)
        self.tape.events.append(["np_mvn", np.atleast_1d(v).tolist()])
        return v

    def __getattr__(self, name):          # choice(), seed(), ... : not on the recorded path
        return getattr(np.random, name)


class _NpShim:
    """`np` as seen by the reference module: numpy, except for `.random`."""

    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, name):
        return getattr(np, name)


def gen_slice_chain():
    models = reference_models()
    out = []
    letter = {rprop.EnsembleSliceDifferential: 0, rprop.EnsembleSliceGaussian: 1,
              rprop.EnsembleSliceCorrelatedGaussian: 2}
    specs = [
        # (model, poolsize, Nmcmc, maxmcmc, nchains, logLmin quantile, scale of pool, mu)
        ("gaussian2d", 16, 4, 30, 4, 0.3, 3.0, 1.0),
        ("gaussian_mean_sigma", 20, 3, 24, 4, 0.3, None, 0.4),     # non-flat prior: stepping out stops on Yp
        ("gaussian10d", 32, 5, 40, 3, 0.2, 1.5, 0.7),
        ("rosenbrock2d", 24, 4, 30, 3, 0.4, 1.0, 1.5),
        ("ring", 24, 4, 30, 3, 0.5, None, 0.5),
        ("gaussian50d", 96, 3, 20, 2, 0.05, 1.2, 2.0),
        ("gaussian2d", 16, 2, 4, 4, 0.97, 3.0, 1.0),               # nearly impossible constraint: max_slices exits
    ]
    real_np_samp, real_np_prop = rsamp.np, rprop.np
    for si, (mname, P, Nmcmc, maxmcmc, nchains, q, scale, mu) in enumerate(specs):
        model = models[mname]
        D = len(model.names)
        rs = np.random.RandomState(300 + si)
        lo = np.array([b[0] for b in model.bounds], dtype=float)
        hi = np.array([b[1] for b in model.bounds], dtype=float)
        if scale is None:
            X = lo + (hi - lo) * rs.uniform(size=(P, D))
        else:
            X = np.clip(rs.normal(size=(P, D)) * scale, lo + 1e-3, hi - 1e-3)
        pts = make_points(model.names, X)
        for p in pts:
            p.logP = float(model.log_prior(p))
            p.logL = float(np.asarray(model.log_likelihood(p)).reshape(-1)[0])
        logLs = np.array([p.logL for p in pts])
        logLmin = float(np.quantile(logLs, q))
        np.random.seed(4321 + si)
        cycle = rprop.EnsembleSliceProposalCycle()
        tape = Tape(7000 + si)
        install_tape(tape)
        rnd = _NpRandomTape(tape, 9000 + si)
        rsamp.np = _NpShim(rnd)
        rprop.np = _NpShim(rnd)
        try:
            mgr = _FakeManager()
            smp = rsamp.SliceSampler(model, maxmcmc, seed=1, poolsize=P, proposal=cycle, manager=mgr)
            # what SliceSampler.reset sets (sampler.py:423-427), without its prior draw / burn-in
            smp.mu = mu
            smp.max_steps_out = smp.maxmcmc
            smp.max_slices = smp.maxmcmc
            smp.tuning_steps = 10 * P
            smp.Nmcmc = Nmcmc
            for p in pts:
                smp.evolution_points.append(p)
            cycle.set_ensemble(smp.evolution_points)
            corr = [p for p in cycle.proposals if isinstance(p, rprop.EnsembleSliceCorrelatedGaussian)][0]
            case = dict(model=mname, D=D, P=P, Nmcmc=Nmcmc, maxmcmc=maxmcmc, logLmin=logLmin, mu0=mu,
                        pool=X.tolist(), pool_logL=logLs.tolist(), pool_logP=[p.logP for p in pts],
                        cycle=[letter[type(p)] for p in cycle.cycle],
                        mean=np.atleast_1d(corr.mean).tolist(), covariance=np.atleast_2d(corr.covariance).tolist(),
                        chains=[])
            gen = smp.yield_sample(logLmin)
            for c in range(nchains):
                tape.events = []
                idx0 = cycle.idx
                mu_before = smp.mu
                steps, outp = next(gen)
                case["chains"].append(dict(cycle_idx_before=idx0, mu_before=fl(mu_before), mu_after=fl(smp.mu),
                                           tape=tape.events, steps=steps, out=list(outp.values),
                                           out_logL=fl(outp.logL), out_logP=fl(outp.logP),
                                           sub_acceptance=fl(smp.sub_acceptance), Ne=fl(smp.Ne), Nc=fl(smp.Nc),
                                           L=fl(smp.L), R=fl(smp.R), logLmax=fl(mgr.logLmax.value)))
        finally:
            rsamp.np, rprop.np = real_np_samp, real_np_prop
        out.append(case)
    dump("slice_chain.json", out)


# --------------------------------------------------------------------------------------
# I. HamiltonianMonteCarloSampler.yield_sample + ConstrainedLeapFrog traces
#    (cpnest/sampler.py:360-402, cpnest/proposal.py:377-844)
# --------------------------------------------------------------------------------------
class _NpRandomTapeHMC(_NpRandomTape):
    def choice(self, a, size=None, replace=True, p=None):
        """np.random.choice(range, p=p) as the legacy RandomState does it (one uniform, searchsorted on the
        normalised cumulative sum); the uniform is recorded so that the selection can be replayed."""
        if p is None or size is not None:
            return np.random.choice(a, size=size, replace=replace, p=p)
        a = list(a)
        u = self.rs.random_sample()
        cdf = np.cumsum(np.asarray(p, dtype=np.float64))
        cdf /= cdf[-1]
        k = int(cdf.searchsorted(u, side="right"))
        self.tape.events.append(["choice_u", float(u)])
        self.tape.events.append(["choice_idx", int(a[k])])
        return a[k]

    def randint(self, low, high=None, size=None):
        v = self.rs.randint(low, high, size)
        self.tape.events.append(["randint", int(v)])
        return v


class _RecordingMomenta:
    def __init__(self, dist, tape, rs):
        self.dist, self.tape, self.rs = dist, tape, rs

    def rvs(self):
        v = np.atleast_1d(self.dist.rvs(random_state=self.rs))
        self.tape.events.append(["p0", v.tolist()])
        return v


def gen_hmc_chain():
    models = reference_models()

    class MeanSigmaWithForce(type(models["gaussian_mean_sigma"])):
        # tests/test_gaussian.py defines no force; the analytic gradient of the potential -log_prior = log(sigma) + const
        def force(self, p):
            f = np.zeros(1, dtype={"names": p.names, "formats": ["f8" for _ in p.names]})
            f["sigma"] = 1.0 / p["sigma"]
            return f

    t2h = _load_module(os.path.join(REFROOT, "tests/test_2d_hmc.py"), "ref_test_2d_hmc")
    models = dict(models, gaussian2d=t2h.GaussianModel(), gaussian_mean_sigma=MeanSigmaWithForce())
    out = []
    specs = [
        # (model, poolsize, Nmcmc, nchains, logLmin quantile, scale of pool)
        ("gaussian2d", 64, 3, 4, 0.3, 3.0),
        ("gaussian_mean_sigma", 64, 3, 4, 0.3, None),
        ("gaussian10d", 96, 2, 3, 0.2, 1.5),
        ("rosenbrock2d", 64, 3, 3, 0.4, 1.0),
        ("gaussian2d", 64, 2, 3, 0.9, 3.0),          # tight constraint: many reflections, rejected trajectories
    ]
    real_np_samp, real_np_prop = rsamp.np, rprop.np
    for si, (mname, P, Nmcmc, nchains, q, scale) in enumerate(specs):
        model = models[mname]
        D = len(model.names)
        rs = np.random.RandomState(500 + si)
        lo = np.array([b[0] for b in model.bounds], dtype=float)
        hi = np.array([b[1] for b in model.bounds], dtype=float)
        if scale is None:
            X = lo + (hi - lo) * rs.uniform(size=(P, D))
        else:
            X = np.clip(rs.normal(size=(P, D)) * scale, lo + 1e-3, hi - 1e-3)
        pts = make_points(model.names, X)
        for p in pts:
            p.logP = float(model.log_prior(p))
            p.logL = float(np.asarray(model.log_likelihood(p)).reshape(-1)[0])
        logLs = np.array([p.logL for p in pts])
        logLmin = float(np.quantile(logLs, q))
        # the chains start from pool members above the threshold (the reference pops the right end of the deque)
        order = np.argsort(logLs)
        pts = [pts[i] for i in order]
        X, logLs = X[order], logLs[order]
        np.random.seed(2468 + si)
        cycle = rprop.HamiltonianProposalCycle(model=model)
        prop = cycle.proposals[0]
        tape = Tape(11000 + si)
        install_tape(tape)
        rnd = _NpRandomTapeHMC(tape, 12000 + si)
        rsamp.np = _NpShim(rnd)
        rprop.np = _NpShim(rnd)
        try:
            mgr = _FakeManager()
            smp = rsamp.HamiltonianMonteCarloSampler(model, 100, seed=1, poolsize=P, proposal=cycle, manager=mgr)
            smp.Nmcmc = Nmcmc
            for p in pts:
                smp.evolution_points.append(p)
            cycle.set_ensemble(smp.evolution_points)
            prop.momenta_distribution = _RecordingMomenta(prop.momenta_distribution, tape, rnd.rs)
            real_normal = prop.unit_normal

            def rec_normal(qq, _f=real_normal):
                n = _f(qq)
                tape.events.append(["normal", np.asarray(n, dtype=float).tolist()])
                return n
            prop.unit_normal = rec_normal
            case = dict(model=mname, D=D, P=P, Nmcmc=Nmcmc, logLmin=logLmin,
                        pool=X.tolist(), pool_logL=logLs.tolist(), pool_logP=[p.logP for p in pts],
                        covariance=np.atleast_2d(prop.covariance).tolist(),
                        inverse_mass=np.atleast_1d(prop.inverse_mass).tolist(),
                        logdeterminant=fl(prop.logdeterminant), base_dt=fl(prop.base_dt), base_L=int(prop.base_L),
                        chains=[])
            gen = smp.yield_sample(logLmin)
            for c in range(nchains):
                tape.events = []
                dt0, L0, scale0 = fl(prop.dt), int(prop.L), fl(prop.scale)
                start = smp.evolution_points[-1]
                steps, outp = next(gen)
                case["chains"].append(dict(dt=dt0, L=L0, scale=scale0, start=list(start.values), start_logL=fl(start.logL),
                                           start_logP=fl(start.logP), tape=tape.events, steps=steps,
                                           out=list(outp.values), out_logL=fl(outp.logL), out_logP=fl(outp.logP),
                                           log_J=fl(prop.log_J), acceptance=fl(smp.acceptance),
                                           dt_after=fl(prop.dt), L_after=int(prop.L), scale_after=fl(prop.scale)))
        finally:
            rsamp.np, rprop.np = real_np_samp, real_np_prop
        out.append(case)
    dump("hmc_chain.json", out)


# --------------------------------------------------------------------------------------
# G. NestedSampler.consume_sample / final sweep with scripted samplers
#    (cpnest/NestedSampling.py:318-375, 433-493)
# --------------------------------------------------------------------------------------
class _ScriptPipe:
    def __init__(self, script):
        self.script = script
        self.sent = []

    def send(self, obj):
        self.sent.append(obj)

    def recv(self):
        return self.script.pop(0)


def gen_ns_loop(tmpdir="/tmp/cpnest_golden_ns"):
    os.makedirs(tmpdir, exist_ok=True)
    out = []
    for nlive, nthreads, nbatches, seed in [(20, 1, 60, 3), (64, 4, 40, 4), (200, 8, 100, 5)]:
        rs = np.random.RandomState(seed)
        names = ["a", "b"]
        mgr = _FakeManager()
        mgr.nthreads = nthreads

        class M(cpnest.model.Model):
            pass
        M.names = names
        M.bounds = [[-10, 10], [-10, 10]]
        model = M()
        ns = NestedSampler(model, manager=mgr, nlive=nlive, output=tmpdir, verbose=0, seed=seed)
        X0 = rs.uniform(-10, 10, size=(nlive, 2))
        L0 = -0.5 * (X0 ** 2).sum(1)
        from cpnest.NestedSampling import OrderedLivePoints
        ns.params = OrderedLivePoints(make_points(names, X0, logL=L0, logP=np.zeros(nlive)))
        ns.initialised = True
        mgr.logLmax.value = float(L0.max())
        case = dict(nlive=nlive, nthreads=nthreads, X0=X0.tolist(), L0=L0.tolist(), batches=[])
        for b in range(nbatches):
            cur = np.array([p.logL for p in ns.params])
            thr = cur[nthreads - 1]
            # scripted replacements: each pipe returns, in order of use, some rejected (logL<=thr)
            # and then an accepted point.  consume_sample recv()s from pipe queue_counter, k reversed.
            pipes = [_ScriptPipe([]) for _ in range(nthreads)]
            qc = ns.queue_counter
            repl = []
            for k in reversed(range(nthreads)):
                seq = []
                if rs.uniform() < 0.3:
                    r = np.sqrt(-2 * (thr - abs(rs.normal()) - 0.01))
                    th = rs.uniform(0, 2 * np.pi)
                    x = np.array([r * np.cos(th), r * np.sin(th)])
                    seq.append(x)
                r = np.sqrt(max(0.0, -2 * thr)) * np.sqrt(rs.uniform())
                th = rs.uniform(0, 2 * np.pi)
                seq.append(np.array([r * np.cos(th), r * np.sin(th)]))
                for x in seq:
                    p = make_points(names, [x], logL=[-0.5 * (x ** 2).sum()], logP=[0.0])[0]
                    pipes[qc].script.append((0.5, 0.5, 10, p))
                    repl.append([k, x.tolist(), p.logL])
                qc = (qc + 1) % nthreads
                mgr.logLmax.value = max(mgr.logLmax.value, p.logL)
            mgr.consumer_pipes = pipes
            n_ins0 = len(ns.insertion_indices)
            ns.consume_sample()
            case["batches"].append(dict(
                logLmin=fl(mgr.logLmin.value), replacements=repl,
                logZ=fl(ns.state.logZ), info=fl(ns.state.info), logw=fl(ns.state.logw),
                condition=fl(ns.condition), iteration=ns.iteration,
                insertion_indices=[fl(v) for v in ns.insertion_indices[n_ins0:]],
                live_logL=[p.logL for p in ns.params] if (b % 10 == 9 or b == nbatches - 1) else None))
        # final sweep (NestedSampling.py:472-480)
        from operator import attrgetter
        ns.params.sort(key=attrgetter("logL"))
        for i, p in enumerate(ns.params):
            ns.state.increment(p.logL, nlive=ns.Nlive - i)
            ns.nested_samples.append(p)
        case["final_logZ_rect"] = fl(ns.state.logZ)
        case["final_info"] = fl(ns.state.info)
        case["final_logZ"] = fl(ns.state.finalise())
        case["nested_logL"] = [p.logL for p in ns.nested_samples]
        out.append(case)
    dump("ns_loop.json", out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["integrator", "nest2pos", "proposals", "cycle", "models", "mh_chain", "slice_chain", "hmc_chain", "ns_loop"]
    for w in which:
        globals()["gen_" + w]()
