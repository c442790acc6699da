"""CPU oracle: a restatement of the reference's hot path (johnveitch/cpnest).

*** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  Nothing under cpnest_b200/
imports it and the product has no CPU fallback: it fails loudly when the CUDA library is
missing.

Pinning: the reference ships no golden vectors for this path, so
this oracle is pinned against outputs of the reference itself: tests/golden/*.json, made
by tests/golden/gen_golden.py from the unmodified reference built in oracle/_ref
(tests/test_oracle_golden.py checks every function here against them at 1e-12 or
bit-exactly).  Citations are `path:line` relative to the reference tree.

Conventions: a point is a 1-D float64 numpy array; an ensemble is an (n, D) array.  Every
random number is an explicit argument (the reference draws them from the unseeded stdlib
`random`; fixtures record them), so each function is deterministic.
"""
from __future__ import annotations

import math

import numpy as np

LOG2 = math.log(2.0)

# proposal ids, in the order of DefaultProposalCycle (cpnest/proposal.py:354-361)
WALK, STRETCH, DE, EIGEN = 0, 1, 2, 3
DEFAULT_WEIGHTS = (5, 5, 10, 10)


# ----------------------------------------------------------------------------------------
# LivePoint scalar semantics (cpnest/parameter.pyx:96-120)
# ----------------------------------------------------------------------------------------
def f32(s):
    """`LivePoint.__mul__/__truediv__` declare the scalar operand as C `float`
    (parameter.pyx:96,103,109,116): the Python float is rounded to fp32 and widened back
    to fp64 before it multiplies the fp64 coordinates."""
    return float(np.float32(s))


def lp_mul(v, s, compat=True):
    return (f32(s) if compat else float(s)) * np.asarray(v, dtype=np.float64)


def lp_div(v, s, compat=True):
    return np.asarray(v, dtype=np.float64) / (f32(s) if compat else float(s))


# ----------------------------------------------------------------------------------------
# Model contract (cpnest/model.py:20-33, 66-82)
# ----------------------------------------------------------------------------------------
def in_bounds(x, lo, hi):
    """model.py:33 -- strict on both sides."""
    x = np.asarray(x)
    return bool(np.all((lo < x) & (x < hi)))


def flat_log_prior(x, lo, hi):
    """model.py:66-82"""
    return 0.0 if in_bounds(x, lo, hi) else -np.inf


class OracleModel:
    """Restatement of one of the reference's test/example models: bounds + log_prior +
    log_likelihood, as plain functions of a float64 vector.  `functor` is the name of the
    device functor in cpnest_b200 that must agree with it."""

    def __init__(self, functor, names, bounds, loglike, logprior=None, params=None):
        self.functor = functor
        self.names = list(names)
        self.bounds = [list(map(float, b)) for b in bounds]
        self.lo = np.array([b[0] for b in self.bounds])
        self.hi = np.array([b[1] for b in self.bounds])
        self._ll = loglike
        self._lp = logprior
        self.params = params or {}

    @property
    def D(self):
        return len(self.names)

    def in_bounds(self, x):
        return in_bounds(x, self.lo, self.hi)

    def log_prior(self, x):
        if not self.in_bounds(x):
            return -np.inf
        return 0.0 if self._lp is None else float(self._lp(np.asarray(x, dtype=np.float64)))

    def log_likelihood(self, x):
        with np.errstate(all="ignore"):
            return float(self._ll(np.asarray(x, dtype=np.float64)))


def _gauss_iid(x):
    # examples/test_50d.py:18-19 ; tests/test_2d.py:15-16 is the D=2 case of the same sum
    return float(np.sum(-0.5 * x * x - 0.5 * math.log(2.0 * math.pi)))


def _eggbox(x):
    # examples/eggbox.py:19-23
    t = 1.0
    for v in x:
        t *= math.cos(v / 2.0)
    return (t + 2.0) ** 5.0


def _rosenbrock(x):
    # examples/rosenbrock.py:20-22
    return float(-np.sum(100.0 * (x[1:] - x[:-1] ** 2.0) ** 2.0 + (1.0 - x[:-1]) ** 2.0))


def _ring(x, radius=1.0, thickness=0.1):
    # tests/test_ring.py:19-23
    r = math.sqrt(x[0] ** 2 + x[1] ** 2)
    return -0.5 * (radius - r) ** 2 / thickness ** 2


def _gauss_mean_sigma_ll(x):
    # tests/test_gaussian.py:15-16
    return -0.5 * x[0] ** 2 / x[1] ** 2 - math.log(x[1]) - 0.5 * math.log(2.0 * math.pi)


def _gauss_mean_sigma_lp(x):
    # tests/test_gaussian.py:18-20
    return -math.log(x[1]) - math.log(10) - math.log(0.95)


def _norm_logpdf(x):
    # tests/test_1d.py:20-21, tests/test_half_gaussian.py:24-25 (scipy.stats.norm.logpdf)
    return -0.5 * x[0] * x[0] - 0.5 * math.log(2.0 * math.pi)


def _ackley_std(x):
    # examples/ackley.py:21-28 is NaN/-inf everywhere, so config 3
    # uses the standard D-dimensional Ackley function, logL = -ackley(x).  No reference parity.
    D = len(x)
    return -(-20.0 * math.exp(-0.2 * math.sqrt(float(np.sum(x * x)) / D))
             - math.exp(float(np.sum(np.cos(2.0 * math.pi * x))) / D) + 20.0 + math.e)


def make_mixture(data):
    data = np.asarray(data, dtype=np.float64)

    def ll(x):
        # examples/gaussianmixture.py:22-27
        w = x[4]
        l1 = np.log(w) - np.log(x[1]) - 0.5 * ((data - x[0]) / x[1]) ** 2
        l2 = np.log(1.0 - w) - np.log(x[3]) - 0.5 * ((data - x[2]) / x[3]) ** 2
        return np.logaddexp(l1, l2).sum()

    def lp(x):
        # examples/gaussianmixture.py:29-31
        return -math.log(x[1]) - math.log(x[3])

    return OracleModel("gaussian_mixture", ["mean1", "sigma1", "mean2", "sigma2", "weight"],
                       [[-3, 3], [0.01, 1], [-3, 3], [0.01, 1], [0.0, 1.0]], ll, lp, params=dict(data=data))


def make_correlated_gaussian(D, seed=1234, lo=-10.0, hi=10.0):
    """SURVEY.md section 8d config 4 ('correlated' flavour): N(0, Sigma), Sigma = A A^T with a
    fixed seed and condition number ~1e2; logZ stays -D log(hi-lo).  Not in the reference;
    logL = -0.5 |L^-1 x|^2 - sum(log diag L) - D/2 log 2pi with L = chol(Sigma)."""
    rs = np.random.RandomState(seed)
    Q, _ = np.linalg.qr(rs.normal(size=(D, D)))
    ev = np.exp(np.linspace(math.log(0.1), math.log(1.0), D)) ** 2  # sigma in [0.1, 1]
    Sigma = (Q * ev) @ Q.T
    Lc = np.linalg.cholesky(Sigma)
    Linv = np.linalg.inv(Lc)
    logdet = float(np.sum(np.log(np.diag(Lc))))

    def ll(x):
        y = Linv @ x
        return -0.5 * float(y @ y) - logdet - 0.5 * D * math.log(2.0 * math.pi)

    return OracleModel("gaussian_corr", [str(i) for i in range(D)], [[lo, hi]] * D, ll,
                       params=dict(Linv=np.tril(Linv), logdet=logdet))


def make_mixture_nd(D, w=0.3, m1=-2.5, s1=1.0, m2=2.5, s2=1.0, lo=-10.0, hi=10.0):
    """BASELINE.json config 5 names a "20-D Gaussian mixture" that the reference does not contain: two
    isotropic Gaussians in parameter space, L = w N(x; m1 1, s1^2 I) + (1-w) N(x; m2 1, s2^2 I); logZ = -D log(hi-lo)."""
    def ll(x):
        l1 = math.log(w) - 0.5 * float(np.sum(((x - m1) / s1) ** 2)) - D * (math.log(s1) + 0.5 * math.log(2 * math.pi))
        l2 = math.log(1 - w) - 0.5 * float(np.sum(((x - m2) / s2) ** 2)) - D * (math.log(s2) + 0.5 * math.log(2 * math.pi))
        return float(np.logaddexp(l1, l2))
    return OracleModel("gaussian_mixture_nd", [str(i) for i in range(D)], [[lo, hi]] * D, ll,
                       params=dict(w=w, m1=m1, s1=s1, m2=m2, s2=s2))


def builtin_model(name, **kw):
    """The reference's models by fixture name (tests/golden/models.json keys)."""
    if name.startswith("gaussian") and name.endswith("d") and name[8:-1].isdigit():
        D = int(name[8:-1])
        if D == 1:
            return OracleModel("gaussian_iid", ["x"], [[-10, 10]], _norm_logpdf)
        names = ["x", "y"] if D == 2 else [str(i) for i in range(D)]
        return OracleModel("gaussian_iid", names, [[-10, 10]] * D, _gauss_iid)
    if name == "half_gaussian":
        return OracleModel("gaussian_iid", ["x"], [[0, kw.get("xmax", 20)]], _norm_logpdf)
    if name == "eggbox":
        return OracleModel("eggbox", ["1", "2"], [[0, 10.0 * math.pi]] * 2, _eggbox)
    if name.startswith("rosenbrock"):
        D = int(name[10:-1])
        return OracleModel("rosenbrock", ["x%d" % i for i in range(D)], [[-5, 5]] * D, _rosenbrock)
    if name.startswith("ackley"):
        D = int(name[6:-1])
        return OracleModel("ackley", ["x%d" % i for i in range(D)], [[-5, 5]] * D, _ackley_std)
    if name == "ring":
        return OracleModel("ring", ["x", "y"], [[-2, 2]] * 2, _ring)
    if name == "gaussian_mean_sigma":
        return OracleModel("gaussian_mean_sigma", ["mean", "sigma"], [[-5, 5], [0.05, 1]],
                           _gauss_mean_sigma_ll, _gauss_mean_sigma_lp)
    if name == "gaussian_mixture":
        return make_mixture(kw["data"])
    raise KeyError(name)


# ----------------------------------------------------------------------------------------
# Proposals (cpnest/proposal.py:209-342)
# ----------------------------------------------------------------------------------------
def ensemble_walk(old, ens, idx3, g3, compat=True):
    """proposal.py:230-235.  idx3: the 3 distinct indices `sample` drew; g3: the 3 N(0,1)."""
    s = [np.asarray(ens[i], dtype=np.float64) for i in idx3]
    com = lp_div((s[0] + s[1]) + s[2], 3.0, compat)           # reduce(__add__) / float(Npoints)
    out = np.array(old, dtype=np.float64)
    for x, g in zip(s, g3):
        out = out + lp_mul(x - com, g, compat)                 # out += (x - com) * gauss(0,1)
    return out, 0.0


def ensemble_stretch(old, ens, ia, u, compat=True):
    """proposal.py:251-261.  u = uniform(-1,1); log_J uses the un-truncated x."""
    a = np.asarray(ens[ia], dtype=np.float64)
    x = u * math.log(2.0)
    Z = math.exp(x)
    out = a + lp_mul(np.asarray(old, dtype=np.float64) - a, Z, compat)
    return out, len(out) * x


def differential_evolution(old, ens, ia, ib, g, compat=True):
    """proposal.py:284-287.  g = gauss(1.0, 1e-4) (the value actually drawn)."""
    a = np.asarray(ens[ia], dtype=np.float64)
    b = np.asarray(ens[ib], dtype=np.float64)
    return np.asarray(old, dtype=np.float64) + lp_mul(b - a, g, compat), 0.0


def ensemble_eigenvector(old, eigen_values, eigen_vectors, i, g):
    """proposal.py:336-342 (full fp64: does not go through LivePoint.__mul__)."""
    jumpsize = math.sqrt(math.fabs(eigen_values[i])) * g
    out = np.array(old, dtype=np.float64)
    for k in range(len(out)):
        out[k] += jumpsize * eigen_vectors[k][i]
    return out, 0.0


def ensemble_covariance(ens):
    """proposal.py:311-323: np.cov of the D x n array (unbiased); D == 1 uses np.var (ddof=0)
    (proposal.py:314-318).  Returns (mean, covariance, eigen_values, eigen_vectors)."""
    X = np.asarray(ens, dtype=np.float64)
    n, D = X.shape
    mean = X.mean(axis=0)
    if D == 1:
        var = np.atleast_1d(np.var(X[:, 0]))
        return mean, np.atleast_2d(var), var, np.eye(1)
    cov = np.cov(X.T)
    w, v = np.linalg.eigh(cov)
    return mean, cov, w, v


def make_cycle(seed_state, weights=DEFAULT_WEIGHTS, cyclelength=100):
    """proposal.py:71-75: np.random.choice(proposals, size, p=weights) on the legacy global
    RandomState.  `seed_state` is a np.random.RandomState positioned where the reference's
    global state is when ProposalCycle.__init__ runs (right after np.random.seed(seed) for
    sampler 0)."""
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    return seed_state.choice(len(w), size=cyclelength, p=w, replace=True).astype(np.int32)


# ----------------------------------------------------------------------------------------
# MetropolisHastingsSampler.yield_sample (cpnest/sampler.py:322-358)
# ----------------------------------------------------------------------------------------
class TapeReader:
    """Reads the event tape recorded by tests/golden/gen_golden.py."""

    def __init__(self, events):
        self.ev = list(events)
        self.i = 0

    def take(self, kind):
        k, v = self.ev[self.i]
        assert k == kind, "tape: expected %s got %s at %d" % (kind, k, self.i)
        self.i += 1
        return v

    def exhausted(self):
        return self.i == len(self.ev)


def mh_chain(model, pool, pool_logL, pool_logP, cycle, cycle_idx, eigen_values, eigen_vectors,
             tape, Nmcmc, maxmcmc, logLmin, compat=True, record=False):
    """One `next(yield_sample(logLmin))` of the reference's MH sampler.

    `pool*` are python lists mirroring the sampler's deque `evolution_points`: the chain
    pops the OLDEST member (sampler.py:328), draws proposals from the remaining members
    (the deque is the ensemble, proposal.py:230,253,284), and appends the result (:351).
    The lists are modified in place.  Returns dict(out, logL, logP, steps, accepted,
    cycle_idx, states)."""
    tp = tape if isinstance(tape, TapeReader) else TapeReader(tape)
    old = np.array(pool.pop(0), dtype=np.float64)
    old_logL = pool_logL.pop(0)
    pool_logP.pop(0)
    logp_old = model.log_prior(old)                                     # :329
    cur_logP = logp_old
    ens = pool
    D = len(old)
    steps = accepted = 0
    states = []
    logLmax = -np.inf
    while True:
        steps += 1                                                       # :333
        cycle_idx = (cycle_idx + 1) % len(cycle)                         # proposal.py:93 (pre-increment)
        kind = int(cycle[cycle_idx])
        if kind == WALK:
            idx = tp.take("sample")
            g = [tp.take("gauss") for _ in range(3)]
            new, log_J = ensemble_walk(old, ens, idx, g, compat)
        elif kind == STRETCH:
            ia = tp.take("choice")
            u = tp.take("uniform")
            new, log_J = ensemble_stretch(old, ens, ia, u, compat)
        elif kind == DE:
            ia, ib = tp.take("sample")
            g = tp.take("gauss")
            new, log_J = differential_evolution(old, ens, ia, ib, g, compat)
        else:
            i = tp.take("randrange")
            g = tp.take("gauss")
            new, log_J = ensemble_eigenvector(old, eigen_values, eigen_vectors, i, g)
        new_logP = model.log_prior(new)                                  # :335
        u = tp.take("random")
        with np.errstate(all="ignore"):
            test = (new_logP - logp_old + log_J) > (math.log(u) if u > 0 else -np.inf)   # :337
        if test:
            new_logL = model.log_likelihood(new)                         # :338
            if new_logL > logLmin:                                       # :339
                logLmax = max(logLmax, new_logL)
                old, old_logL, cur_logP = new, new_logL, new_logP
                logp_old = new_logP
                accepted += 1
        if record:
            states.append(old.copy())                                    # :345
        if (steps >= Nmcmc and accepted > 0) or steps >= maxmcmc:        # :347
            break
    pool.append(old)
    pool_logL.append(old_logL)
    pool_logP.append(cur_logP)
    return dict(out=old, logL=old_logL, logP=cur_logP, steps=steps, accepted=accepted,
                cycle_idx=cycle_idx, states=states, logLmax=logLmax)


# ----------------------------------------------------------------------------------------
# SliceSampler.yield_sample (cpnest/sampler.py:414-569) with the direction proposals of
# EnsembleSliceProposalCycle (cpnest/proposal.py:114-207)
# ----------------------------------------------------------------------------------------
SLICE_DIFF, SLICE_GAUSS, SLICE_CORR = 0, 1, 2      # order of EnsembleSliceProposalCycle (proposal.py:196)


def slice_direction(kind, ens, mu, tp, compat=True):
    """proposal.py:126-133 (differential: `reduce(__sub__, sample(ens, 2)) * mu`, a LivePoint times a
    scalar, i.e. mu rounded to fp32), :180-187 (unit Gaussian direction, numpy: full fp64) and
    :167-173 (`mu * multivariate_normal(mean, cov)`; the mean is NOT subtracted).  For the two
    numpy kinds the tape holds the vector numpy returned (np.random.normal / multivariate_normal)."""
    if kind == SLICE_DIFF:
        ia, ib = tp.take("sample")
        d = np.asarray(ens[ia], dtype=np.float64) - np.asarray(ens[ib], dtype=np.float64)
        return lp_mul(d, mu, compat)
    if kind == SLICE_GAUSS:
        z = np.array(tp.take("np_normal"), dtype=np.float64)
        z = z / np.linalg.norm(z)
        return z * mu
    z = np.array(tp.take("np_mvn"), dtype=np.float64)
    return mu * z


def slice_chain(model, pool, pool_logL, pool_logP, cycle, cycle_idx, mu, tape, Nmcmc, maxmcmc, logLmin,
                compat=True, mcmc_counter=0, tuning_steps=None):
    """One `next(yield_sample(logLmin))` of the reference's SliceSampler.  `pool*` mirror the
    deque `evolution_points` and are modified in place.  Returns dict(out, logL, logP, steps,
    accepted, cycle_idx, mu, Ne, Nc, L, R, evals)."""
    tp = tape if isinstance(tape, TapeReader) else TapeReader(tape)
    P = len(pool)
    max_steps_out = max_slices = maxmcmc                                  # sampler.py:424-425
    # :471-478 first pool member above the threshold (at most poolsize pops; the last popped one is
    # kept even if it fails, exactly as the reference does)
    j = 0
    while j < P:
        old = np.array(pool.pop(0), dtype=np.float64)
        old_logL = pool_logL.pop(0)
        old_logP = pool_logP.pop(0)
        if old_logL > logLmin:
            break
        pool.append(old); pool_logL.append(old_logL); pool_logP.append(old_logP)
        j += 1
    ens = pool
    sub_counter = sub_accepted = evals = 0
    Ne = Nc = 0.0
    L = R = 0.0
    while True:
        L = -tp.take("np_uniform")                                        # reset_boundaries :444-447
        R = L + 1.0
        Ne = Nc = 0.0
        sub_counter += 1
        cycle_idx = (cycle_idx + 1) % len(cycle)                          # proposal.py:203
        direction = slice_direction(int(cycle[cycle_idx]), ens, mu, tp, compat)
        Y = logLmin
        Yp = old_logP - tp.take("np_exponential")                         # :490
        J = math.floor(max_steps_out * tp.take("np_uniform"))             # :491
        K = (max_steps_out - 1) - J
        while J > 0:                                                      # :495-508
            left = lp_mul(direction, L, compat) + old
            if model.in_bounds(left):
                if Yp > model.log_prior(left):
                    break
                L -= 1.0; Ne += 1; J -= 1
            else:
                break
        while K > 0:                                                      # :512-524
            right = lp_mul(direction, R, compat) + old
            if model.in_bounds(right):
                if Yp > model.log_prior(right):
                    break
                R += 1.0; Ne += 1; K -= 1
            else:
                break
        sl = 0
        while sl < max_slices:                                            # :529-551
            Xp = tp.take("np_uniform")                                    # np.random.uniform(L, R): value drawn
            new = lp_mul(direction, Xp, compat) + old
            new_logP = model.log_prior(new)
            if new_logP > Yp:
                new_logL = model.log_likelihood(new)
                evals += 1
                if new_logL > Y:
                    old, old_logL, old_logP = new, new_logL, new_logP
                    sub_accepted += 1
                    break
            if Xp < 0.0:
                L = Xp; Nc += 1
            elif Xp > 0.0:
                R = Xp; Nc += 1
            sl += 1
        if sub_counter > Nmcmc and sub_accepted > 0:                      # :553-554
            break
        if sub_counter > maxmcmc:                                         # :556-557
            break
    pool.append(old); pool_logL.append(old_logL); pool_logP.append(old_logP)
    mcmc_counter += sub_counter
    if tuning_steps is None:
        tuning_steps = 10 * (P)                                           # :426
    if mcmc_counter < tuning_steps:                                       # :566-567, adapt_length_scale :429-437
        mu = mu * (2 * (max(1, Ne) / (max(1, Ne) + max(1, Nc))))
    return dict(out=old, logL=old_logL, logP=old_logP, steps=sub_counter, accepted=sub_accepted,
                cycle_idx=cycle_idx, mu=mu, Ne=Ne, Nc=Nc, L=L, R=R, evals=evals, mcmc_counter=mcmc_counter)


# ----------------------------------------------------------------------------------------
# HamiltonianMonteCarloSampler.yield_sample (cpnest/sampler.py:360-402) driving
# ConstrainedLeapFrog (cpnest/proposal.py:377-844)
# ----------------------------------------------------------------------------------------
def hmc_set_integration_parameters(bounds):
    """proposal.py:524-536 -> (base_L, base_dt)"""
    ranges = [b[1] - b[0] for b in bounds]
    d = len(bounds)
    l, h = np.min(ranges), np.max(ranges)
    base_L = 10 + int((h / l) * d ** (1. / 4.))
    base_dt = (1.0 / base_L) * l / h
    return base_L, base_dt


def hmc_kinetic_energy(p, C, logdet):
    """proposal.py:562-575: 0.5 p.C.p - logdet(M) - d/2 log 2pi, C = ensemble covariance = inverse mass matrix"""
    return 0.5 * np.dot(p, np.dot(C, p)) - logdet - 0.5 * len(p) * np.log(2.0 * np.pi)


def hmc_trajectory(model, grad_V, q0, p0, C, inverse_mass, logdet, dt, L, logLmin, tp):
    """ConstrainedLeapFrog.evolve_trajectory + sample_trajectory (proposal.py:793-844).  `grad_V(x)` is
    Model.force (the gradient of the potential -log_prior); the unit normals used at reflections off
    the likelihood boundary are taken from the tape (the reference estimates them with per-dimension
    spline fits over the ensemble, proposal.py:417-486).  Returns (q, p, logL, logP) of the picked point."""
    lo, hi = model.lo, model.hi
    D = len(q0)
    traj = [(np.array(q0), np.array(p0), None, None)]

    def position(p, q):                                                   # :729-741
        for j in range(D):
            q[j] += dt * p[j] * inverse_mass[j]
            while q[j] < lo[j] or q[j] > hi[j]:
                if q[j] > hi[j]:
                    q[j] = hi[j] - (q[j] - hi[j])
                    p[j] *= -1
                if q[j] < lo[j]:
                    q[j] = lo[j] + (lo[j] - q[j])
                    p[j] *= -1

    def momentum(p, q):                                                   # :760-773
        dV = grad_V(q)
        logP = model.log_prior(q)                                         # check_constraint :790-791
        logL = model.log_likelihood(q)
        if logL - logLmin > 0:
            p += -dt * dV
        else:
            normal = np.array(tp.take("normal"))
            p += -2.0 * np.dot(p, normal) * normal
        return logL, logP

    for sign in (1.0, -1.0):
        p = sign * np.array(p0, dtype=np.float64)
        q = np.array(q0, dtype=np.float64)
        p += -0.5 * dt * grad_V(q)                                        # half step :757-759
        for _ in range(L):
            position(p, q)
            logL, logP = momentum(p, q)
            traj.append((q.copy(), p.copy(), logL, logP))
    # (the closing half step of :829 changes p only; the last point is excluded from the draw below)
    with np.errstate(all="ignore"):
        logw = np.array([-(hmc_kinetic_energy(pp, C, logdet) - model.log_prior(qq)) for qq, pp, _, _ in traj[1:-1]])
        m = np.max(logw)
        norm = m + np.log(np.sum(np.exp(logw - m)))                       # scipy logsumexp
        prob = np.exp(logw - norm)
    u = tp.take("choice_u")
    cdf = np.cumsum(prob)
    cdf /= cdf[-1]
    idx = 1 + int(cdf.searchsorted(u, side="right"))                      # np.random.choice(range(1, n-1), p=prob)
    assert idx == tp.take("choice_idx")
    return traj[idx]


def hmc_chain(model, grad_V, start, start_logL, start_logP, C, inverse_mass, logdet, dt, L, logLmin, tape,
              max_trajectories=10**9):
    """One `next(yield_sample(logLmin))` of the reference's HMC sampler: trajectories from `start` until
    one is accepted (sampler.py:375-385).  Returns dict(out, logL, logP, steps, log_J, evals)."""
    tp = tape if isinstance(tape, TapeReader) else TapeReader(tape)
    q0 = np.array(start, dtype=np.float64)
    steps = 0
    while True:
        steps += 1
        p0 = np.array(tp.take("p0"), dtype=np.float64)
        E0 = hmc_kinetic_energy(p0, C, logdet) - model.log_prior(q0)      # hamiltonian(p0, q0) :621
        q, p, logL, logP = hmc_trajectory(model, grad_V, q0, p0, C, inverse_mass, logdet, dt, L, logLmin, tp)
        E1 = hmc_kinetic_energy(p, C, logdet) - model.log_prior(q)        # :625
        with np.errstate(all="ignore"):
            log_J = min(0.0, E0 - E1)                                     # :626
        u = tp.take("random")
        if log_J > (math.log(u) if u > 0 else -np.inf):                   # sampler.py:380
            if logL > logLmin:                                            # :382
                return dict(out=q, logL=logL, logP=logP, steps=steps, log_J=log_J, evals=2 * L * steps)
        if steps >= max_trajectories:
            return dict(out=q0, logL=start_logL, logP=start_logP, steps=steps, log_J=log_J, evals=2 * L * steps)


def hmc_production_chain(model, grad_logL, grad_V, flat_prior, start, start_logL, start_logP, C, dt, logLmin, trajectories,
                         nmcmc, max_traj):
    """The PRODUCTION HMC kernel's chain (cpnest_b200/csrc/hmc_kernel.cuh) on recorded inputs.  It is the reference's
    constrained leapfrog (position step with reflection off the bounds, proposal.py:729-741; momentum step :760-773; the two
    half trajectories of evolve_trajectory, :793-835; energies and log_J of get_sample, :605-627) with the stated differences:
    the reflection normal is the analytic gradient of log_likelihood (the reference fits splines, :417-486); the reflection is
    p -= 2 (n.C.p / n.C.n) n, in the metric of the kinetic energy (:769-771 uses the Euclidean one); the trajectory point is
    drawn in one pass by weighted reservoir sampling with weights exp(-H) over the points 1 .. 2L-1, summed in the linear domain
    relative to a reference energy (same distribution as the multinomial draw of sample_trajectory, :837-844); the kinetic energy is re-evaluated only after the momentum can have
    changed it (first point, a bound reflection, a non-flat prior); a chain runs trajectories until steps >= nmcmc and one was
    accepted (the MH rule, sampler.py:347; the reference stops at the first accepted one, :375).
    trajectories: list of dict(L, p0, u, u_acc) as cpn_trace_hmc records them.  Returns dict(out, logL, logP, steps, accepted, evals)."""
    lo, hi = model.lo, model.hi
    D = len(start)
    C = np.asarray(C, dtype=np.float64)
    imass = np.diag(C).copy()                                             # inverse_mass = diag(covariance), :519
    x = np.array(start, dtype=np.float64)
    logL, logP = float(start_logL), float(start_logP)
    steps = accepted = evals = reflections = 0
    neg_inf = -np.inf
    for tr in trajectories:
        L, p0, us, u_acc = int(tr["L"]), np.array(tr["p0"], dtype=np.float64), list(tr["u"]), float(tr["u_acc"])
        E0 = 0.5 * float(p0 @ (C @ p0)) - logP                            # hamiltonian(p0, q0), :621
        Wsum = Href = Hprev = wprev = 0.0                                 # running sum of the weights exp(Href - H)
        qc, logLc, logPc, Hc = x.copy(), neg_inf, neg_inf, 0.0
        point = 0
        for sign in (1.0, -1.0):
            pm = sign * p0
            q = x.copy()
            pm = pm + (-0.5 * dt) * grad_V(q)                             # half step, :757-759
            T, dirty = 0.0, True
            for i in range(L):
                point += 1
                flipped = False
                for j in range(D):                                        # :729-741
                    q[j] += dt * pm[j] * imass[j]
                    while q[j] < lo[j] or q[j] > hi[j]:
                        if q[j] > hi[j]:
                            q[j] = hi[j] - (q[j] - hi[j]); pm[j] *= -1; flipped = True
                        if q[j] < lo[j]:
                            q[j] = lo[j] + (lo[j] - q[j]); pm[j] *= -1; flipped = True
                lp = model.log_prior(q) if model.in_bounds(q) else neg_inf
                ll = model.log_likelihood(q)
                evals += 1
                gv = grad_V(q)
                if ll - logLmin > 0:                                      # :764-766
                    pm = pm + (-dt) * gv
                else:
                    n = np.array(grad_logL(q), dtype=np.float64)
                    n2 = float(n @ n)
                    n = n * (1.0 / math.sqrt(n2)) if n2 > 0 else n * 0.0
                    cn = C @ n
                    ncn = float(cn @ n)
                    pn = float(cn @ pm) / ncn if ncn > 0 else 0.0
                    pm = pm + (-2.0 * pn) * n
                    reflections += 1
                dirty = dirty or (not flat_prior) or flipped
                if dirty:
                    T = 0.5 * float(pm @ (C @ pm))
                    dirty = False
                H = T - lp
                if not (sign < 0 and i == L - 1):                         # trajectory[1:-1], :841
                    u = us[point - 1]
                    if H < np.inf:                                        # a point outside the box has weight 0
                        if not Wsum > 0.0:
                            Href, wgt = H, 1.0
                        elif H == Hprev:
                            wgt = wprev
                        else:
                            if H < Href - 50.0:                           # the reference energy follows H downwards
                                Wsum *= math.exp(H - Href)
                                Href = H
                            wgt = math.exp(Href - H)
                        Hprev, wprev = H, wgt
                        Wsum += wgt
                        if u * Wsum < wgt:                                # keep this point with probability w / (sum so far)
                            qc, logLc, logPc, Hc = q.copy(), ll, lp, H
        log_J = min(0.0, E0 - Hc)                                         # :626
        acc = (log_J > math.log(u_acc)) and (logLc > logLmin)             # sampler.py:380-382
        if acc:
            x, logL, logP = qc, logLc, logPc
        steps += 1
        accepted += 1 if acc else 0
        if (steps >= nmcmc and accepted > 0) or steps >= max_traj:
            break
    return dict(out=x, logL=logL, logP=logP, steps=steps, accepted=accepted, evals=evals, reflections=reflections)


def production_hmc_draw(seed, chain_id, attempt, D, base_L, eig_scale, eigen_vectors):
    """hmc_kernel.cuh, production branch: what trajectory number `attempt` of chain `chain_id` draws from the Philox stream:
    L = base_L + 10 + index(v0 of block 2, 40) (cf. proposal.py:553-559; chains sharing a warp use the first chain's),
    p0 = sum_i z_i |lambda_i|^-1/2 v_i with z_d = the first fp32 Box-Muller Gaussian of block 0x80000000 | d (momenta ~ N(0, M),
    M = C^-1, proposal.py:516-522, 620), u_acc = u(v0 of block 1); the uniform of point k of the weighted draw is word k & 3 of
    block 0x40000000 | (k >> 2).  Returns (L, p0, u_acc, uniform_of_point)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    clo, chi = chain_id & 0xFFFFFFFF, (chain_id >> 32) & 0xFFFFFFFF
    L = base_L + 10 + u32_to_index(philox4x32_10((clo, chi, attempt, 2), key)[0], 40)
    z = np.zeros(D)
    for d in range(D):
        q = philox4x32_10((clo, chi, attempt, 0x80000000 | d), key)
        z[d] = box_muller_f32(q[0], q[1])[0] / eig_scale[d] if eig_scale[d] > 0 else 0.0
    p0 = np.zeros(D)
    for i in range(D):
        p0 = p0 + z[i] * np.asarray(eigen_vectors)[:, i]
    u_acc = u32_to_unit_open(philox4x32_10((clo, chi, attempt, 1), key)[0])

    def uniform_of_point(k):
        return u32_to_unit_open(philox4x32_10((clo, chi, attempt, 0x40000000 | (k >> 2)), key)[k & 3])
    return L, p0, u_acc, uniform_of_point


def hmc_update_time_step(scale, acceptance, base_dt, target=0.5, adaptation=0.001):
    """proposal.py:539-550 -> (scale, dt)"""
    scale = float(np.exp(np.log(scale) + adaptation * (acceptance - target)))
    return scale, base_dt * scale


# ----------------------------------------------------------------------------------------
# Evidence integrator (cpnest/NestedSampling.py:88-144) and helpers (cpnest/nest2pos.py:17-42)
# ----------------------------------------------------------------------------------------
def logsubexp(x, y):
    """nest2pos.py:17-31"""
    return x + np.log1p(-np.exp(y - x))


def log_integrate_log_trap(log_func, log_support):
    """nest2pos.py:33-42"""
    log_func = np.asarray(log_func, dtype=np.float64)
    log_support = np.asarray(log_support, dtype=np.float64)
    with np.errstate(all="ignore"):
        log_func_sum = np.logaddexp(log_func[:-1], log_func[1:]) - np.log(2)
        log_dxs = logsubexp(log_support[:-1], log_support[1:])
        return float(np.logaddexp.reduce(log_func_sum + log_dxs))


class IntegralState:
    """_NSintegralState (NestedSampling.py:88-144)."""

    def __init__(self, nlive):
        self.nlive = nlive
        self.iteration = 0
        self.logZ = -np.inf
        self.logw = 0.0
        self.info = 0.0
        self.logLs = [-np.inf]
        self.log_vols = [0.0]

    def increment(self, logL, nlive=None, nreplace=1):
        if nlive is None:
            nlive = self.nlive
        oldZ = self.logZ
        logt = -nreplace / nlive                                         # :121
        with np.errstate(all="ignore"):
            Wt = self.logw + logL + logsubexp(0, logt)                   # :122
            self.logZ = float(np.logaddexp(self.logZ, Wt))               # :123
            if np.isfinite(oldZ) and np.isfinite(self.logZ) and np.isfinite(logL):   # :125
                self.info = float(np.exp(Wt - self.logZ) * logL
                                  + np.exp(oldZ - self.logZ) * (self.info + oldZ) - self.logZ)
                if math.isnan(self.info):
                    self.info = 0.0
        self.logw += logt                                                # :131
        self.iteration += 1
        self.logLs.append(logL)
        self.log_vols.append(self.logw)

    def finalise(self):
        """:136-144"""
        self.logZ = log_integrate_log_trap(np.array(self.logLs), np.array(self.log_vols))
        return self.logZ


def compute_weights(data, Nlive):
    """nest2pos.py:45-71"""
    data = np.asarray(data, dtype=np.float64)
    start_data = np.concatenate(([float("-inf")], data[:-Nlive]))
    end_data = data[-Nlive:]
    log_vols_start = np.cumsum(np.ones(len(start_data) + 1) * np.log1p(-1. / Nlive)) - np.log1p(-1. / Nlive)
    log_vols_end = np.zeros(len(end_data))
    log_vols_end[-1] = -np.inf
    log_vols_end[0] = log_vols_start[-1] + np.log1p(-1.0 / Nlive)
    for i in range(len(end_data) - 1):
        log_vols_end[i + 1] = log_vols_end[i] + np.log1p(-1.0 / (Nlive - i))
    log_likes = np.concatenate((start_data, end_data, [end_data[-1]]))
    log_vols = np.concatenate((log_vols_start, log_vols_end))
    log_ev = log_integrate_log_trap(log_likes, log_vols)
    with np.errstate(all="ignore"):
        log_dXs = logsubexp(log_vols[:-1], log_vols[1:])
    log_wts = log_likes[1:-1] + log_dXs[:-1]
    log_wts -= log_ev
    return log_ev, log_wts


def autocorrelation(x):
    """nest2pos.py:175-187"""
    x = np.asarray(x, dtype=np.float64)
    m = x.mean()
    v = np.var(x)
    xp = x - m
    cf = np.fft.fft(xp)
    sf = cf.conjugate() * cf
    return np.fft.ifft(sf).real / v / len(x)


def acl(x, tolerance=0.01):
    """nest2pos.py:189-199"""
    T = 1
    i = 0
    acf = autocorrelation(x)
    while i < len(acf) and acf[i] > tolerance:
        T += 2 * acf[i]
        i += 1
    return T


def estimate_nmcmc_on_the_fly(Nmcmc_exact, sub_acceptance, maxmcmc, poolsize, safety=5, tau=None):
    """sampler.py:156-176.  Returns (Nmcmc_exact, Nmcmc)."""
    if tau is None:
        tau = poolsize / safety
    if sub_acceptance == 0.0:
        Nmcmc_exact = (1.0 + 1.0 / tau) * Nmcmc_exact
    else:
        Nmcmc_exact = (1.0 - 1.0 / tau) * Nmcmc_exact + (safety / tau) * (2.0 / sub_acceptance - 1.0)
    Nmcmc_exact = float(min(Nmcmc_exact, maxmcmc))
    return Nmcmc_exact, max(safety, int(Nmcmc_exact))


# ----------------------------------------------------------------------------------------
# NestedSampler.consume_sample (cpnest/NestedSampling.py:318-375) on a sorted logL list
# ----------------------------------------------------------------------------------------
class NSLoopOracle:
    """The reference's batch step on the (sorted) live logL list.  `replacements` is the
    list of accepted new logL values in the order the reference receives them (k reversed).
    mode='reference': one integrator increment per batch, logt = -K/N at the K-th worst
    (NestedSampling.py:324-326).  mode='sequential': per-point shrinkage -1/(N-i) over the
    K sorted dead points, the rule the reference itself applies in its final sweep
    (NestedSampling.py:474-476); this is cpnest_b200's default (SURVEY.md hard part 5)."""

    def __init__(self, live_logL, mode="reference"):
        self.N = len(live_logL)
        self.keys = sorted(float(v) for v in live_logL)
        self.state = IntegralState(self.N)
        self.mode = mode
        self.iteration = 0
        self.insertion_indices = []
        self.nested_logL = []
        self.condition = np.inf

    def batch(self, K, replacements, logLmax):
        import bisect
        logLmin = self.keys[K - 1]                                       # :374
        if self.mode == "reference":
            self.state.increment(logLmin, nreplace=K)                    # :326
        else:
            for i in range(K):
                self.state.increment(self.keys[i], nlive=self.N - i)
        self.nested_logL.extend(self.keys[:K])                           # :327
        with np.errstate(all="ignore"):
            self.condition = float(np.logaddexp(self.state.logZ, logLmax - self.iteration / float(self.N))
                                   - self.state.logZ)                    # :332
        assert len(replacements) == K
        for j, k in enumerate(reversed(range(K))):
            self.iteration += 1
            new = float(replacements[j])
            assert new > logLmin
            index = bisect.bisect(self.keys, new)                        # KeyOrderedList.search :42-46
            self.keys.insert(index, new)
            self.insertion_indices.append((index - K) / (self.N - k - 1))   # :350
        del self.keys[:K]                                                # :366
        return logLmin

    def survivor_ranks(self, K, replacements):
        """cpnest_b200's insertion statistic: rank of each replacement among the N-K
        survivors only (bisect-right), normalised by N-K.  Coincides with the reference's
        (index - nreplace)/(Nlive - k - 1) when K == 1."""
        import bisect
        surv = self.keys[K:]
        return [bisect.bisect(surv, float(v)) / (self.N - K) for v in replacements]

    def finish(self):
        """NestedSampling.py:472-480"""
        for i, l in enumerate(self.keys):
            self.state.increment(l, nlive=self.N - i)
            self.nested_logL.append(l)
        rect = self.state.logZ
        return rect, self.state.info, self.state.finalise()


# ----------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011, "Parallel random numbers: as easy as 1, 2, 3"; the
# generator cpnest_b200's kernels use instead of the reference's unseeded MT19937) and the
# integer -> real transforms of cpnest_b200/csrc/rng.cuh, for known-answer tests.
# ----------------------------------------------------------------------------------------
PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    c = [int(v) & 0xFFFFFFFF for v in ctr]
    k = [int(v) & 0xFFFFFFFF for v in key]
    for _ in range(10):
        p0 = PHILOX_M0 * c[0]
        p1 = PHILOX_M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + PHILOX_W0) & 0xFFFFFFFF, (k[1] + PHILOX_W1) & 0xFFFFFFFF]
    return c


def u32_to_unit_open(r):
    """(r + 0.5) * 2^-32 in fp64: uniform on (0,1), never 0 or 1."""
    return (int(r) + 0.5) * (1.0 / 4294967296.0)


def u32_to_index(r, n):
    """mulhi(r, n): uniform index in [0, n)."""
    return (int(r) * int(n)) >> 32


# ----------------------------------------------------------------------------------------
# Step-level oracle of the PRODUCTION MH kernel (cpnest_b200/csrc/mh_kernel.cuh, the instantiation
# every run launches): the kernel records what each step consumed besides the chain state
# (cpn_trace_mh); the two functions below restate (a) how those inputs follow from the Philox
# stream and (b) the reference's chain given those inputs.
# ----------------------------------------------------------------------------------------
def box_muller_f32(r0, r1):
    """rng.cuh box_muller_f: two N(0,1) from two u32, fp32 arithmetic (the device uses the fast
    fp32 log / sincos units, so agreement is to ~1e-5 absolute, not bit-exact)."""
    f = np.float32
    u1 = (f(int(r0) >> 8) + f(0.5)) * f(1.0 / 16777216.0)
    th = (f(int(r1) >> 8) + f(0.5)) * f(2.0 / 16777216.0) - f(1.0)
    rad = np.sqrt(f(-2.0) * np.log(u1, dtype=np.float32), dtype=np.float32)
    ang = f(3.14159265358979) * th
    return float(rad * np.cos(ang, dtype=np.float32)), float(rad * np.sin(ang, dtype=np.float32))


def production_draw(seed, chain_id, step, kind, ens_n, D, eig_scale, flat_prior=True, self_slot=None):
    """mh_kernel.cuh DrawFeed::refill: the inputs of MCMC step `step` of chain `chain_id` as a function of the Philox
    stream (counter = (chain id lo, hi, step, block), key = seed).  Returns [kind, i0, i1, i2, c0, c1, c2, thr]:
      WALK    three distinct members (`random.sample(ensemble, 3)`, proposal.py:230), c_j = g_j - mean(g) in fp32
      STRETCH one member (`choice`, :253), c0 = Z = 2^t (fp32), t = uniform(-1,1); thr = log u - D t log 2 (:255-260)
      DE      two distinct members (:284), c0 = g = 1 + 1e-4 N(0,1) in fp32 (:285)
      EIGEN   i0 = randrange(D) (:338), c0 = sqrt|lambda_i0| N(0,1) in fp64 (:339)
    thr = log(u) - log_J of the prior MH test (sampler.py:337); -1 under a flat prior for the kinds with log_J = 0.
    self_slot: the live slot the chain started from; the members are then drawn from the other ens_n - 1 slots (the reference
    pops a chain's point from the pool before evolving it, sampler.py:325), compact index c -> slot c + (c >= self_slot)."""
    f = np.float32
    if self_slot is not None:
        ens_n -= 1
    unc = (lambda c: c + (c >= self_slot)) if self_slot is not None else (lambda c: c)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    clo, chi = chain_id & 0xFFFFFFFF, (chain_id >> 32) & 0xFFFFFFFF
    r = philox4x32_10((clo, chi, step, 0), key)
    thr = -1.0 if (flat_prior and kind != STRETCH) else math.log(u32_to_unit_open(r[3]))
    i0 = u32_to_index(r[0], D if kind == EIGEN else ens_n)
    i1 = i2 = i0
    c0 = c1 = c2 = 0.0
    if kind == WALK:
        i1 = u32_to_index(r[1], ens_n - 1)
        i1 += i1 >= i0
        i2 = u32_to_index(r[2], ens_n - 2)
        lo2, hi2 = min(i0, i1), max(i0, i1)
        i2 += i2 >= lo2
        i2 += i2 >= hi2
        i0, i1, i2 = unc(i0), unc(i1), unc(i2)
        q = philox4x32_10((clo, chi, step, 1), key)
        ga, gb = box_muller_f32(q[0], q[1])
        gc, _ = box_muller_f32(q[2], q[3])
        gbar = (f(ga) + f(gb) + f(gc)) * f(1.0 / 3.0)
        c0, c1, c2 = float(f(ga) - gbar), float(f(gb) - gbar), float(f(gc) - gbar)
    elif kind == STRETCH:
        i0 = unc(i0)                                        # (i1, i2 are not used by this kind)
        t = f(2.0 * u32_to_unit_open(r[1]) - 1.0)
        c0 = float(np.exp2(t, dtype=np.float32))
        thr -= D * (float(t) * 0.69314718055994530942)
    elif kind == DE:
        i1 = u32_to_index(r[1], ens_n - 1)
        i1 += i1 >= i0
        i0, i1 = unc(i0), unc(i1)
        ga, _ = box_muller_f32((r[2] << 16) & 0xFFFFFFFF, r[2] & 0xFFFF0000)
        c0 = float(f(1e-4) * f(ga) + f(1.0))
    else:
        ga, _ = box_muller_f32(r[1], r[2])
        c0 = float(eig_scale[i0]) * float(f(ga))
    return [float(kind), float(i0), float(i1), float(i2), c0, c1, c2, thr]


def production_slice_draw(seed, chain_id, s, kind, ens_n, D, mean=None, eig_scale=None, eigen_vectors=None):
    """slice_kernel.cuh, production branch: what slice number `s` of chain `chain_id` draws from the Philox stream before the
    shrinkage (counter = (chain id lo, hi, slice, block), key = seed).  Returns (uL, direction data, Exp(1), uJ):
      block 0: uL = u(v0) (reset_boundaries, sampler.py:444), uJ = u(v1) (:491), Exp(1) = -log u(v2) (:490),
               differential direction: first member = index(v3, n); block 1: second member = index(v0, n - 1), skipping the first
               (`sample(list(ensemble), 2)`, proposal.py:131)
      block 0x80000000 | d: one N(0,1) per coordinate d (fp32 Box-Muller of v0, v1) -> the vector np.random.normal returns
               (proposal.py:183), or z of mean + sum_i z_i sqrt|lambda_i| v_i = the multivariate_normal sample (:169-171)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    clo, chi = chain_id & 0xFFFFFFFF, (chain_id >> 32) & 0xFFFFFFFF
    r = philox4x32_10((clo, chi, s, 0), key)
    uL, uJ = u32_to_unit_open(r[0]), u32_to_unit_open(r[1])
    eexp = -math.log(u32_to_unit_open(r[2]))
    if kind == SLICE_DIFF:
        i0 = u32_to_index(r[3], ens_n)
        i1 = u32_to_index(philox4x32_10((clo, chi, s, 1), key)[0], ens_n - 1)
        i1 += i1 >= i0
        return uL, (i0, i1), eexp, uJ
    z = np.empty(D)
    for d in range(D):
        q = philox4x32_10((clo, chi, s, 0x80000000 | d), key)
        z[d] = box_muller_f32(q[0], q[1])[0]
    if kind == SLICE_GAUSS:
        return uL, z, eexp, uJ
    v = np.array(mean, dtype=np.float64)
    for i in range(D):
        v = v + (z[i] * eig_scale[i]) * np.asarray(eigen_vectors)[:, i]
    return uL, v, eexp, uJ


def production_slice_uniform(seed, chain_id, s, k):
    """The uniform behind the k-th shrinkage candidate of slice `s` (slice_kernel.cuh: word k & 3 of block
    0x40000000 | (k >> 2)); the candidate is X' = fma(R - L, u, L), np.random.uniform(L, R) of sampler.py:530."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10((chain_id & 0xFFFFFFFF, (chain_id >> 32) & 0xFFFFFFFF, s, 0x40000000 | (k >> 2)), key)
    return u32_to_unit_open(r[k & 3])


def mh_chain_from_draws(model, ens, eigen_vectors, start, start_logL, start_logP, draws, Nmcmc, maxmcmc, logLmin):
    """MetropolisHastingsSampler.yield_sample (sampler.py:322-358) on the per-step inputs `draws[s] = [kind, i0, i1, i2,
    c0, c1, c2, thr]` recorded by the production kernel.  The ensemble is the fixed live snapshot `ens` (the device does
    not pop/append, DESIGN.md "pool = live set").  Proposals in the reference's own operation order with the scalars as
    given (they already are fp32 values, the rounding LivePoint.__mul__ applies, parameter.pyx:96-120):
      STRETCH a + (old - a) Z (proposal.py:258), DE old + (b - a) g (:286), EIGEN old + jump v_i (:340-341);
      WALK old + sum_j c_j s_j, which is old + sum_j g_j (s_j - com) (:230-234) written with c_j = g_j - mean(g).
    Prior MH test `new_logP - old_logP + log_J > log u` (sampler.py:337) as `dlogP > thr`."""
    ens = np.asarray(ens, dtype=np.float64)
    V = np.asarray(eigen_vectors, dtype=np.float64)
    old = np.array(start, dtype=np.float64)
    old_logL, logp_old = float(start_logL), float(start_logP)
    steps = accepted = evals = 0
    for s in range(maxmcmc):
        kind, i0, i1, i2, c0, c1, c2, thr = draws[s]
        kind, i0, i1, i2 = int(kind), int(i0), int(i1), int(i2)
        steps += 1                                                       # :333
        if kind == WALK:
            assert len({i0, i1, i2}) == 3 and abs(c0 + c1 + c2) < 1e-6
            new = ((old + c0 * ens[i0]) + c1 * ens[i1]) + c2 * ens[i2]
        elif kind == STRETCH:
            new = ens[i0] + lp_mul(old - ens[i0], c0, compat=False)
        elif kind == DE:
            assert i0 != i1
            new = old + lp_mul(ens[i1] - ens[i0], c0, compat=False)
        else:
            new = old + c0 * V[:, i0]
        new_logP = model.log_prior(new)                                  # :335
        with np.errstate(all="ignore"):
            test = (new_logP - logp_old) > thr                           # :337
        if test:
            new_logL = model.log_likelihood(new)                         # :338
            evals += 1
            if new_logL > logLmin:                                       # :339
                old, old_logL, logp_old = new, new_logL, new_logP
                accepted += 1
        if (steps >= Nmcmc and accepted > 0) or steps >= maxmcmc:        # :347
            break
    return dict(out=old, logL=old_logL, logP=logp_old, steps=steps, accepted=accepted, evals=evals)
