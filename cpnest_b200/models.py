"""Ready-made models: the reference's test/example models with their device functors.
The Python `log_likelihood` of each class restates the reference's callable (cited) and serves as
the host-side parity check of the functor; it is never used for sampling."""
import math

import numpy as np

from .model import Model


class GaussianModel(Model):
    """n-dimensional unit Gaussian: examples/test_50d.py:5-19, tests/test_2d.py:5-16 (dim=2)."""

    def __init__(self, dim=50, bound=10.0):
        self.dim = dim
        self.names = ["x", "y"] if dim == 2 else ["{0}".format(i) for i in range(dim)]
        self.bounds = [[-bound, bound] for _ in range(dim)]
        self.analytic_log_Z = -sum(math.log(b[1] - b[0]) for b in self.bounds)
        self.device_functor = ("gaussian_iid", {"mu": 0.0, "sigma": 1.0})

    def log_likelihood(self, p):
        return float(np.sum([-0.5 * v * v - 0.5 * math.log(2.0 * math.pi) for v in p.values]))


class HalfGaussianModel(GaussianModel):
    """tests/test_half_gaussian.py:11-25"""

    def __init__(self, xmax=10.0):
        from scipy import stats
        GaussianModel.__init__(self, dim=1)
        self.names = ["x"]
        self.bounds = [[0, xmax]]
        self.analytic_log_Z = math.log(stats.norm.cdf(xmax) - stats.norm.cdf(0)) - math.log(xmax)


class EggboxModel(Model):
    """examples/eggbox.py:5-23"""
    names = ["1", "2"]
    bounds = [[0, 10.0 * math.pi], [0, 10.0 * math.pi]]
    device_functor = ("eggbox", {})
    analytic_log_Z = 235.85594     # quadrature, SURVEY.md 8c

    def log_likelihood(self, x):
        tmp = 1.0
        for v in x.values:
            tmp *= math.cos(v / 2.0)
        return (tmp + 2.0) ** 5.0


class RosenbrockModel(Model):
    """examples/rosenbrock.py:7-22"""
    device_functor = ("rosenbrock", {})

    def __init__(self, ndims=2):
        self.ndims = ndims
        self.names = ["x" + str(i) for i in range(ndims)]
        self.bounds = [[-5, 5] for _ in range(ndims)]

    def log_likelihood(self, x):
        v = np.array(x.values)
        return float(-np.sum(100.0 * (v[1:] - v[:-1] ** 2.0) ** 2.0 + (1.0 - v[:-1]) ** 2.0))


class AckleyModel(Model):
    """Standard D-dimensional Ackley function, logL = -ackley(x).  examples/ackley.py:21-28 takes
    log() of a non-positive number and is finite nowhere; SURVEY.md 8d defines the config this way."""
    device_functor = ("ackley", {})

    def __init__(self, ndims=2):
        self.names = ["x" + str(i) for i in range(ndims)]
        self.bounds = [[-5, 5] for _ in range(ndims)]

    def log_likelihood(self, x):
        v = np.array(x.values)
        D = len(v)
        return -(-20.0 * math.exp(-0.2 * math.sqrt(float(np.sum(v * v)) / D))
                 - math.exp(float(np.sum(np.cos(2.0 * math.pi * v))) / D) + 20.0 + math.e)


class RingModel(Model):
    """tests/test_ring.py:5-23"""
    names = ["x", "y"]
    bounds = [[-2, 2], [-2, 2]]
    analytic_log_Z = -2.31836
    device_functor = ("ring", {"radius": 1.0, "thickness": 0.1})

    def log_likelihood(self, p):
        r = math.sqrt(p["x"] ** 2 + p["y"] ** 2)
        return -0.5 * (1.0 - r) ** 2 / 0.1 ** 2


class GaussianMeanSigmaModel(Model):
    """tests/test_gaussian.py:5-20: parameters (mean, sigma), 1/sigma prior."""
    names = ["mean", "sigma"]
    bounds = [[-5, 5], [0.05, 1]]
    device_functor = ("gaussian_mean_sigma", {})

    def log_likelihood(self, x):
        return -0.5 * x["mean"] ** 2 / x["sigma"] ** 2 - math.log(x["sigma"]) - 0.5 * math.log(2.0 * math.pi)

    def log_prior(self, p):
        if not self.in_bounds(p):
            return -np.inf
        return -math.log(p["sigma"]) - math.log(10) - math.log(0.95)


class GaussianMixtureModel(Model):
    """examples/gaussianmixture.py:5-31 with the data drawn from a seeded generator (the reference
    draws them unseeded)."""

    def __init__(self, seed=1234, ndata=10000):
        self.names = ["mean1", "sigma1", "mean2", "sigma2", "weight"]
        self.bounds = [[-3, 3], [0.01, 1], [-3, 3], [0.01, 1], [0.0, 1.0]]
        rs = np.random.RandomState(seed)
        data = []
        for _ in range(ndata):
            if rs.uniform(0.0, 1.0) < 0.1:
                data.append(rs.normal(0.5, 0.5))
            else:
                data.append(rs.normal(-1.5, 0.03))
        self.data = np.array(data)
        self.device_functor = ("gaussian_mixture", {"data": self.data})

    def log_likelihood(self, x):
        w = x["weight"]
        l1 = np.log(w) - np.log(x["sigma1"]) - 0.5 * ((self.data - x["mean1"]) / x["sigma1"]) ** 2
        l2 = np.log(1.0 - w) - np.log(x["sigma2"]) - 0.5 * ((self.data - x["mean2"]) / x["sigma2"]) ** 2
        return float(np.logaddexp(l1, l2).sum())

    def log_prior(self, p):
        if not self.in_bounds(p):
            return -np.inf
        return -math.log(p["sigma1"]) - math.log(p["sigma2"])


class CorrelatedGaussianModel(Model):
    """N(0, Sigma) with Sigma = Q diag(s^2) Q^T, s log-spaced in [0.1, 1] (the
    'correlated' flavour; not in the reference).  logZ = -D log(2 bound)."""

    def __init__(self, dim=50, seed=1234, bound=10.0):
        rs = np.random.RandomState(seed)
        Q, _ = np.linalg.qr(rs.normal(size=(dim, dim)))
        ev = np.exp(np.linspace(math.log(0.1), math.log(1.0), dim)) ** 2
        self.Sigma = (Q * ev) @ Q.T
        Lc = np.linalg.cholesky(self.Sigma)
        self.Linv = np.tril(np.linalg.inv(Lc))
        self.logdet = float(np.sum(np.log(np.diag(Lc))))
        self.names = [str(i) for i in range(dim)]
        self.bounds = [[-bound, bound] for _ in range(dim)]
        self.analytic_log_Z = -dim * math.log(2 * bound)
        self.device_functor = ("gaussian_corr", {"Linv": self.Linv, "logdet": self.logdet})

    def log_likelihood(self, p):
        y = self.Linv @ np.array(p.values)
        return -0.5 * float(y @ y) - self.logdet - 0.5 * len(self.names) * math.log(2.0 * math.pi)
