"""A model whose log_likelihood / log_prior is written by the user (CUDA C++ source on the Model subclass), compiled
at run time with NVRTC into the same chain kernels as the built-in functors: the device form of subclassing
cpnest.model.Model (cpnest/model.py:55-82).  The Python callables of the model remain the parity oracle."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SRC_GAUSS = """
__device__ double cpn_user_log_likelihood(const double* x, int D, const double* params) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) { const double t = (x[d] - params[0]) / params[1]; s += -0.5 * t * t - log(params[1]) - 0.91893853320467274178; }
    return s;
}
"""

SRC_MEAN_SIGMA = """
#define CPN_USER_HAS_PRIOR
__device__ double cpn_user_log_likelihood(const double* x, int D, const double* params) {
    return -0.5 * x[0] * x[0] / (x[1] * x[1]) - log(x[1]) - 0.91893853320467274178;
}
__device__ double cpn_user_log_prior(const double* x, int D, const double* params) {
    return -log(x[1]) - log(10.0) - log(0.95);
}
"""


def test_user_functor_matches_builtin_and_python():
    from cpnest_b200.engine import Engine
    rs = np.random.RandomState(0)
    for D in (2, 10, 50):
        with Engine("user", [[-10, 10]] * D, nlive=64, params=dict(source=SRC_GAUSS, params=[0.3, 1.7])) as e, \
                Engine("gaussian_iid", [[-10, 10]] * D, nlive=64, params=dict(mu=0.3, sigma=1.7)) as b:
            X = rs.uniform(-9, 9, size=(100, D))
            lu, pu = e.eval_model(X)
            lb, pb = b.eval_model(X)
            ref = np.array([np.sum(-0.5 * ((x - 0.3) / 1.7) ** 2 - math.log(1.7) - 0.5 * math.log(2 * math.pi)) for x in X])
            assert np.allclose(lu, ref, rtol=1e-12, atol=0) and np.allclose(lu, lb, rtol=1e-12, atol=0)
            assert np.all(pu == 0) and np.all(pb == 0)
            gl, gv = e.eval_gradients(X[:8])                 # finite differences of the user function
            assert np.allclose(gl, -(X[:8] - 0.3) / 1.7 ** 2, rtol=1e-5, atol=1e-6) and np.all(gv == 0)


def test_compile_error_is_reported():
    from cpnest_b200.engine import Engine
    from cpnest_b200._lib import CpnError
    with pytest.raises(CpnError, match="compiling the user functor failed"):
        Engine("user", [[-1, 1]] * 2, nlive=16, params=dict(source="__device__ double cpn_user_log_likelihood(const double* x, int D, const double* p) { return oops; }"))


@pytest.mark.parametrize("sampler", ["mh", "slice", "hmc"])
def test_user_functor_end_to_end_logz(sampler):
    from cpnest_b200.engine import Engine
    with Engine("user", [[-10, 10]] * 4, nlive=1000, maxmcmc=200, seed=5, sampler=sampler,
                params=dict(source=SRC_GAUSS, params=[0.0, 1.0])) as e:
        e.init_prior()
        e.run(64, tolerance=0.1)
        rect, trap = e.finalise()
        s = e.state()
        assert abs(trap + 4 * math.log(20.0)) < 3 * math.sqrt(s["info"] / 1000), (trap, s["info"])


def test_user_model_with_prior_through_the_facade(tmp_path):
    """tests/test_gaussian.py's (mean, sigma) model, its likelihood and 1/sigma prior supplied as CUDA source."""
    import cpnest_b200 as cpnest
    import cpnest_b200.model
    from scipy import integrate

    class GaussianModel(cpnest.model.Model):
        names = ["mean", "sigma"]
        bounds = [[-5, 5], [0.05, 1]]
        device_source = SRC_MEAN_SIGMA

        def log_likelihood(self, p):
            return -0.5 * p["mean"] ** 2 / p["sigma"] ** 2 - np.log(p["sigma"]) - 0.5 * np.log(2.0 * np.pi)

        def log_prior(self, p):
            if not self.in_bounds(p):
                return -np.inf
            return -np.log(p["sigma"]) - np.log(10) - np.log(0.95)

    work = cpnest.CPNest(GaussianModel(), verbose=0, nthreads=8, nlive=1000, maxmcmc=200, output=str(tmp_path))
    work.run()
    f = lambda m, s: math.exp(-0.5 * m * m / (s * s)) / (s * math.sqrt(2 * math.pi)) / s
    Z, _ = integrate.dblquad(f, 0.05, 1.0, lambda s: -5, lambda s: 5, epsabs=1e-10)
    Z /= 10.0 * math.log(1.0 / 0.05)
    assert abs(work.NS.logZ - math.log(Z)) < 3.0 * np.sqrt(work.NS.state.info / work.NS.Nlive)
