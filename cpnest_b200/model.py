"""The user-model contract of cpnest (cpnest/model.py:11-159) plus the device-functor hook.

A model is still described by `names`, `bounds`, `log_likelihood(param)` and `log_prior(param)`.
Python callables cannot run inside a CUDA kernel and there is no CPU fallback, so a model also
names the device functor that evaluates it on the GPU:

    class MyModel(cpnest_b200.model.Model):
        names = ['x', 'y']; bounds = [[-10, 10], [-10, 10]]
        device_functor = ('gaussian_iid', {'mu': 0.0, 'sigma': 1.0})
        def log_likelihood(self, p): ...      # kept: it is the parity oracle of the functor

or carries its own functor as CUDA C++ source, compiled at run time with NVRTC into the same kernels:

    class MyModel(cpnest_b200.model.Model):
        names = ['a', 'b']; bounds = [[-5, 5], [-5, 5]]
        device_source = '''
        __device__ double cpn_user_log_likelihood(const double* x, int D, const double* params) {
            return -0.5 * (x[0] * x[0] + x[1] * x[1]) / (params[0] * params[0]);
        }'''
        device_params = [2.0]
        def log_likelihood(self, p): return -0.5 * (p['a']**2 + p['b']**2) / 4.0

`CPNest` checks the functor against the Python callables on random prior points before a run.
"""
from array import array

import numpy as np
from numpy import inf
from numpy.random import uniform

from .parameter import LivePoint


class Model(object):
    names = []     # parameter names, e.g. ['p1', 'p2']
    bounds = []    # prior box, e.g. [(min1, max1), (min2, max2)]
    device_functor = None   # (functor name, parameter dict), see cpnest_b200.engine.functor_blob
    device_source = None    # or: CUDA C++ source of the model's own functor (contract: csrc/user_model.cuh)
    device_params = ()      # doubles handed to the user functor as `params`

    def in_bounds(self, param):
        """Strict box test (cpnest/model.py:33)."""
        return all(self.bounds[i][0] < param.values[i] < self.bounds[i][1] for i in range(param.dimension))

    def new_point(self):
        """A point drawn uniformly in the box with finite log_prior (cpnest/model.py:35-53)."""
        logP = -inf
        while logP == -inf:
            p = LivePoint(self.names, d=array("d", [uniform(b[0], b[1]) for b in self.bounds]))
            logP = self.log_prior(p)
        return p

    def log_likelihood(self, param):
        raise NotImplementedError("Model.log_likelihood must be implemented by the user model")

    def log_prior(self, param):
        """Flat within the bounds (cpnest/model.py:66-82)."""
        return 0.0 if self.in_bounds(param) else -inf

    def potential(self, param):
        return -self.log_prior(param)

    def force(self, param):
        pass

    def strsample(self, sample):
        """One line of the chain file (cpnest/model.py:108-121): tab-joined '{:.20e}' values and the
        logL appended with '{:20e}' and no separator."""
        line = "\t".join("{0:.20e}".format(sample[n]) for n in sample.names)
        line += "{0:20e}".format(sample.logL)
        return line

    def header(self):
        return "\t".join(self.names) + "\tlogL"

    def from_normalised(self, normalised_value):
        d = array("d", [b[0] + normalised_value[i] * (b[1] - b[0]) for i, b in enumerate(self.bounds)])
        return LivePoint(self.names, d=d)

    def to_normalised(self, point):
        """Maps a point onto the unit cube.  (The reference's version references an undefined name,
        cpnest/model.py:158-159; this is the evident intent.)"""
        return array("d", [(v - b[0]) / (b[1] - b[0]) for v, b in zip(point.values, self.bounds)])

    # ---- device hook -------------------------------------------------------------------------
    def get_device_functor(self):
        f = self.device_functor
        if callable(f):
            f = f()
        if f is None and self.device_source:
            # the model's own log_likelihood / log_prior written in CUDA C++, compiled at run time (NVRTC)
            return "user", dict(source=self.device_source, params=list(self.device_params))
        if f is None:
            raise NotImplementedError(
                "%s has no device functor. cpnest_b200 evaluates log_prior/log_likelihood inside its CUDA kernels "
                "and has no CPU fallback: set `device_functor = (name, params)` to one of the built-in functors "
                "(gaussian_iid, gaussian_corr, eggbox, rosenbrock, ackley, ring, gaussian_mean_sigma, "
                "gaussian_mixture; ready-made models in cpnest_b200/models.py), or set `device_source` to CUDA C++ "
                "defining cpn_user_log_likelihood (contract in cpnest_b200/csrc/user_model.cuh)." % type(self).__name__)
        if isinstance(f, str):
            f = (f, {})
        return f[0], dict(f[1] or {})
