#!/usr/bin/env python
"""Complete runs of the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh) on its own test models:
log-evidence, information and a thinned set of posterior samples, stored as tests/golden/ref_runs.json.

BASELINE.json's north_star asks that end-to-end runs agree with "the reference's own run" (logZ within 3 sigma,
posterior means / variances through a KS / moment test).  The reference cannot travel to the GPU box, so its runs are
recorded here once and the GPU tests compare against the record (tests/test_gpu_vs_reference_runs.py).

Run:  bash oracle/build_ref.sh && python tests/golden/gen_ref_runs.py        (about two minutes on 8 cores)
"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import cpnest  # noqa: E402  (the reference)
import cpnest.model  # noqa: E402


class Gaussian2D(cpnest.model.Model):
    """tests/test_2d.py:5-16"""
    names = ["x", "y"]
    bounds = [[-10, 10], [-10, 10]]

    def log_likelihood(self, p):
        return -0.5 * (p["x"] ** 2 + p["y"] ** 2) - np.log(2.0 * np.pi)


class Ring(cpnest.model.Model):
    """tests/test_ring.py:5-23"""
    names = ["x", "y"]
    bounds = [[-2, 2], [-2, 2]]

    def log_likelihood(self, p):
        r = np.sqrt(p["x"] ** 2 + p["y"] ** 2)
        return -0.5 * (1.0 - r) ** 2 / 0.1 ** 2


class GaussianMeanSigma(cpnest.model.Model):
    """tests/test_gaussian.py:5-20 (1/sigma prior)"""
    names = ["mean", "sigma"]
    bounds = [[-5, 5], [0.05, 1]]

    def log_likelihood(self, x):
        return -0.5 * x["mean"] ** 2 / x["sigma"] ** 2 - np.log(x["sigma"]) - 0.5 * np.log(2.0 * np.pi)

    def log_prior(self, p):
        if not self.in_bounds(p):
            return -np.inf
        return -np.log(p["sigma"]) - np.log(10) - np.log(0.95)


class GaussianND(cpnest.model.Model):
    """examples/test_50d.py:5-28 (iid unit Gaussian in a [-10, 10]^D box; no force: the HMC runs use the flat-prior force)"""

    def __init__(self, dim):
        self.names = ["{0}".format(i) for i in range(dim)]
        self.bounds = [[-10, 10] for _ in range(dim)]

    def log_likelihood(self, p):
        return np.sum([-0.5 * p[n] ** 2 - 0.5 * np.log(2.0 * np.pi) for n in p.names])

    def force(self, p):
        return np.zeros(1, dtype={"names": p.names, "formats": ["f8" for _ in p.names]})


class Gaussian10D(GaussianND):
    def __init__(self):
        GaussianND.__init__(self, 10)


class Gaussian50D(GaussianND):
    def __init__(self):
        GaussianND.__init__(self, 50)


class Gaussian2DForce(Gaussian2D):
    """tests/test_2d_hmc.py:5-24"""

    def force(self, p):
        return np.zeros(1, dtype={"names": p.names, "formats": ["f8" for _ in p.names]})


RUNS = [
    # name, model, CPNest keywords (the reference's own tests: tests/test_2d.py:30, tests/test_ring.py:31,
    # tests/test_gaussian.py:31, tests/test_2d_slice.py:30 with smaller maxmcmc to keep the recipe short)
    ("gaussian2d_mh", Gaussian2D, dict(nlive=1000, nthreads=8, maxmcmc=1000, poolsize=1000, seed=1234)),
    ("ring_mh", Ring, dict(nlive=1000, nthreads=8, maxmcmc=1000, poolsize=1000, seed=1234)),
    ("gaussian_mean_sigma_mh", GaussianMeanSigma, dict(nlive=1000, nthreads=8, maxmcmc=500, poolsize=1000, seed=1234)),
    ("gaussian2d_slice", Gaussian2D, dict(nlive=1000, nthreads=8, nslice=8, maxmcmc=1000, poolsize=1000, seed=4321)),
    # round 2: beyond two dimensions, and the third sampler (tests/test_2d_hmc.py:34; examples/test_50d.py:37 with MH)
    ("gaussian10d_mh", Gaussian10D, dict(nlive=1000, nthreads=8, maxmcmc=1000, poolsize=1000, seed=1234)),
    ("gaussian2d_hmc", Gaussian2DForce, dict(nlive=1000, nthreads=8, nhamiltonian=8, maxmcmc=200, poolsize=1000, seed=1234)),
    ("gaussian50d_mh", Gaussian50D, dict(nlive=1000, nthreads=8, maxmcmc=100, poolsize=1000, seed=1234)),
]


def main():
    # python gen_ref_runs.py [name ...]: only these runs are (re)made, the others are kept from the existing file (the
    # reference draws from the unseeded stdlib `random`: a rerun gives another realisation)
    only = set(sys.argv[1:])
    path = os.path.join(HERE, "ref_runs.json")
    old = {r["name"]: r for r in json.load(open(path))} if (only and os.path.exists(path)) else {}
    out = []
    for name, cls, kw in RUNS:
        if only and name not in only:
            if name in old:
                out.append(old[name])
            continue
        tmp = tempfile.mkdtemp(prefix="cpnest_ref_run_")
        t0 = time.time()
        work = cpnest.CPNest(cls(), verbose=0, output=tmp, **kw)
        work.run()
        dt = time.time() - t0
        pos = work.posterior_samples
        names = list(work.user.names)
        rs = np.random.RandomState(0)
        keep = np.sort(rs.permutation(len(pos))[:1500])
        rec = dict(name=name, kwargs=kw, names=names, bounds=work.user.bounds, logZ=float(work.NS.logZ), info=float(work.NS.state.info),
                   n_posterior=int(len(pos)), n_nested=int(len(work.nested_samples)), wall_s=round(dt, 2),
                   posterior={n: [float(v) for v in np.asarray(pos[n]).ravel()[keep]] for n in names},
                   posterior_mean={n: float(np.mean(pos[n])) for n in names},
                   posterior_var={n: float(np.var(pos[n])) for n in names})
        print(name, "logZ %.4f H %.3f posterior %d samples, %.1f s" % (rec["logZ"], rec["info"], rec["n_posterior"], dt), flush=True)
        out.append(rec)
        shutil.rmtree(tmp, ignore_errors=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
