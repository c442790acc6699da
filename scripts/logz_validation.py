"""Many seeds per sampler on problems with an analytic log-evidence: the deviations (logZ - analytic) / sqrt(H / nlive) should
look like N(0, 1) draws (BASELINE.json north_star: within 3 sigma).  One line per configuration.
usage: python scripts/logz_validation.py [seeds] [first seed] [only cases whose tag contains this]"""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from cpnest_b200.engine import Engine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
BASE = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ONLY = sys.argv[3] if len(sys.argv) > 3 else ""
CASES = [
    # tag, functor, D, nlive, K, maxmcmc, kwargs
    ("50-D Gaussian, MH", "gaussian_iid", 50, 10000, 4704, 5000, dict(sampler="mh")),
    ("10-D Gaussian, MH", "gaussian_iid", 10, 4000, 1000, 1000, dict(sampler="mh")),
    ("10-D Gaussian, slice", "gaussian_iid", 10, 4000, 1000, 1000, dict(sampler="slice")),
    ("10-D Gaussian, HMC", "gaussian_iid", 10, 4000, 1000, 200, dict(sampler="hmc")),
    ("10-D Gaussian, MH:slice:HMC = 2:1:1", "gaussian_iid", 10, 4000, 1000, 1000, dict(mix=(2, 1, 1))),
    ("2-D Gaussian, MH", "gaussian_iid", 2, 1000, 250, 1000, dict(sampler="mh")),
]
for tag, functor, D, N, K, maxmcmc, kw in CASES:
    if ONLY not in tag:
        continue
    dev, t0, evals = [], time.perf_counter(), 0
    for seed in range(S):
        e = Engine(functor, [[-10, 10]] * D, nlive=N, maxmcmc=maxmcmc, seed=BASE + seed, **kw)
        e.init_prior()
        e.run(K, tolerance=0.1)
        rect, trap = e.finalise()
        s = e.state()
        evals += s["total_evals"]
        dev.append((trap + D * math.log(20.0)) / math.sqrt(s["info"] / N))
        e.close()
    dev = np.array(dev)
    print("%-40s nlive %6d K %5d seeds %2d: mean %+.2f +- %.2f sigma, rms %.2f, max |dev| %.2f, %.2f s/run, %.2e evals/run" % (
        tag, N, K, S, dev.mean(), dev.std(ddof=1) / math.sqrt(S), math.sqrt((dev ** 2).mean()), np.abs(dev).max(),
        (time.perf_counter() - t0) / S, evals / S), flush=True)
