"""The likelihood-bound workload: the reference's Gaussian-mixture example (examples/gaussianmixture.py: 5 parameters,
1e4 data per likelihood call) with MH chains.  Short driver for ncu and for a throughput figure.
usage: python scripts/profile_mixture.py [nlive] [K] [batches] [maxmcmc]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from cpnest_b200.engine import Engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 592
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 12
maxmcmc = int(sys.argv[4]) if len(sys.argv) > 4 else 200
rs = np.random.RandomState(1234)
n = 10000
data = np.concatenate((rs.normal(0.5, 0.5, size=n // 10), rs.normal(-1.5, 0.03, size=n - n // 10)))
b = [[-3, 3], [0.01, 1], [-3, 3], [0.01, 1], [0.0, 1.0]]
e = Engine("gaussian_mixture", b, nlive=N, maxmcmc=maxmcmc, seed=1234, params=dict(data=data), sampler="mh")
e.init_prior()
e.synchronize()
s0 = e.state()
t0 = time.perf_counter()
for _ in range(nb):
    e.ns_step(K)
e.synchronize()
dt = time.perf_counter() - t0
s = e.state()
ev = s["total_evals"] - s0["total_evals"]
# per data term: 16 flops + 1 exp (+ 1 log1p in the reference's form)
print("mixture mh: batches", s["batches"], "nmcmc", s["nmcmc"], "evals", ev, "evals/s %.4g" % (ev / dt),
      "data terms/s %.4g" % (ev * n / dt), "algorithmic TFLOP/s (16 flops/term) %.3f" % (ev * n * 16 / dt / 1e12),
      "rho", s["rho_max"], "logLmin", s["logLmin"])
