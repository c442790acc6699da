"""Short driver for ncu: a few batches of a sampler on a mid-size problem.
usage: python scripts/profile_sampler.py <mh|slice|hmc> [functor] [D] [N] [K] [batches]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
sampler = sys.argv[1]
functor = sys.argv[2] if len(sys.argv) > 2 else "rosenbrock"
D = int(sys.argv[3]) if len(sys.argv) > 3 else 10
N = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
K = int(sys.argv[5]) if len(sys.argv) > 5 else 2500
nb = int(sys.argv[6]) if len(sys.argv) > 6 else 30
b = 10.0 if functor == "gaussian_iid" else 5.0
e = Engine(functor, [[-b, b]] * D, nlive=N, maxmcmc=1000, seed=1234, sampler=sampler)
e.init_prior()
for _ in range(nb):
    e.ns_step(K)
e.synchronize()
s = e.state()
print(sampler, functor, D, "batches", s["batches"], "nmcmc", s["nmcmc"], "steps", s["total_steps"], "evals", s["total_evals"], "mu", s["slice_mu"], "dt", s["hmc_dt"])
