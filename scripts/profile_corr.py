"""Short driver for ncu: the correlated 50-D Gaussian (D^2 likelihood), nlive 1e4, K = 4704, a few batches into the run."""
import sys
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
from oracle import cpnest_oracle as O          # (measurement tooling, not the product: builds the covariance of the model)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = O.make_correlated_gaussian(50)
np.random.seed(1234)
e = Engine("gaussian_corr", m.bounds, nlive=10000, maxmcmc=5000, seed=1234, params=m.params, lanes=lanes)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
for _ in range(steps):
    e.ns_step(4704)
e.synchronize()
s = e.state()
print("batches", s["batches"], "nmcmc", s["nmcmc"], "steps", s["total_steps"], "evals", s["total_evals"], "rho", s["rho_max"])
