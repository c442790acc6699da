"""Short driver for ncu: a few nested-sampling batches of the bench workload (50-D Gaussian,
nlive=1e4, K=2500) after a warm-up that brings the run into its constrained regime."""
import sys
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4704
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
np.random.seed(1234)
e = Engine("gaussian_iid", [[-10, 10]] * 50, nlive=N, maxmcmc=5000, seed=1234, lanes=lanes)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
for _ in range(steps):
    e.ns_step(K)
e.synchronize()
s = e.state()
print("batches", s["batches"], "nmcmc", s["nmcmc"], "steps", s["total_steps"], "evals", s["total_evals"], "rho", s["rho_max"])
