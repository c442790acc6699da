"""Jacobi sweeps of the asynchronous eigen-decomposition refreshes, warm start (previous eigenvectors) vs cold (identity):
python scripts/eigen_sweeps.py [functor: gaussian_iid | gaussian_corr]   (CPN_DEBUG prints the count at every ensemble_stats)"""
import sys, os
import numpy as np
sys.path.insert(0, ".")
os.environ["CPN_DEBUG"] = "1"
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
from oracle import cpnest_oracle as O          # (measurement tooling: builds the correlated model's parameters)
functor = sys.argv[1] if len(sys.argv) > 1 else "gaussian_iid"
params = O.make_correlated_gaussian(50).params if functor == "gaussian_corr" else None
for cold in (False, True):
    e = Engine(functor, [[-10, 10]] * 50, nlive=10000, maxmcmc=5000, seed=1234, cold_eigen=cold, params=params)
    e.set_cycle(DefaultProposalCycle().device_cycle())
    e.init_prior()
    print("cold" if cold else "warm", flush=True)
    for b in range(60):
        e.ns_step(4704)
        if b % 6 == 5:
            e.ensemble_stats()
    e.close()
