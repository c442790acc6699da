import sys, numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
np.random.seed(1234)
e = Engine("gaussian_iid", [[-10, 10]] * 50, nlive=10000, maxmcmc=5000, seed=1234)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
e.ensemble_stats()
for i in range(6):
    e.ns_step(2500); e.ensemble_stats()
for i in range(200):
    e.ns_step(2500)
for i in range(3):
    e.ns_step(2500); m,c,w,v = e.ensemble_stats()
print(w[:3], w[-3:])
