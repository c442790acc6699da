"""Trace of the chain-length controllers on the bimodal 20-D mixture: python scripts/twolag_probe.py"""
import sys, math
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
mix = dict(w=0.3, m1=-2.5, s1=1.0, m2=2.5, s2=1.0)
np.random.seed(1)
e = Engine("gaussian_mixture_nd", [[-10, 10]] * 20, nlive=4000, maxmcmc=2000, seed=31, params=mix)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
for b in range(200):
    e.ns_step(1000)
    s = e.state()
    if b % 5 == 0 or s["done"]:
        print(b, "nmcmc", s["nmcmc"], "rho %.3f" % s["rho_max"], "twolag", s["two_lag_kind"][0], "decay %.3f" % s["rho_decay_kind"][0], "cond %.2f" % s["condition"], flush=True)
    if s["done"] or not (s["condition"] > 0.1):
        break
