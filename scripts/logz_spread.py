"""Spread of (logZ - analytic) / sqrt(H / nlive) over seeds on the bench workload (50-D Gaussian, nlive 1e4) as a function of
the batch size K and of the decorrelation target of the chain-length control.
usage: python scripts/logz_spread.py [seeds]"""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
from cpnest_b200.engine import Engine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
D, N = 50, 10000
for K, rho in ((4704, 0.1), (4704, 0.05), (2352, 0.1), (1176, 0.1), (1176, 0.05)):
    dev, nm, Hs = [], [], []
    for seed in range(S):
        e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=5000, seed=2000 + seed, rho_target=rho)
        e.init_prior()
        e.run(K, tolerance=0.1)
        nm.append(e.state()["nmcmc"])
        rect, trap = e.finalise()
        s = e.state()
        Hs.append(s["info"])
        dev.append((trap + D * math.log(20.0)) / math.sqrt(s["info"] / N))
        e.close()
    dev = np.array(dev)
    f = K / N
    infl = math.sqrt((1.0 / (1.0 - f) - 1.0) / -math.log(1.0 - f))
    print("K %5d rho_target %.2f: mean %+.2f +- %.2f, rms %.2f (batch-replacement variance alone predicts %.2f), max %.2f, nmcmc %d, H %.2f" % (
        K, rho, dev.mean(), dev.std(ddof=1) / math.sqrt(S), math.sqrt((dev ** 2).mean()), infl, np.abs(dev).max(), int(np.mean(nm)), np.mean(Hs)), flush=True)
