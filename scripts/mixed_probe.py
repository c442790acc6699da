"""A mixed sampler population on the 10-D Gaussian: what the three per-kind controllers settle on.
usage: python scripts/mixed_probe.py [dim] [nlive] [K]"""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from cpnest_b200.engine import Engine

D = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4704
for mix in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1), (2, 1, 1)):
    e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=2000, seed=1234, mix=mix)
    e.init_prior()
    e.synchronize()
    t0 = time.perf_counter()
    e.run(K, tolerance=0.1)
    e.synchronize()
    dt = time.perf_counter() - t0
    mid = e.state()
    rect, trap = e.finalise()
    s = e.state()
    sigma = math.sqrt(s["info"] / N)
    print("mix", mix, "chains", mid["chains_kind"], "nmcmc", mid["nmcmc_kind"], "rho", ["%.3f" % r for r in mid["rho_kind"]],
          "run %.3f s" % dt, "batches", mid["batches"], "evals %.3e" % s["total_evals"],
          "logZ %.4f (%.2f sigma)" % (trap, (trap + D * math.log(20.0)) / sigma), flush=True)
    e.close()
