"""Wall time of the pieces of one complete run: python scripts/run_breakdown.py NLIVE K [functor dim]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4704
np.random.seed(1234)
e = Engine("gaussian_iid", [[-10, 10]] * 50, nlive=N, maxmcmc=5000, seed=1234)
e.set_cycle(DefaultProposalCycle().device_cycle())
for rep in range(2):
    e.set_seed(1234 + rep)
    t0 = time.perf_counter(); e.init_prior(); e.synchronize(); t1 = time.perf_counter()
    nb = e.run(K, tolerance=0.1); e.synchronize(); t2 = time.perf_counter()
    s = e.state()
    rect, trap = e.finalise(); e.synchronize(); t3 = time.perf_counter()
    print("N %d K %d rep %d: init %.4f s, run %.4f s (%d batches, %.3f ms/batch), finalise %.4f s, steps %.3e, logZ %.3f" % (
        N, K, rep, t1 - t0, t2 - t1, nb, 1e3 * (t2 - t1) / nb, t3 - t2, s["total_steps"], trap))
