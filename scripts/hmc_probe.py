"""HMC kernel diagnostics: a few batches at a time, printing the controller state."""
import sys, time, math
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
D = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 128
maxmcmc = int(sys.argv[4]) if len(sys.argv) > 4 else 100
nb = int(sys.argv[5]) if len(sys.argv) > 5 else 40
e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=maxmcmc, seed=1234, sampler="hmc")
e.init_prior()
prev = e.state()
for b in range(nb):
    t0 = time.time()
    e.ns_step(K); e.synchronize()
    dt = time.time() - t0
    s = e.state()
    tr = s["total_steps"] - prev["total_steps"]; ac = s["total_accepted"] - prev["total_accepted"]; ev = s["total_evals"] - prev["total_evals"]
    if b % 4 == 0 or b == nb - 1:
        print("batch %3d %.3fs nmcmc %4d dt %.4f traj/chain %.2f acc %.3f evals/traj %.0f failed %d rho %.3f logLmin %.2f" % (
            b, dt, s["nmcmc"], s["hmc_dt"], tr / K, ac / max(1, tr), ev / max(1, tr), s["failed_chains"], s["rho_max"], s["logLmin"]), flush=True)
    prev = s
