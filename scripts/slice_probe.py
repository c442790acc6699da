"""Slice sampler on the 10-D configs (BASELINE config 3): throughput and validity probe."""
import sys, time, math
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
which = sys.argv[1] if len(sys.argv) > 1 else "gaussian_iid"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
K = int(sys.argv[4]) if len(sys.argv) > 4 else 2500
sampler = sys.argv[5] if len(sys.argv) > 5 else "slice"
lanes = int(sys.argv[6]) if len(sys.argv) > 6 else 0
seed = int(sys.argv[7]) if len(sys.argv) > 7 else 1234
rho = float(sys.argv[8]) if len(sys.argv) > 8 else 0.0
maxmcmc = int(sys.argv[9]) if len(sys.argv) > 9 else 5000
b = 10.0 if which == "gaussian_iid" else 5.0
e = Engine(which, [[-b, b]] * D, nlive=N, maxmcmc=maxmcmc, seed=seed, sampler=sampler, lanes=lanes, rho_target=rho)
e.init_prior()
e.synchronize()
t0 = time.time()
nb = e.run(K, tolerance=0.1)
e.synchronize()
dt = time.time() - t0
s = e.state()
rect, trap = e.finalise()
sf = e.state()
ana = -D * math.log(2 * b) if which == "gaussian_iid" else float("nan")
print("%s D=%d N=%d K=%d %s: %.2fs batches %d nmcmc %d mu %.3g evals %.3e (%.3e/s) steps %.3e logZ %.3f (analytic %.3f, sigma %.3f) H %.2f rho %.3f dt %.3g failed %d" % (
    which, D, N, K, sampler, dt, nb, s["nmcmc"], s["slice_mu"], s["total_evals"], s["total_evals"] / dt, s["total_steps"], trap, ana,
    math.sqrt(sf["info"] / N), sf["info"], s["rho_max"], s["hmc_dt"], s["failed_chains"]))
