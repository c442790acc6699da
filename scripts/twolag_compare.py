"""Two-lag chain-length control on / off on the bimodal 20-D mixture and on the 5-D data mixture: steps, final chain length, logZ."""
import sys, math, time
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
mix = dict(w=0.3, m1=-2.5, s1=1.0, m2=2.5, s2=1.0)
rs = np.random.RandomState(1234)
n = 10000
data = np.concatenate((rs.normal(0.5, 0.5, size=n // 10), rs.normal(-1.5, 0.03, size=n - n // 10)))
CASES = [("mix20 nlive 4000", "gaussian_mixture_nd", [[-10, 10]] * 20, 4000, 1000, 2000, mix, -20 * math.log(20.0)),
         ("data mixture 5-D", "gaussian_mixture", [[-3, 3], [0.01, 1], [-3, 3], [0.01, 1], [0.0, 1.0]], 2000, 592, 500, dict(data=data), None)]
for tag, f, b, N, K, mx, prm, anchor in CASES:
    for two in (False, True):
        for seed in (31, 32, 33):
            np.random.seed(seed)
            e = Engine(f, b, nlive=N, maxmcmc=mx, seed=seed, params=prm, two_lag=two)
            e.set_cycle(DefaultProposalCycle().device_cycle())
            t0 = time.perf_counter(); e.init_prior(); e.run(K, tolerance=0.1); s = e.state(); rect, trap = e.finalise(); dt = time.perf_counter() - t0
            sf = e.state(); sig = math.sqrt(sf["info"] / N)
            print("%-18s two_lag %-5s seed %d: %.3f s, steps %.3e, nmcmc_final %4d, rho %.2f decay %.3f, logZ %.3f%s" % (
                tag, two, seed, dt, s["total_steps"], s["nmcmc"], s["rho_max"], s["rho_decay_kind"][0], trap,
                "" if anchor is None else " (%+.2f sigma)" % ((trap - anchor) / sig)), flush=True)
            e.close()
