import sys, math
sys.path.insert(0, ".")
import numpy as np
from scipy import integrate
from cpnest_b200.engine import Engine
f = lambda m, s: math.exp(-0.5 * m * m / (s * s)) / (s * math.sqrt(2 * math.pi)) / s
Z, _ = integrate.dblquad(f, 0.05, 1.0, lambda s: -5, lambda s: 5, epsabs=1e-10)
Z /= 10.0 * math.log(1.0 / 0.05)
for sampler in ("hmc", "mh", "slice"):
    for seed in (1, 2, 3):
        e = Engine("gaussian_mean_sigma", [[-5, 5], [0.05, 1]], nlive=1000, maxmcmc=100, seed=seed, sampler=sampler)
        e.init_prior(); e.run(64, tolerance=0.1); r, t = e.finalise(); s = e.state()
        sig = math.sqrt(s["info"] / 1000)
        print(sampler, seed, "logZ %.3f analytic %.3f dev %+.2f sigma  H %.2f nmcmc %d acc %.2f dt %.4g rho %.3f" % (t, math.log(Z), (t - math.log(Z)) / sig, s["info"], s["nmcmc"], s["total_accepted"] / s["total_steps"], s["hmc_dt"], s["rho_max"]))
        e.close()
