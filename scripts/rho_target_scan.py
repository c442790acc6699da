"""Bias of logZ against the decorrelation target of the chain-length control: many seeds of the 10-D Gaussian (MH, HMC) and
of the 50-D Gaussian (MH).  usage: python scripts/rho_target_scan.py [seeds]"""
import math, sys, time
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 48
CASES = [("10-D MH", 10, 4000, 1000, 1000, "mh"), ("10-D HMC", 10, 4000, 1000, 200, "hmc"), ("50-D MH", 50, 10000, 4704, 5000, "mh")]
for tag, D, N, K, maxmcmc, smp in CASES:
    for rt in (0.1, 0.07, 0.05):
        S_ = S if D < 50 else max(8, S // 3)
        dev, nm, t0, st = [], [], time.perf_counter(), 0
        for seed in range(S_):
            e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=maxmcmc, seed=3000 + seed, sampler=smp, rho_target=rt)
            e.init_prior()
            e.run(K, tolerance=0.1)
            s0 = e.state()
            rect, trap = e.finalise()
            s = e.state()
            dev.append((trap + D * math.log(20.0)) / math.sqrt(s["info"] / N))
            nm.append(s0["nmcmc"]); st += s["total_steps"]
            e.close()
        dev = np.array(dev)
        print("%-9s rho_target %.2f seeds %3d: mean %+.3f +- %.3f sigma, rms %.2f, nmcmc %5.0f, steps/run %.3e, %.3f s/run" % (
            tag, rt, S_, dev.mean(), dev.std(ddof=1) / math.sqrt(S_), math.sqrt((dev ** 2).mean()), np.mean(nm), st / S_,
            (time.perf_counter() - t0) / S_), flush=True)
