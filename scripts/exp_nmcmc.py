"""Experiment: logZ bias of the D-dim Gaussian against the chain length (fixed, or adapted to a
start/end correlation target)."""
import math, sys, time
import numpy as np
from scipy import stats
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from oracle import cpnest_oracle as O

D = int(sys.argv[1]) if len(sys.argv) > 1 else 50
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
mode = sys.argv[4] if len(sys.argv) > 4 else "fixed"
vals = [float(v) for v in (sys.argv[5].split(",") if len(sys.argv) > 5 else "50,100,200,500,1000,2000".split(","))]
lanes = int(sys.argv[6]) if len(sys.argv) > 6 else 0
ana = -D * math.log(20.0)
for v in vals:
    if mode == "fixed":
        e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=max(int(v), 10), seed=11, adapt_nmcmc=False, nmcmc_init=int(v), lanes=lanes)
    else:
        e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=5000, seed=11, adapt_nmcmc=True, rho_target=v, lanes=lanes)
    e.set_cycle(O.make_cycle(np.random.RandomState(11)))
    t0 = time.time(); e.init_prior(); t1 = time.time()
    nb = e.run(K); e.synchronize(); t2 = time.time()
    ins = e.insertion_indices()
    s0 = e.state()
    rect, trap = e.finalise()
    s = e.state()
    sig = math.sqrt(s["info"] / N)
    print("D=%d N=%d K=%d %s=%g: logZ=%9.3f bias=%+7.3f = %+6.1f sigma  H=%6.2f  KS p=%.3g  acc=%.3f nmcmc_end=%d rho=%.3f steps=%.3g fail=%d batches=%d run %.2fs -> %.3g steps/s" % (
        D, N, K, mode, v, trap, trap - ana, (trap - ana) / sig, s["info"], stats.kstest(ins, "uniform").pvalue,
        s["total_accepted"] / max(1, s["total_steps"]), s0["nmcmc"], s0["rho_max"], s["total_steps"], s["failed_chains"], nb, t2 - t1,
        s["total_steps"] / (t2 - t1)), flush=True)
    e.close()
