import sys, time, numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
N, K = 10000, 2500
np.random.seed(1234)
e = Engine("gaussian_iid", [[-10, 10]] * 50, nlive=N, maxmcmc=5000, seed=1234)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
prev = e.state()
for b in range(int(sys.argv[1]) if len(sys.argv) > 1 else 200):
    t0 = time.perf_counter(); e.ns_step(K); e.synchronize(); dt = time.perf_counter() - t0
    s = e.state()
    st = s["total_steps"] - prev["total_steps"]; ev = s["total_evals"] - prev["total_evals"]; ac = s["total_accepted"] - prev["total_accepted"]
    if b % 10 == 0 or b < 5:
        _, steps, acc = e.last_batch(K)
        print("b=%3d %.2f ms nmcmc(next)=%4d steps/chain mean %.0f max %d  evals/steps %.2f acc %.3f  rho %.3f  %.3g steps/s logLmin %.1f fail %d" % (
            b, 1e3 * dt, s["nmcmc"], steps.mean(), steps.max(), ev / st, ac / st, s["rho_max"], st / dt, s["logLmin"], s["failed_chains"]), flush=True)
    prev = s
