import sys, time, numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
N, K = 10000, 2500
np.random.seed(1234)
e = Engine("gaussian_iid", [[-10, 10]] * 50, nlive=N, maxmcmc=5000, seed=1234)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
for b in range(20): e.ns_step(K)
e.synchronize()
t0 = time.perf_counter()
ts = []
for b in range(300):
    t1 = time.perf_counter(); e.ns_step(K); ts.append(time.perf_counter() - t1)
t_enq = time.perf_counter() - t0
e.synchronize()
dt = time.perf_counter() - t0
ts = np.array(ts)
print("300 steps: enqueue %.3f s, total %.3f s; per-call host time median %.1f us, max %.1f ms; slow calls (>5ms): %s" % (
    t_enq, dt, 1e6 * np.median(ts), 1e3 * ts.max(), [(i, round(1e3 * v)) for i, v in enumerate(ts) if v > 5e-3]))
