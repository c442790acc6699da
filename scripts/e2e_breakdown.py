"""Where the time of one host-protocol step (bench.py e2e leg) goes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
DIM, N, K = 50, 10000, 4704
np.random.seed(1234)
e = Engine("gaussian_iid", [[-10, 10]] * DIM, nlive=N, maxmcmc=5000, seed=7)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
e.run(K, max_batches=60)
live_t = torch.from_numpy(e.live()).pin_memory(); live = live_t.numpy(); e.set_live(live)
rs = np.random.Generator(np.random.PCG64(1))
pin = lambda *s, dt=torch.float64: torch.empty(s, dtype=dt).pin_memory()
st_t, ac_t = pin(K, dt=torch.int32), pin(K, dt=torch.int32)
stp, acc_ = st_t.numpy(), ac_t.numpy()
nm = e.state()["nmcmc"]
T = np.zeros(3)
n = 40
for i in range(n + 5):
    t0 = time.perf_counter()
    logL = live[:, DIM]
    part = np.argpartition(logL, K - 1).astype(np.int32)
    worst, surv = part[:K], part[K:]
    logLmin = logL[part[K - 1]]
    t1 = time.perf_counter()
    starts = surv[rs.integers(0, N - K, size=K, dtype=np.int32)]
    t2 = time.perf_counter()
    e.evolve_host_table(live, starts, worst, logLmin, nm, steps=stp, accepted=acc_, rows_async=True)
    t3 = time.perf_counter()
    if i >= 5:
        T += [t1 - t0, t2 - t1, t3 - t2]
print("per step [ms]: select %.3f  draw starts %.3f  cpn_evolve_host_table %.3f  (chain length %d)" % (*(1e3 * T / n), nm))
