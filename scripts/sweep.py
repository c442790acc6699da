"""Throughput sweep of the MH kernel over lanes-per-chain and chains-per-batch (bench workload)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle

D = 50
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 250
for K in [int(v) for v in sys.argv[3].split(",")]:
    for lanes in [int(v) for v in sys.argv[4].split(",")]:
        np.random.seed(1234)
        e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=5000, seed=1234, lanes=lanes)
        e.set_cycle(DefaultProposalCycle().device_cycle())
        e.init_prior()
        nb = int(warm * 2500 / K)
        for _ in range(nb):
            e.ns_step(K)
        e.synchronize()
        s0 = e.state(); t0 = time.perf_counter()
        for _ in range(20):
            e.ns_step(K)
        e.synchronize()
        dt = time.perf_counter() - t0; s1 = e.state()
        print("N=%d K=%d lanes=%d: nmcmc=%d  %.3f ms/batch  %.3g steps/s  %.3g evals/s  acc=%.3f" % (
            N, K, lanes, s1["nmcmc"], 1e3 * dt / 20, (s1["total_steps"] - s0["total_steps"]) / dt,
            (s1["total_evals"] - s0["total_evals"]) / dt,
            (s1["total_accepted"] - s0["total_accepted"]) / max(1, s1["total_steps"] - s0["total_steps"])), flush=True)
        e.close()
