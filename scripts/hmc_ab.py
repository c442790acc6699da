"""A/B timing of a sampler on one workload with the package found under <root> (default: this repo).
usage: python scripts/hmc_ab.py <root> [sampler] [D] [N] [K] [batches]"""
import sys
import time
root = sys.argv[1]
sys.path.insert(0, root)
import numpy as np
import torch
from cpnest_b200.engine import Engine
sampler = sys.argv[2] if len(sys.argv) > 2 else "hmc"
D = int(sys.argv[3]) if len(sys.argv) > 3 else 20
N = int(sys.argv[4]) if len(sys.argv) > 4 else 100000
K = int(sys.argv[5]) if len(sys.argv) > 5 else 4704
nb = int(sys.argv[6]) if len(sys.argv) > 6 else 60
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=200, seed=1234, sampler=sampler)
e.set_stream(stream.cuda_stream)
e.init_prior()
for _ in range(20):
    e.ns_step(K)
torch.cuda.synchronize()
s0 = e.state()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(nb)]
t0 = time.perf_counter()
for i in range(nb):
    e.select_and_accumulate(K)
    ev[i][0].record(stream); e.evolve_batch(K); ev[i][1].record(stream)
    e.insert_batch(K)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
s = e.state()
k = np.array([a.elapsed_time(b) for a, b in ev])
print(root, sampler, "D", D, "nmcmc", s["nmcmc"], "hmc_dt %.5f" % s["hmc_dt"], "evals/batch %.4g" % ((s["total_evals"] - s0["total_evals"]) / nb),
      "steps/batch %.4g" % ((s["total_steps"] - s0["total_steps"]) / nb), "wall/batch %.3f ms" % (1e3 * dt / nb),
      "kernel ms median %.3f mean %.3f" % (np.median(k), k.mean()), "evals/s %.4g" % ((s["total_evals"] - s0["total_evals"]) / dt))
