"""Event-timed launches of the MH kernel on the correlated 50-D Gaussian: python scripts/time_corr.py [lanes]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
from oracle import cpnest_oracle as O          # (measurement tooling: builds the model's covariance)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 0
K = 4704
m = O.make_correlated_gaussian(50)
np.random.seed(1234)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
e = Engine("gaussian_corr", m.bounds, nlive=10000, maxmcmc=5000, seed=1234, params=m.params, lanes=lanes)
e.set_stream(stream.cuda_stream)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
for _ in range(30):
    e.ns_step(K)
ms, st, ev = [], [], []
for _ in range(10):
    e.select_and_accumulate(K)
    a = e.state()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); e.evolve_batch(K); e1.record(stream)
    e.insert_batch(K); torch.cuda.synchronize()
    b = e.state()
    ms.append(e0.elapsed_time(e1)); st.append(b["total_steps"] - a["total_steps"]); ev.append(b["total_evals"] - a["total_evals"])
ms, st, ev = np.array(ms), np.array(st), np.array(ev)
print("lanes %d: %.3f ms/launch, %.3e steps, %.3e evals: %.4f ns/chain-step, %.4f ns/eval, %.3f TFLOP/s algorithmic (2650 flop/eval + 308/step)" % (
    lanes, ms.mean(), st.mean(), ev.mean(), 1e6 * ms.sum() / st.sum(), 1e6 * ms.sum() / ev.sum(),
    (ev.sum() * 2650.0 + st.sum() * 308.0) / (ms.sum() * 1e-3) / 1e12))
