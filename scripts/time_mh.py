"""Event-timed launches of the production MH kernel on the bench workload (50-D Gaussian, nlive 1e4, K chains per batch).
usage: python scripts/time_mh.py [K] [lanes] [warm batches] [timed batches] [functor]
prints ns per chain-step (kernel time / MCMC steps of the launch) over the timed batches."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4704
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 30
timed = int(sys.argv[4]) if len(sys.argv) > 4 else 20
N = int(sys.argv[5]) if len(sys.argv) > 5 else 10000
np.random.seed(1234)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
e = Engine("gaussian_iid", [[-10, 10]] * 50, nlive=N, maxmcmc=5000, seed=1234, lanes=lanes)
e.set_stream(stream.cuda_stream)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
for _ in range(warm):
    e.ns_step(K)
ms, st = [], []
for _ in range(timed):
    e.select_and_accumulate(K)
    a = e.state()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e.evolve_batch(K)
    e1.record(stream)
    e.insert_batch(K)
    torch.cuda.synchronize()
    b = e.state()
    ms.append(e0.elapsed_time(e1))
    st.append(b["total_steps"] - a["total_steps"])
ms, st = np.array(ms), np.array(st)
print("K %d lanes %d: kernel %.4f ms/launch, %.3e steps/launch, %.4f ns/chain-step (min %.4f), nmcmc %d, evals/step %.3f" % (
    K, lanes, ms.mean(), st.mean(), 1e6 * ms.sum() / st.sum(), 1e6 * (ms / st).min(), b["nmcmc"],
    b["total_evals"] / max(1, b["total_steps"])))
