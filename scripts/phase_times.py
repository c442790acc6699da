"""CUDA-event timing of the three phases of a batch (select+accumulate, evolve, insert) in mid-run."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4704
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
np.random.seed(1234)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
e = Engine("gaussian_iid", [[-10, 10]] * 50, nlive=N, maxmcmc=5000, seed=1234, lanes=lanes)
e.set_stream(stream.cuda_stream)
e.set_cycle(DefaultProposalCycle().device_cycle())
e.init_prior()
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 60
for _ in range(warm):
    e.ns_step(K)
torch.cuda.synchronize()
n = int(sys.argv[5]) if len(sys.argv) > 5 else 60
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n)]
for i in range(n):
    ev[i][0].record(stream); e.select_and_accumulate(K)
    ev[i][1].record(stream); e.evolve_batch(K)
    ev[i][2].record(stream); e.insert_batch(K)
    ev[i][3].record(stream)
torch.cuda.synchronize()
t = np.array([[ev[i][k].elapsed_time(ev[i][k + 1]) for k in range(3)] for i in range(n)])
tot = ev[0][0].elapsed_time(ev[n - 1][3]) / n
print("lanes=%d " % lanes + "K=%d mean ms: select %.3f evolve %.3f insert %.3f | sum %.3f, wall/step %.3f" % (K, *t.mean(0), t.mean(0).sum(), tot))
pass
ev_t = t[:, 1]
print("evolve ms: min %.3f median %.3f max %.3f; batches slower than 1.1 x median: %d of %d" % (
    ev_t.min(), np.median(ev_t), ev_t.max(), int((ev_t > 1.1 * np.median(ev_t)).sum()), n))
print("select ms: median %.3f max %.3f; insert ms: median %.3f max %.3f" % (
    np.median(t[:, 0]), t[:, 0].max(), np.median(t[:, 2]), t[:, 2].max()))
print("per batch [select evolve insert]:", " ".join("[%.3f %.3f %.3f]" % tuple(r) for r in t[:16]))
