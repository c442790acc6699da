"""A small tour of every kernel (written for compute-sanitizer memcheck, which this GPU pool does not allow; it still serves as
a quick whole-library exercise): MH with zero-padded rows and with masks, slice, HMC, a mixed population, the mixture functor in
team mode, the host-table protocol, checkpoint, ensemble statistics, the final sweep.
usage: [compute-sanitizer --tool memcheck] python scripts/sanitizer_workload.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cpnest_b200.engine import Engine


def tour(tag, **kw):
    K = kw.pop("K", 48)
    nb = kw.pop("nb", 4)
    e = Engine(**kw)
    e.init_prior()
    for _ in range(nb):
        e.ns_step(K)
    blob = e.save_checkpoint()
    e.load_checkpoint(blob)
    e.ns_step(K)
    e.run(K, max_batches=3)
    e.ensemble_stats()
    e.finalise()
    s = e.state()
    print(tag, "ok: iteration", s["iteration"], "logZ %.3f" % s["logZ"], flush=True)
    e.close()


b10 = [[-10, 10]] * 10
tour("mh 10-D padded rows", functor="gaussian_iid", bounds=b10, nlive=200, maxmcmc=40)
os.environ["CPN_NO_ROW_PAD"] = "1"
tour("mh 10-D masked", functor="gaussian_iid", bounds=b10, nlive=200, maxmcmc=40)
del os.environ["CPN_NO_ROW_PAD"]
tour("mh 50-D", functor="gaussian_iid", bounds=[[-10, 10]] * 50, nlive=300, maxmcmc=60, K=100)
tour("mh 7-D lanes 32", functor="gaussian_iid", bounds=[[-10, 10]] * 7, nlive=150, maxmcmc=30, lanes=32, K=37)
tour("slice 10-D", functor="rosenbrock", bounds=[[-5, 5]] * 10, nlive=200, maxmcmc=30, sampler="slice")
tour("hmc 10-D", functor="gaussian_iid", bounds=b10, nlive=200, maxmcmc=20, sampler="hmc")
tour("mixed 2:1:1", functor="gaussian_iid", bounds=[[-10, 10]] * 5, nlive=200, maxmcmc=30, mix=(2, 1, 1), K=50)
tour("eggbox", functor="eggbox", bounds=[[0, 10 * np.pi]] * 2, nlive=200, maxmcmc=30)
rs = np.random.RandomState(1)
data = np.concatenate((rs.normal(0.5, 0.5, size=100), rs.normal(-1.5, 0.03, size=901)))
tour("mixture team mode", functor="gaussian_mixture", bounds=[[-3, 3], [0.01, 1], [-3, 3], [0.01, 1], [0.0, 1.0]], nlive=120, maxmcmc=20,
     params=dict(data=data), K=19, nb=2)
# host-table protocol
D, N, K = 6, 200, 50
e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=30, seed=3)
e.init_prior()
table_t = torch.from_numpy(e.live()).pin_memory()
table = table_t.numpy()
e.set_live(table)
for _ in range(3):
    part = np.argpartition(table[:, D], K - 1).astype(np.int32)
    e.evolve_host_table(table, part[K:][:K], part[:K], table[part[K - 1], D], 20, rows_async=True)
req = np.ascontiguousarray(table[:K])
e.evolve_host(req, float(table[:, D].min()), 20, slots=np.arange(K, dtype=np.int32))
print("host protocols ok", flush=True)
e.close()
