import sys, hashlib, numpy as np
sys.path.insert(0, ".")
from cpnest_b200.engine import Engine
from cpnest_b200.proposal import DefaultProposalCycle
def run(N, K, nb, D=50, lanes=0):
    np.random.seed(1234)
    e = Engine("gaussian_iid", [[-10, 10]] * D, nlive=N, maxmcmc=3000, seed=5, lanes=lanes)
    e.set_cycle(DefaultProposalCycle().device_cycle())
    e.init_prior()
    for _ in range(nb):
        e.ns_step(K)
    s = e.state()
    h = hashlib.sha1(e.dead().tobytes() + e.live().tobytes()).hexdigest()[:12]
    e.close()
    return s["logZ"], s["total_steps"], s["nmcmc"], h
for cfg in [(10000, 2500, 30), (2000, 256, 60), (1000, 400, 40)]:
    res = [run(*cfg) for _ in range(5)]
    print(cfg, "OK" if len(set(res)) == 1 else "NONDETERMINISTIC", res[0] if len(set(res)) == 1 else res, flush=True)
a = run(10000, 2500, 20, lanes=16); b = run(10000, 2500, 20, lanes=32); c = run(10000, 2500, 20, lanes=8)
print("lanes 16/32/8:", a, b, c)
