#!/usr/bin/env python
"""bench.py -- likelihood evaluations / second of the nested-sampling hot path on B200.

Workload (BASELINE.json north_star: "50-D Gaussian at nlive=10^4 on 1 B200"): the model of the
reference's examples/test_50d.py (iid unit Gaussian, bounds [-10,10]^50, analytic logZ
-149.786614), nlive = 10^4 per GPU, MH sampler with the DefaultProposalCycle
(Walk/Stretch/DE/EigenVector, weights 5:5:10:10).

A "step" is one nested-sampling batch = one `consume_sample()` of the reference
(cpnest/NestedSampling.py:318-366): select the K worst live points + accumulate the evidence,
evolve K constrained MCMC chains (the hot kernel), insert the K replacements.

  value       : evals/s with the live set resident in HBM, CUDA-event timed, max over ranks, over K steps in the bulk
                of a run (the run is fast-forwarded, untimed, past its first FAST_FORWARD batches, where chains are
                several times longer and most proposals leave the prior box: evals/s differs 2x between the phases)
  logz_wall_s : wall time of ONE complete run, prior draw -> final evidence, on the N GPUs; full_run.evals_per_s is the
                whole-run average (every phase)
  e2e         : the same batches driven through the sampler pipe protocol of the C-ABI with HOST
                buffers (cpn_evolve_host_table): the host keeps the live set in one pinned table and
                selects the worst points; every step the K request rows travel host->device and the K
                replies device->host (read from / written into the table by the device itself).
  matched     : both arms evolve chains on the same kind of ensemble, threshold and chain length (pool drawn from N(0,1)^50,
                threshold at its 20% logL quantile, Nmcmc = maxmcmc = 100) - here through cpn_evolve_host, in the reference
                arm by its own MetropolisHastingsSampler objects
  roofline / cpu_baseline : see DESIGN.md "Measurement".

`--impl reference` times the UNMODIFIED reference (oracle/_ref) through its own public API,
cpnest.CPNest(model, nlive, nthreads=cores, ...).run(), on the host cores, observed for a bounded time.
One process per GPU under torchrun (RANK/LOCAL_RANK/WORLD_SIZE).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM = 50
NLIVE_PER_GPU = 10000
K_PER_GPU = 4704          # 147 SMs x 32 resident chains: one wave of the chain kernel, with room left on one SM for the
                          # ensemble-statistics kernels that run beside it on the side stream
MAXMCMC = 5000
REF_MAXMCMC = 100         # the reference arm: examples/test_50d.py:37 (with maxmcmc = 5000 its start-up alone - poolsize x
                          # Nmcmc = 500 steps per sampler, then nlive chains - takes more than 7 minutes on 8 cores)
SEED = 1234
FAST_FORWARD = 40         # batches of a run before the bulk phase (chain length has settled, cloud well inside the box)
ANALYTIC_LOGZ = -DIM * math.log(20.0)
METRIC = "likelihood evals/sec (50-D Gaussian, nlive=1e4 per GPU, MH DefaultProposalCycle)"
UNIT = "likelihood evals/s"

WORKLOADS = {
    # name: (functor, dim, bounds, params, sampler, analytic logZ or None)
    "gauss50": ("gaussian_iid", 50, (-10.0, 10.0), None, "mh", -50 * math.log(20.0)),
    "rosen10": ("rosenbrock", 10, (-5.0, 5.0), None, "slice", None),
    "gauss20": ("gaussian_iid", 20, (-10.0, 10.0), None, "hmc", -20 * math.log(20.0)),
    # (HMC as the reference specifies it - mass matrix from the global ensemble covariance - has no inter-mode move and
    # loses a mode of the separated mixture, scripts/baseline_configs.py: the mixture runs with MH)
    "mix20": ("gaussian_mixture_nd", 20, (-10.0, 10.0), None, "mh", -20 * math.log(20.0)),
    # BASELINE config 4, correlated flavour: N(0, Sigma), condition number 1e2; the D^2 likelihood runs on the
    # FP64 tensor cores (models.cuh GaussCorr::log_likelihood_block); parameters filled in below
    "corr50": ("gaussian_corr", 50, (-10.0, 10.0), "corr", "mh", -50 * math.log(20.0)),
}


def shared_config(args):
    """The workload, named identically by both arms (the driver compares the two `config` objects)."""
    return {"workload": "gaussian50d_iid_bounds[-10,10]_nlive%d_per_gpu_mh_default_cycle" % args.nlive, "dim": DIM,
            "nlive_per_gpu": args.nlive, "sampler": "mh", "proposals": "DefaultProposalCycle (5:5:10:10)",
            "seed": SEED, "model": "examples/test_50d.py GaussianModel(dim=50)"}


def fp64_peak():
    """Measured DFMA peak of this pool's B200 (scripts/fp_peak.cu -> profiles/fp_peaks.json); MEASURED_PEAKS.json has
    no FP64 figure."""
    p = os.path.join(ROOT, "profiles", "fp_peaks.json")
    if os.path.exists(p):
        return float(json.load(open(p))["fp64_fma_tflops"]), "measured (profiles/fp_peaks.json, scripts/fp_peak.cu)"
    return 34.0, "fallback (earlier measurement on this pool)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle sampling during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.maxclk = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxclk = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxclk, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxclk, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host cores, through its public API."""
    if rank != 0:
        return
    from oracle import ref_bench
    cores = os.cpu_count() or 1
    cfg = shared_config(args)
    # a step is a bounded observation window of the run; the whole arm stays within a few minutes whatever --steps / --warmup
    per_step_s = min(5.0, max(0.25, 90.0 / max(1, args.steps + min(args.warmup, 5))))
    warm = min(args.warmup, 5)
    matched = None
    if ref_bench.reference_available():
        kind = "reference"
        # The stock path.  The reference's start-up (NestedSampler.reset, cpnest/NestedSampling.py:404-431) hung in one of
        # three starts when tried at these sizes (a reply that fails its finiteness check desynchronises the pipes and the
        # parent waits for ever: SURVEY.md 8b "errors: none propagated"); the run is given 150 s to reach its nested-sampling
        # loop and is started afresh (another seed) up to three times.
        marks, prog, t_ready, attempts = None, None, None, 0
        for attempt in range(3):
            attempts += 1
            run = ref_bench.StockReferenceRun(dim=DIM, nlive=args.nlive, cores=cores, maxmcmc=REF_MAXMCMC, poolsize=1000,
                                              seed=SEED + attempt)
            try:
                if not run.wait_ready(timeout_s=150.0):
                    continue
                t_ready = time.perf_counter() - run.t0
                for _ in range(warm):
                    time.sleep(per_step_s)
                marks = [run.counts()]
                for _ in range(args.steps):
                    time.sleep(per_step_s)
                    marks.append(run.counts())
                prog = run.progress()
                break
            finally:
                run.stop()
        if marks is not None and marks[-1][0] > marks[0][0]:
            e0, s0, t0 = marks[0]
            e1, s1, t1 = marks[-1]
            wall = t1 - t0
            value = (e1 - e0) / wall
            steps_per_s = (s1 - s0) / wall
            sample = ("unmodified reference (oracle/_ref): cpnest.CPNest(GaussianModel(50), nlive=%d, nthreads=%d, maxmcmc=%d, poolsize=1000, "
                      "seed=%d).run() in a child process group (start %d of 3), %d windows of %.2f s after %.0f s of start-up (sampler "
                      "reset + initial live set); likelihood calls and MCMC steps counted by a wrapper model (multiprocessing.Value, "
                      "bumped every 1000 calls); run at NS iteration %s, Nmcmc %s when stopped" % (
                          args.nlive, cores, REF_MAXMCMC, SEED + attempts - 1, attempts, args.steps, per_step_s, t_ready,
                          prog["iteration"] if prog else "?", prog["nmcmc"] if prog else "?"))
        else:
            value = None
        # the matched regime: the reference's own sampler objects on a fixed ensemble / threshold / chain length
        pool = ref_bench.ReferencePool(dim=DIM, cores=cores)
        pool.leg(1.0)
        m = pool.leg(8.0)
        pool.close()
        matched = {"value": m["value"], "unit": UNIT, "steps_per_s": m["steps_per_s"], "cores": cores,
                   "regime": "pool of 1000 points ~ N(0,1)^50 per process, threshold at the 20% logL quantile, Nmcmc = maxmcmc = 100",
                   "sample": m["sample"]}
        if value is None:
            # the stock run never reached its loop: the sampler-level figure stands in, and says so
            value, steps_per_s, wall = m["value"], m["steps_per_s"], per_step_s * args.steps
            sample = "stock CPNest.run() hung in its start-up 3 times; " + m["sample"]
    else:
        kind = "port"
        res = [ref_bench.time_reference(dim=DIM, cores=cores, budget_s=per_step_s) for _ in range(min(args.steps, 8))]
        value = float(np.mean([r["value"] for r in res]))
        steps_per_s = float(np.mean([r["steps_per_s"] for r in res]))
        wall = per_step_s * args.steps
        sample = res[-1]["sample"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "arm": {"step": "bounded observation window of %.2f s of the running reference" % per_step_s, "nthreads": cores,
                "maxmcmc": REF_MAXMCMC,
                "phase": "start of the run (a 50-D run of the reference takes hours; its evals/s grows ~3x towards the bulk, "
                         "BASELINE.md section 2)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mcmc_steps_per_s": steps_per_s,
        "matched": matched,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def state_hash(eng, n_dead=None):
    """SHA-1 over the live set (sorted by logL, as cpn_get_live returns it) and the dead points so far."""
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(eng.live()).tobytes())
    nd = eng.num_dead() if n_dead is None else n_dead
    if nd > 0:
        h.update(np.ascontiguousarray(eng.dead(0, nd)).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nlive", type=int, default=NLIVE_PER_GPU)
    ap.add_argument("--k", type=int, default=K_PER_GPU)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--comm", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU exchange of the replacements: fused into the chain kernel over peer memory, or NCCL all-gather")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: nlive and K per GPU fixed (default); strong: --nlive and --k are totals shared by all GPUs")
    ap.add_argument("--workload", default="gauss50", choices=sorted(WORKLOADS),
                    help="gauss50 is the headline; the others (BASELINE configs 3 and 5) add the `sampler_run` key only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-run", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from cpnest_b200.engine import Engine
    from cpnest_b200.proposal import DefaultProposalCycle

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; cpnest_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    # a real (non-default) stream shared by torch's events/collectives and the engine's kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    functor, dim, (blo, bhi), fparams, sampler, analytic = WORKLOADS[args.workload]
    if fparams == "corr":
        from cpnest_b200.models import CorrelatedGaussianModel
        fparams = CorrelatedGaussianModel(dim=dim, bound=bhi).device_functor[1]
    bounds = [[blo, bhi]] * dim
    np.random.seed(SEED)
    cycle = DefaultProposalCycle().device_cycle()        # drawn like the reference's (proposal.py:71-75)
    if args.scaling == "weak":
        # nlive and chains per batch grow with the number of GPUs; every rank holds a replica of the live set and evolves
        # its slice of the chains (cpnest_b200/sharded.py)
        N, K = args.nlive * world, args.k * world
    else:
        N, K = args.nlive, args.k

    def make_engine(seed, n=None, rk=rank, ws=world, lanes=None):
        e = Engine(functor, bounds, nlive=n or N, maxmcmc=MAXMCMC, seed=seed, device=local_rank, lanes=args.lanes if lanes is None else lanes,
                   rank=rk, world_size=ws, params=fparams, sampler=sampler)
        e.set_stream(stream.cuda_stream)
        e.set_cycle(cycle)
        return e

    def make_driver(e):
        if world > 1:
            from cpnest_b200.sharded import ShardedEngine
            return ShardedEngine(e, comm=args.comm)
        return e

    def sigma_inflation(k, n):
        # a batch removes its K worst points before any replacement arrives: the variance of the log-volume per unit of
        # shrinkage grows by (1/(1-f) - 1) / (-ln(1-f)), f = K / nlive (profiles/r01_logz_validation.txt)
        f = min(k / float(n), 0.999)
        return math.sqrt((1.0 / (1.0 - f) - 1.0) / -math.log1p(-f))

    # ------------------------------------------------------------------------------------------
    # one complete run: logZ wall time and the whole-run rate
    # ------------------------------------------------------------------------------------------
    full = None
    logz_wall_s = None
    if not args.no_full_run:
        eng = make_engine(SEED)
        drv = make_driver(eng)
        eng.init_prior()                                 # warm (module load, allocations, first use of the exchange): not
        for _ in range(3):                               # part of the timed run
            drv.ns_step(K)
        eng.synchronize()
        eng.set_seed(SEED + 1)
        barrier()
        t0 = time.perf_counter()
        eng.init_prior()                                 # same seed everywhere: identical replicas, no broadcast
        drv.run(K, tolerance=0.1)
        st = eng.state()
        rect, trap = eng.finalise()
        torch.cuda.synchronize()
        logz_wall_s = max_over_ranks(time.perf_counter() - t0)
        sf = eng.state()
        sigma = math.sqrt(sf["info"] / N)
        infl = sigma_inflation(K, N)
        full = {"logZ": trap, "analytic_logZ": analytic, "sigma": sigma, "sigma_batch_inflated": sigma * infl,
                "nsigma": (trap - analytic) / sigma if analytic is not None else None,
                "nsigma_batch_inflated": (trap - analytic) / (sigma * infl) if analytic is not None else None,
                "H": sf["info"], "batches": st["batches"], "dead_points": sf["iteration"], "nmcmc_final": st["nmcmc"],
                "total_evals": sf["total_evals"], "total_mcmc_steps": sf["total_steps"], "total_accepted": sf["total_accepted"],
                "failed_chains": sf["failed_chains"], "wall_s": logz_wall_s,
                "evals_per_s": sf["total_evals"] / logz_wall_s, "mcmc_steps_per_s": sf["total_steps"] / logz_wall_s,
                "accepted_chain_steps_per_s": sf["total_accepted"] / logz_wall_s,
                "what": "init_prior -> run to the stopping condition 0.1 -> final sweep, wall clock, max over ranks"}
        if world > 1:
            drv.close()
        eng.close()

    if args.workload != "gauss50":
        # BASELINE configs 3 / 5 on the full box: the other sampler kinds sharded over the GPUs; not a headline
        if rank == 0:
            print(json.dumps({"metric": "whole-run likelihood evals/s (%s, %s sampler)" % (args.workload, sampler), "unit": UNIT,
                              "value": full["evals_per_s"] if full else None, "n_gpus": world, "scaling": args.scaling,
                              "config": {"workload": args.workload, "functor": functor, "dim": dim, "nlive": N, "K": K, "sampler": sampler},
                              "logz_wall_s": logz_wall_s, "sampler_run": full,
                              "algorithmic_fp64_tflops": ((full["total_evals"] * 2650.0 + full["total_mcmc_steps"] * 308.0) / logz_wall_s / 1e12
                                                          if full and args.workload == "corr50" else None)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------------------------------
    # device-resident run: `value`
    # ------------------------------------------------------------------------------------------
    # The cost of a batch and the number of likelihood calls it makes both change along a run (first batches: chains 3x
    # longer, 0.13 calls per step; last batches: 0.34), so K consecutive steps give a number that depends on where they
    # sit.  The K timed steps are therefore spread evenly over one complete run (every `stride`-th batch is timed, each
    # with its own pair of events; the batches between them run untimed): `value` estimates the whole-run rate for any K.
    eng = make_engine(SEED + 2)
    drv = make_driver(eng)
    B_est = int(full["batches"]) if full else 190
    stride = max(1.0, B_est / float(args.steps))
    eng.init_prior()
    for _ in range(FAST_FORWARD + args.warmup):          # warm-up: untimed batches into the bulk of a throw-away run
        drv.ns_step(K)
    eng.set_seed(SEED + 3)
    eng.init_prior()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    rec = []                                             # per timed step: ms, kernel ms, d steps, d evals, d accepted, launches
    nm_sched = []                                        # chain length of every batch of the run (for the host-driven run)
    timed_at = []
    g, next_t, restarts = 0, 0.5 * stride, 0
    while len(rec) < args.steps:
        a = eng.state()
        if a["done"] or not (a["condition"] > 0.1):
            restarts += 1
            eng.set_seed(SEED + 3 + 7919 * restarts)
            eng.init_prior()
            a = eng.state()
        if restarts == 0:
            nm_sched.append(a["nmcmc"])
        if g >= int(next_t):
            l0 = eng.launch_count()
            e0, e1, k0, k1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record(stream)
            eng.select_and_accumulate(K)
            k0.record(stream)
            eng.evolve_batch(K)                          # exactly one launch of k_mh_chains (this rank's chains)
            k1.record(stream)
            if world > 1:
                drv.exchange(K)
            eng.insert_batch(K)
            e1.record(stream)
            b = eng.state()                              # synchronises
            rec.append((e0.elapsed_time(e1), k0.elapsed_time(k1), b["total_steps"] - a["total_steps"],
                        b["total_evals"] - a["total_evals"], b["total_accepted"] - a["total_accepted"], eng.launch_count() - l0))
            timed_at.append(g)
            next_t += stride
        else:
            drv.ns_step(K)
        g += 1
    barrier()
    clk = clocks.result()
    # finish the run of the schedule (untimed) so that the host-driven run below can follow all of it
    while restarts == 0:
        a = eng.state()
        if a["done"] or not (a["condition"] > 0.1):
            break
        nm_sched.append(a["nmcmc"])
        drv.ns_step(K)
    rec = np.array(rec, dtype=np.float64)
    dt_ms = max_over_ranks(rec[:, 0].sum())
    # the counters are sums over the gathered per-chain outcomes: already global on every rank
    steps_mcmc, evals, accepted = rec[:, 2].sum(), rec[:, 3].sum(), rec[:, 4].sum()
    launches = int(rec[:, 5].sum())
    value = evals / (dt_ms * 1e-3)
    nm_bulk = int(np.median(nm_sched[len(nm_sched) // 3:])) if nm_sched else 1000

    # ---- the dominant kernel alone: k_mh_chains over the same timed steps ----
    k_ms = float(rec[:, 1].mean())
    k_steps = float(rec[:, 2].mean()) / world
    k_evals = float(rec[:, 3].mean()) / world
    # algorithmic work of one launch: FP64: 6.17*D flops per proposal + box test, 3*D per likelihood
    # evaluation of the iid Gaussian; gathers: 13.33*D bytes per MCMC step of the default cycle + read and write one
    # point (8D+16 bytes) per chain
    alg_flops = k_steps * 6.17 * DIM + k_evals * 3.0 * DIM
    f64_peak, f64_src = fp64_peak()
    f64_achieved = alg_flops / (k_ms * 1e-3) / 1e12
    bytes_per_step = 8.0 * DIM * (3 / 6 + 1 / 6 + 2 / 3 + 1 / 3)
    alg_bytes = k_steps * bytes_per_step + (K / world) * 2.0 * (8 * DIM + 16)
    hbm_peak, hbm_src = peaks()
    traffic = None
    ncu = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            ncu = json.load(open(tp))
            traffic = ncu.get("k_mh_chains_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # The kernel is bound by neither HBM nor the tensor cores: the live set is L2 resident (DRAM sees each row once per
    # launch) and the work is FP64.  Of the two figures SURVEY 8d asks for, the FP64 one measures a resource the kernel
    # uses and leads; the gather rate is reported as what it is, bytes served by L2/L1.
    roofline = {"bound": "latency (dependent FP64/shuffle chain per MCMC step at the occupancy K chains allow; not hbm, not tensor)",
                "achieved": f64_achieved, "peak": f64_peak, "unit": "TFLOP/s", "frac": f64_achieved / f64_peak,
                "traffic": traffic, "kernel": "k_mh_chains", "ms_per_launch": k_ms, "mcmc_steps_per_launch": k_steps,
                "evals_per_launch": k_evals, "ns_per_chain_step": 1e6 * k_ms / k_steps,
                "kernel_share_of_step": float(rec[:, 1].sum() / rec[:, 0].sum()),
                "algorithmic_flops_per_launch": alg_flops, "peak_source": f64_src,
                "issue_slots_pct": ncu.get("k_mh_chains_issue_active_pct"), "fp64_pipe_pct": ncu.get("k_mh_chains_fp64_pipe_pct"),
                "l1_data_pipe_pct": ncu.get("k_mh_chains_l1tex_lsu_wavefronts_pct"),
                "warp_instructions_per_chain_step": ncu.get("k_mh_chains_warp_inst_per_chain_step"),
                "gathers_l2_served": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "unit": "GB/s", "algorithmic_bytes_per_launch": alg_bytes,
                                      "served_from": "L2/L1 (live set %.1f MB)" % (N * 64 * 8 / 1e6),
                                      "hbm_peak_for_scale": hbm_peak, "hbm_peak_source": hbm_src},
                "hbm": {"achieved": (traffic / (k_ms * 1e-3) / 1e9) if traffic else None, "peak": hbm_peak, "unit": "GB/s",
                        "note": "actual DRAM traffic of the launch (ncu) over its duration"}}
    if world > 1:
        drv.close()
    eng.close()

    # ------------------------------------------------------------------------------------------
    # multi-GPU: the sharded run equals the single-GPU run of the same (nlive, K), bit for bit
    # ------------------------------------------------------------------------------------------
    sharded_equals_single = None
    replicas_identical = None
    if world > 1:
        NB = 6
        # same lanes per chain in both (the likelihood sum is reduced over them: its rounding depends on their number; a
        # single GPU with all K chains would by itself pick narrower groups than a rank with K / world chains)
        probe = make_engine(SEED + 5)
        lanes_eq = probe.pick_lanes(K // world)[0]
        probe.close()
        e_sh = make_engine(SEED + 5, lanes=lanes_eq)
        d_sh = make_driver(e_sh)
        e_sh.init_prior()
        for _ in range(NB):
            d_sh.ns_step(K)
        torch.cuda.synchronize()
        h_sh = state_hash(e_sh)
        d_sh.close()
        e_sh.close()
        hs = [None] * world
        dist.all_gather_object(hs, h_sh)
        replicas_identical = len(set(hs)) == 1
        if rank == 0:
            e_1 = make_engine(SEED + 5, rk=0, ws=1, lanes=lanes_eq)      # all K chains of every batch on one GPU
            e_1.init_prior()
            for _ in range(NB):
                e_1.ns_step(K)
            torch.cuda.synchronize()
            sharded_equals_single = bool(state_hash(e_1) == h_sh)
            e_1.close()
        barrier()

    # ------------------------------------------------------------------------------------------
    # host-buffer protocol: `e2e` (every rank serves its own K requests per step from its own host)
    # ------------------------------------------------------------------------------------------
    # The caller (bench.py, playing NestedSampler.consume_sample) drives a whole run from the host: its live set is one
    # pinned host table, it selects the K worst points and the threshold on the host and draws the start points; the chain
    # length of batch b is the one the device-resident run used at batch b.  Timed: the same stratified steps as above,
    # each from the selection to the point where the replies have landed in the host table.
    Ne, Ke = (args.nlive, args.k) if args.scaling == "weak" else (max(8, args.nlive // world), max(1, args.k // world))
    eng = make_engine(SEED + 1000 * rank + 7, n=Ne, rk=0, ws=1)
    eng.init_prior()
    live_t = torch.from_numpy(eng.live()).pin_memory()  # the caller's live set: ONE pinned host table
    live = live_t.numpy()
    rs = np.random.Generator(np.random.PCG64(SEED + rank))        # host draws of the start points
    st_t = torch.empty(Ke, dtype=torch.int32).pin_memory()
    ac_t = torch.empty(Ke, dtype=torch.int32).pin_memory()
    stp, acc_ = st_t.numpy(), ac_t.numpy()
    eng.set_live(live)                                  # the device pool mirrors the host live set slot by slot

    def host_step(nm):
        # the caller's side of the protocol, as NestedSampler.consume_sample does it (NestedSampling.py:318-366): pick the
        # K worst live points and the threshold, send K requests, put the K replies in their place.  The live set stays
        # where it is (row i <-> device pool slot i); a partial selection replaces the reference's sorted list.
        logL = live[:, DIM]
        part = np.argpartition(logL, Ke - 1).astype(np.int32)   # int32: the index type of the C-ABI, no conversions later
        worst, surv = part[:Ke], part[Ke:]
        logLmin = logL[part[Ke - 1]]                    # the K-th smallest sits at position K-1 of a partition
        starts = surv[rs.integers(0, Ne - Ke, size=Ke, dtype=np.int32)]  # requests: start points drawn among the survivors
        # request j = row starts[j] of the host table, reply j -> row worst[j]: the device reads and writes the pinned table
        eng.evolve_host_table(live, starts, worst, logLmin, nm, steps=stp, accepted=acc_, rows_async=True)

    e2e_value = None
    e2e_ms = None
    if not args.no_e2e and nm_sched:
        barrier()
        tset = set(timed_at) if restarts == 0 else set(range(0, len(nm_sched), max(1, len(nm_sched) // args.steps)))
        tset = set(b_ for b_ in tset if b_ >= 3)        # the first three batches of the run are the warm-up of this path
        t_sum, ev_sum = 0.0, 0
        for bi in range(len(nm_sched)):
            nm_b = max(20, nm_sched[bi])
            if bi in tset:
                eng.synchronize()
                ev_a = eng.state()["total_evals"]
                t0 = time.perf_counter()
                host_step(nm_b)
                eng.synchronize()                       # the step's rows have landed in the host table
                t_sum += time.perf_counter() - t0
                ev_sum += eng.state()["total_evals"] - ev_a
            else:
                host_step(nm_b)
        eng.synchronize()
        torch.cuda.synchronize()
        t_e2e = max_over_ranks(t_sum)
        e2e_value = sum_over_ranks(ev_sum) / t_e2e if t_e2e > 0 else None
        e2e_ms = 1e3 * t_e2e / max(1, len(tset))
    nm = nm_bulk
    h2d = Ke * (DIM + 2) * 8 + 2 * Ke * 4               # K request rows read from the host table + start / slot indices
    d2h = Ke * (DIM + 2) * 8 + 2 * Ke * 4               # K reply rows written into the host table + steps / accepted
    eng.close()

    # ------------------------------------------------------------------------------------------
    # matched regime (rank 0): the reference arm's ensemble, threshold and chain length through cpn_evolve_host
    # ------------------------------------------------------------------------------------------
    matched = None
    if rank == 0:
        cores = os.cpu_count() or 1
        Pm = 1000 * cores                                # the reference arm holds one pool of 1000 points per process
        rsm = np.random.RandomState(SEED)
        X = rsm.normal(size=(Pm, DIM))
        rows = np.zeros((Pm, DIM + 2))
        rows[:, :DIM] = X
        rows[:, DIM] = -0.5 * (X * X).sum(1) - 0.5 * DIM * math.log(2 * math.pi)
        em = Engine("gaussian_iid", [[-10.0, 10.0]] * DIM, nlive=Pm, maxmcmc=100, seed=SEED, device=local_rank, lanes=args.lanes, adapt_nmcmc=False)
        em.set_stream(stream.cuda_stream)
        em.set_cycle(cycle)
        em.set_live(rows)
        logLmin = float(np.quantile(rows[:, DIM], 0.2))
        good = np.where(rows[:, DIM] > logLmin)[0]
        Km = Pm - 8
        req = torch.from_numpy(rows[good[rsm.randint(0, len(good), size=Km)]]).pin_memory().numpy()
        rep = torch.empty((Km, DIM + 2), dtype=torch.float64).pin_memory().numpy()
        ms_, ma_ = np.empty(Km, dtype=np.int32), np.empty(Km, dtype=np.int32)
        for _ in range(3):
            em.evolve_host(req, logLmin, 100, replies=rep, steps=ms_, accepted=ma_)
        ea = em.state()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            em.evolve_host(req, logLmin, 100, replies=rep, steps=ms_, accepted=ma_)
        tm = time.perf_counter() - t0
        eb = em.state()
        matched = {"value": (eb["total_evals"] - ea["total_evals"]) / tm, "unit": UNIT,
                   "steps_per_s": (eb["total_steps"] - ea["total_steps"]) / tm,
                   "regime": "pool of %d points ~ N(0,1)^50, threshold at the 20%% logL quantile, Nmcmc = maxmcmc = 100, %d chains per "
                             "call through cpn_evolve_host (host request / reply buffers, copies inside the call)" % (Pm, Km),
                   "ms_per_call": 1e3 * tm / reps, "evals_per_step": (eb["total_evals"] - ea["total_evals"]) / max(1, eb["total_steps"] - ea["total_steps"])}
        em.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_bench
        cpu = ref_bench.time_reference(dim=DIM, budget_s=args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "steps_per_s")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": shared_config(args),
            "arm": {"chains_per_batch_per_gpu": args.k if args.scaling == "weak" else args.k / world, "maxmcmc": MAXMCMC,
                    "multi_gpu": ("chains sharded over %d ranks, live set (nlive=%d) replicated, exchange=%s" % (world, N, args.comm)) if world > 1 else "single GPU",
                    "chain_length": "adaptive (start/end correlation target 0.1)", "chain_length_bulk": nm_bulk,
                    "step": "one nested-sampling batch: select+accumulate, evolve K chains, insert",
                    "phase": "the %d timed steps are spread evenly over one complete run of ~%d batches (every %.2f-th batch, each "
                             "timed with its own CUDA events; batches in between untimed), after %d untimed warm-up batches of a "
                             "throw-away run: `value` is the whole-run rate whatever --steps" % (args.steps, B_est, stride, FAST_FORWARD + args.warmup),
                    "timed_batches": timed_at,
                    "l2_policy": "working set per step (live set + staging, %.1f MB) is re-gathered at random; "
                                 "no explicit flush: every step reads points rewritten by the previous one" % ((N + K) * 64 * 8 / 1e6)},
            "logz_wall_s": logz_wall_s,
            "runs_restarted_in_timed_region": restarts,
            "mcmc_steps_per_s": steps_mcmc / (dt_ms * 1e-3),
            "accepted_chain_steps_per_s": accepted / (dt_ms * 1e-3),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "cpn_evolve_host_table (sampler pipe protocol over the caller's pinned host live table: the device reads the K request rows from it and writes the K replies into it; selection of the worst points on the host while the coordinates of the previous replies are still landing)",
                    "chain_length": "that of the device-resident run, batch by batch (bulk: %d)" % nm, "ms_per_step": e2e_ms,
                    "what": "a whole run driven from the host through the protocol; timed at the same stratified batches"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "full_run": full,
            "matched": matched,
        }
        if world > 1:
            line["sharded_equals_single"] = sharded_equals_single
            line["replicas_identical"] = replicas_identical
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
