#!/usr/bin/env python
"""bench.py -- likelihood evaluations / second of the nested-sampling hot path on B200.

Workload (BASELINE.json north_star: "50-D Gaussian at nlive=10^4 on 1 B200"): the model of the
reference's examples/test_50d.py (iid unit Gaussian, bounds [-10,10]^50, analytic logZ
-149.786614), nlive = 10^4 per GPU, MH sampler with the DefaultProposalCycle
(Walk/Stretch/DE/EigenVector, weights 5:5:10:10).

A "step" is one nested-sampling batch = one `consume_sample()` of the reference
(cpnest/NestedSampling.py:318-366): select the K worst live points + accumulate the evidence,
evolve K constrained MCMC chains (the hot kernel), insert the K replacements.

  value : evals/s with the live set resident in HBM, CUDA-event timed, max over ranks
  e2e   : the same batches driven through the sampler pipe protocol of the C-ABI with HOST
          buffers (cpn_evolve_host_table): the host keeps the live set in one pinned table and
          selects the worst points; every step the K request rows travel host->device and the K
          replies device->host (read from / written into the table by the device itself).
  roofline / cpu_baseline : see DESIGN.md "Measurement".

`--impl reference` times the reference's own MH sampler (oracle/_ref, unmodified) on the host
cores instead.  One process per GPU under torchrun (RANK/LOCAL_RANK/WORLD_SIZE).
"""
from __future__ import annotations

import argparse
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
K_PER_GPU = 4704          # 147 SMs x 16 resident warps x 2 chains per warp: one wave of the chain kernel, with room
                          # left on one SM for the ensemble-statistics kernels that run beside it on the side stream
MAXMCMC = 5000
SEED = 1234
ANALYTIC_LOGZ = -DIM * math.log(20.0)
METRIC = "likelihood evals/sec (50-D Gaussian, nlive=1e4 per GPU, MH DefaultProposalCycle)"
UNIT = "likelihood evals/s"


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
            time.sleep(0.2)      # the recipe samples every 200 ms; NVML queries are not free

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxclk, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxclk, "reasons": sorted(self.reasons)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host cores."""
    if rank != 0:
        return
    from oracle import ref_bench
    cores = os.cpu_count() or 1
    # a step is a bounded sample of the workload; the whole arm stays within ~2.5 minutes whatever --steps / --warmup
    per_step_s = min(6.0, max(0.25, 120.0 / max(1, args.steps)))
    res = []
    if ref_bench.reference_available():
        pool = ref_bench.ReferencePool(dim=DIM, cores=cores)          # processes + sampler pools are set up once
        for _ in range(min(args.warmup, 3)):
            pool.leg(min(1.0, per_step_s))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res.append(pool.leg(per_step_s))
        wall = time.perf_counter() - t0
        pool.close()
    else:
        nsteps = min(args.steps, 8)                                   # the port pays its set-up per call
        t0 = time.perf_counter()
        for _ in range(nsteps):
            res.append(ref_bench.time_reference(dim=DIM, cores=cores, budget_s=per_step_s))
        wall = (time.perf_counter() - t0) * args.steps / nsteps
    value = float(np.mean([r["value"] for r in res]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the workload of the GPU arm (same model, sampler and proposal cycle); the reference's sampler processes evolve
        # one chain each from a pool of 1000 points, so nlive / K enter only through the sample described in `step`
        "config": {"workload": "gaussian50d_iid_bounds[-10,10]_nlive%d_per_gpu_K%d_mh_default_cycle" % (args.nlive, args.k), "dim": DIM,
                   "step": "bounded sample: %d forked sampler processes x %.1f s of yield_sample (poolsize 1000, Nmcmc 100)" % (cores, per_step_s)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": res[-1]["kind"], "sample": res[-1]["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "steps_per_s": float(np.mean([r["steps_per_s"] for r in res])),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nlive", type=int, default=NLIVE_PER_GPU)
    ap.add_argument("--k", type=int, default=K_PER_GPU)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--comm", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU exchange of the replacements: fused into the chain kernel over peer memory, or NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-run", action="store_true")
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

    # weak scaling: nlive and chains per batch grow with the number of GPUs; every rank holds a replica
    # of the live set and evolves its slice of the chains (cpnest_b200/sharded.py)
    N, K = args.nlive * world, args.k * world
    bounds = [[-10.0, 10.0]] * DIM
    np.random.seed(SEED)
    cycle = DefaultProposalCycle().device_cycle()        # drawn like the reference's (proposal.py:71-75)

    # ------------------------------------------------------------------------------------------
    # device-resident run: `value`
    # ------------------------------------------------------------------------------------------
    eng = Engine("gaussian_iid", bounds, nlive=N, maxmcmc=MAXMCMC, seed=SEED, device=local_rank, lanes=args.lanes,
                 rank=rank, world_size=world)
    eng.set_stream(stream.cuda_stream)
    eng.set_cycle(cycle)
    eng.init_prior()                                     # same seed everywhere: identical replicas, no broadcast
    if world > 1:
        from cpnest_b200.sharded import ShardedEngine
        drv = ShardedEngine(eng, comm=args.comm)
    else:
        drv = eng
    # A run has about nlive * (H + 3 sqrt(H)) / K batches (H = 79 nats): ~190 at the default sizes.  Steps beyond the end
    # of a run would be no-ops, so the state is polled every CHECK steps and a finished run is followed by a fresh one
    # (new seed); the re-initialisation (prior draw + sort, a few ms) stays inside the timed region.
    CHECK = 16
    runs = [0]

    carried = dict(total_evals=0, total_steps=0, total_accepted=0)

    def fresh_run():
        st = eng.state()
        for k in carried:
            carried[k] += st[k]                   # cpn_init_prior resets the run's counters
        runs[0] += 1
        eng.set_seed(SEED + 7919 * runs[0])
        eng.init_prior()

    def totals():
        st = eng.state()
        return {k: carried[k] + st[k] for k in carried}, st

    def do_steps(n):
        done = 0
        while done < n:
            m = min(CHECK, n - done)
            for _ in range(m):
                drv.ns_step(K)
            done += m
            st = eng.state()
            if st["done"] or not (st["condition"] > 0.1):
                fresh_run()

    do_steps(args.warmup)
    torch.cuda.synchronize()
    s0, _ = totals()
    r0 = runs[0]
    l0 = eng.launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    do_steps(args.steps)
    ev1.record(stream)
    barrier()
    dt_ms = max_over_ranks(ev0.elapsed_time(ev1))
    clk = clocks.result()
    t1, s1 = totals()
    launches = eng.launch_count() - l0
    # the counters are sums over the gathered per-chain outcomes: already global on every rank; they are cumulative
    # over the runs of this context
    evals = t1["total_evals"] - s0["total_evals"]
    steps_mcmc = t1["total_steps"] - s0["total_steps"]
    accepted = t1["total_accepted"] - s0["total_accepted"]
    value = evals / (dt_ms * 1e-3)
    restarts = runs[0] - r0
    if s1["done"] or not (s1["condition"] > 0.1) or s1["batches"] < 12:
        fresh_run()
        do_steps(30)

    # ---- the dominant kernel alone: per-launch duration of k_mh_chains, live, with events ----
    kev = []
    mh_steps = []
    mh_evals = []
    for _ in range(min(10, args.steps)):
        eng.select_and_accumulate(K)
        a = eng.state()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.evolve_batch(K)                 # exactly one launch of k_mh_chains (this rank's chains)
        e1.record(stream)
        if world > 1:
            drv.exchange(K)
        eng.insert_batch(K)
        torch.cuda.synchronize()
        b = eng.state()
        kev.append(e0.elapsed_time(e1))
        mh_steps.append((b["total_steps"] - a["total_steps"]) / world)
        mh_evals.append((b["total_evals"] - a["total_evals"]) / world)
    k_ms = float(np.mean(kev))
    k_steps = float(np.mean(mh_steps))
    # algorithmic bytes of one launch: gathers 13.33*D bytes per MCMC step of the
    # default cycle + read and write one point (8D+16 bytes) per chain
    bytes_per_step = 8.0 * DIM * (3 / 6 + 1 / 6 + 2 / 3 + 1 / 3)
    alg_bytes = k_steps * bytes_per_step + (K / world) * 2.0 * (8 * DIM + 16)
    peak, peak_src = peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_mh_chains_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # algorithmic FP64 work of the same launch: 6.17*D flops per proposal + box test, 3*D per
    # likelihood evaluation of the iid Gaussian
    k_evals = float(np.mean(mh_evals))
    alg_flops = k_steps * 6.17 * DIM + k_evals * 3.0 * DIM
    f64_peak, f64_src = fp64_peak()
    f64_achieved = alg_flops / (k_ms * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_mh_chains", "ms_per_launch": k_ms, "mcmc_steps_per_launch": k_steps,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "fp64": {"achieved": f64_achieved, "peak": f64_peak, "unit": "TFLOP/s", "frac": f64_achieved / f64_peak,
                         "algorithmic_flops_per_launch": alg_flops, "evals_per_launch": k_evals, "peak_source": f64_src},
                "note": "gathers are served from L2 (live set %.1f MB, DRAM traffic per launch = `traffic`); the kernel is "
                        "bound by dependent-instruction latency at the occupancy K chains allow, see DESIGN.md" % (
                    N * 64 * 8 / 1e6)}

    # ---- finish the run: log-evidence wall time and validity --------------------------------------
    full = None
    if not args.no_full_run:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        drv.run(K, tolerance=0.1)
        torch.cuda.synchronize()
        t_rest = time.perf_counter() - t0
        st = eng.state()
        rect, trap = eng.finalise()
        sf = eng.state()
        sigma = math.sqrt(sf["info"] / N)
        full = {"logZ": trap, "analytic_logZ": ANALYTIC_LOGZ, "sigma": sigma, "nsigma": (trap - ANALYTIC_LOGZ) / sigma,
                "H": sf["info"], "batches": st["batches"], "dead_points": sf["iteration"], "nmcmc_final": st["nmcmc"],
                "total_evals": sf["total_evals"], "total_mcmc_steps": sf["total_steps"],
                "remaining_run_wall_s": t_rest, "rank": rank}
    if world > 1:
        drv.close()
    eng.close()

    # ------------------------------------------------------------------------------------------
    # host-buffer protocol: `e2e` (every rank serves its own K requests per step from its own host)
    # ------------------------------------------------------------------------------------------
    N, K = args.nlive, args.k
    eng = Engine("gaussian_iid", bounds, nlive=N, maxmcmc=MAXMCMC, seed=SEED + 1000 * rank + 7, device=local_rank, lanes=args.lanes)
    eng.set_stream(stream.cuda_stream)
    eng.set_cycle(cycle)
    eng.init_prior()
    live_t = torch.from_numpy(eng.live()).pin_memory()  # the caller's live set: ONE pinned host table, ascending logL at start
    live = live_t.numpy()
    rs = np.random.Generator(np.random.PCG64(SEED + rank))        # host draws of the start points (3x faster than RandomState.randint)
    st_t = torch.empty(K, dtype=torch.int32).pin_memory()
    ac_t = torch.empty(K, dtype=torch.int32).pin_memory()
    stp, acc_ = st_t.numpy(), ac_t.numpy()
    # the device pool mirrors the host live set slot by slot
    eng.set_live(live)
    nm = max(20, s1["nmcmc"])                           # the chain length the device-resident run settled on
    e2e_evals = 0

    def host_step():
        # the caller's side of the protocol, as NestedSampler.consume_sample does it (NestedSampling.py:318-366): pick the
        # K worst live points and the threshold, send K requests, put the K replies in their place.  The live set stays
        # where it is (row i <-> device pool slot i); a partial selection replaces the reference's sorted list.
        nonlocal e2e_evals
        logL = live[:, DIM]
        part = np.argpartition(logL, K - 1).astype(np.int32)   # int32: the index type of the C-ABI, no conversions later
        worst, surv = part[:K], part[K:]
        logLmin = logL[part[K - 1]]                     # the K-th smallest sits at position K-1 of a partition: max of the K worst
        starts = surv[rs.integers(0, N - K, size=K, dtype=np.int32)]     # requests: start points drawn among the survivors
        # request j = row starts[j] of the host table, reply j -> row worst[j]: the device reads and writes the pinned table
        eng.evolve_host_table(live, starts, worst, logLmin, nm, steps=stp, accepted=acc_, rows_async=True)

    for _ in range(args.warmup):
        host_step()
    barrier()
    ev_before = eng.state()["total_evals"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    eng.synchronize()                                   # the last step's rows have landed in the host table
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_evals = eng.state()["total_evals"] - ev_before
    e2e_value = sum_over_ranks(e2e_evals) / t_e2e
    h2d = K * (DIM + 2) * 8 + 2 * K * 4                 # K request rows read from the host table + start / slot indices
    d2h = K * (DIM + 2) * 8 + 2 * K * 4                 # K reply rows written into the host table + steps / accepted
    eng.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_bench
        cpu = ref_bench.time_reference(dim=DIM, budget_s=args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "steps_per_s")}

    if world > 1:
        fulls = [None] * world
        dist.all_gather_object(fulls, full)
    else:
        fulls = [full]
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "gaussian50d_iid_bounds[-10,10]_nlive%d_per_gpu_K%d_mh_default_cycle" % (N, K),
                       "multi_gpu": ("chains sharded over %d ranks, live set (nlive=%d) replicated, exchange=%s" % (world, N * world, args.comm)) if world > 1 else "single GPU",
                       "dim": DIM, "nlive_per_gpu": N, "chains_per_batch_per_gpu": K, "maxmcmc": MAXMCMC,
                       "chain_length": "adaptive (start/end correlation target 0.1)", "seed": SEED,
                       "step": "one nested-sampling batch: select+accumulate, evolve K chains, insert",
                       "l2_policy": "working set per step (live set + staging, %.1f MB) is re-gathered at random; "
                                    "no explicit flush: every step reads points rewritten by the previous one" % (
                                        (N + K) * 64 * 8 / 1e6)},
            "runs_restarted_in_timed_region": restarts,
            "mcmc_steps_per_s": steps_mcmc / (dt_ms * 1e-3),
            "accepted_chain_steps_per_s": accepted / (dt_ms * 1e-3),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "cpn_evolve_host_table (sampler pipe protocol over the caller's pinned host live table: the device reads the K request rows from it and writes the K replies into it; selection of the worst points on the host while the coordinates of the previous replies are still landing)",
                    "chain_length": nm},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "full_run": fulls,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
