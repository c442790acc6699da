#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/r2_t_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_s20.log 2>&1; echo "bench s20 rc=$?"
python bench.py > gpurun_out/r2_bench_default.log 2>&1; echo "bench default rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_s20", "r2_bench_default"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % f).read().strip().splitlines()[-1])
        r = d["roofline"]; fr = d["full_run"]
        print(f, "value %.3e e2e %.3e ms/step %.3f | kernel %.4f ms %.4f ns/step fp64 %.3f | logz_wall %.3f s full evals/s %.3e nsigma %.2f (infl %.2f) failed %d | matched %.3e (%.2f ms/call) | cpu %.3e" % (
            d["value"], d["e2e"]["value"], d["ms_per_step"], r["ms_per_launch"], r["ns_per_chain_step"], r["frac"], d["logz_wall_s"], fr["evals_per_s"],
            fr["nsigma"], fr["nsigma_batch_inflated"], fr["failed_chains"], d["matched"]["value"], d["matched"]["ms_per_call"], d["cpu_baseline"]["value"]))
    except Exception as e:
        print(f, "parse failed", e)
        print(open("gpurun_out/%s.log" % f).read()[-1500:])
PY
