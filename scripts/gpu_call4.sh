#!/bin/bash
# one GPU session: full GPU test suite, phase times, bench, ncu launch list of the bench command, BASELINE configurations
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/t_all4.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/t_all4.log)
timeout 120 python scripts/phase_times.py 4704 0 10000 60 60 2>&1 | head -3
timeout 300 python bench.py > gpurun_out/bench_v8.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_v8.log").read().strip().splitlines()[-1])
print("value %.4e ms/step %.4f e2e %.4e kernel ms %.4f nsigma %.2f cpu %.4e launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["full_run"][0]["nsigma"], d["cpu_baseline"]["value"], d["gpu_launches"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01h_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench8.log 2>&1; echo "ncu launches rc=$?"
timeout 200 python scripts/baseline_configs.py 1 2 3 4 5 6 > gpurun_out/baseline_configs_v8.jsonl 2> gpurun_out/baseline_configs_v8.err; echo "configs rc=$?"; cut -c1-200 gpurun_out/baseline_configs_v8.jsonl
