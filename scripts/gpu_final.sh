#!/bin/bash
# round-end check on one GPU: the full GPU suite, smoke(), the bench line
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -q > gpurun_out/t_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_final.log)
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.log").read().strip().splitlines()[-1])
print("value %.4e ms/step %.4f e2e %.4e kernel ms %.4f nsigma %.2f cpu %.4e launches %d frac %.3f fp64 frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["full_run"][0]["nsigma"], d["cpu_baseline"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"]))
PY
