#!/bin/bash
# round 2, call 1: new production parity tests, whole GPU suite, bench at lanes 16 / 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_production_parity.py -x -q -m gpu > gpurun_out/r2_t_prod.log 2>&1; echo "prod parity rc=$?" 
python -m pytest tests -x -q -m gpu > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?"
tail -3 gpurun_out/r2_t_prod.log; tail -3 gpurun_out/r2_t_all.log
for L in 16 8; do
  python bench.py --lanes $L --steps 120 --warmup 30 --no-cpu-baseline > gpurun_out/r2_bench_l$L.log 2>&1; echo "bench lanes $L rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_l$L.log").read().strip().splitlines()[-1])
    print("lanes $L value %.3e e2e %.3e ms/step %.3f kernel ms %.3f steps/launch %.3e fp64 %.3f nsigma %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["mcmc_steps_per_launch"], d["roofline"]["fp64"]["frac"], d["full_run"][0]["nsigma"]))
except Exception as e:
    print("parse failed", e)
PY
done
