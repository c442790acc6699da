#!/bin/bash
# one point of the scaling curves on an N-GPU box: scripts/gpu_r2g_scale.sh N [weak|strong|tests ...]
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$1; shift
TAG=r02g
run() {  # name args...
  local name=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n$N.log 2> gpurun_out/${TAG}_${name}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n$N.log 2> gpurun_out/${TAG}_${name}_n$N.err
  fi
  echo "$name N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${name}_n$N.log").read().strip().splitlines()[-1])
    fr = d.get("full_run") or d.get("sampler_run")
    print("   value %.4e ms/step %s logz_wall %.3f s nsigma %s e2e %s eq %s kernel_share %s" % (d["value"], d.get("ms_per_step"), d["logz_wall_s"], fr and fr.get("nsigma"),
          (d.get("e2e") or {}).get("value"), d.get("sharded_equals_single"), (d.get("roofline") or {}).get("kernel_share_of_step")))
except Exception as e:
    print("   parse failed", e)
PY
}
for what in "$@"; do
  case $what in
    tests) python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${TAG}_multi_tests_n$N.log 2>&1; echo "multi tests rc=$?"; tail -2 gpurun_out/${TAG}_multi_tests_n$N.log;;
    weak) run weak --steps 20 --warmup 5 --no-cpu-baseline;;
    strong) run strong --steps 20 --warmup 5 --scaling strong --nlive 100000 --k 37632 --no-cpu-baseline;;
    corr50) run corr50 --workload corr50;;
    rosen10|gauss20|mix20) run $what --workload $what;;
  esac
done
