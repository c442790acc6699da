#!/bin/bash
# scaling session on an 8-GPU box: weak curve, strong curve at nlive = 1e5 (BASELINE config 4), other samplers sharded
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02}
run() {  # n tag args...
  local n=$1 name=$2; shift 2
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n$n.log 2> gpurun_out/${TAG}_${name}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n "$@" > gpurun_out/${TAG}_${name}_n$n.log 2> gpurun_out/${TAG}_${name}_n$n.err
  fi
  echo "$name N=$n rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${name}_n$n.log").read().strip().splitlines()[-1])
    fr = d.get("full_run") or d.get("sampler_run")
    print("   value %.4e ms/step %s logz_wall %.3f s nsigma %s e2e %s eq %s" % (d["value"], d.get("ms_per_step"), d["logz_wall_s"], fr and fr.get("nsigma"),
          (d.get("e2e") or {}).get("value"), d.get("sharded_equals_single")))
except Exception as e:
    print("   parse failed", e)
PY
}
for n in 4 8; do run $n weak --steps 20 --warmup 5 --no-cpu-baseline; done
for n in 1 4 8; do run $n strong --steps 20 --warmup 5 --scaling strong --nlive 100000 --k 37632 --no-cpu-baseline; done
for w in rosen10 gauss20 mix20; do run 8 $w --workload $w; done
run 4 gauss20 --workload gauss20
