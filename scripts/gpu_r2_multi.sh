#!/bin/bash
# multi-GPU: bit-identity tests + bench under torchrun at N ranks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02}
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${TAG}_t_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/${TAG}_t_multi.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ${@:3} > gpurun_out/${TAG}_bench_n$N.log 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
tail -3 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n$N.log").read().strip().splitlines()[-1])
    r = d["roofline"]; fr = d["full_run"]
    print("N=$N value %.3e e2e %.3e ms/step %.3f | kernel %.4f ms share %.3f | logz_wall %.3f s nsigma %.2f | sharded==single %s replicas %s" % (
        d["value"], d["e2e"]["value"], d["ms_per_step"], r["ms_per_launch"], r["kernel_share_of_step"], d["logz_wall_s"], fr["nsigma"],
        d.get("sharded_equals_single"), d.get("replicas_identical")))
except Exception as e:
    print("parse failed", e)
PY
