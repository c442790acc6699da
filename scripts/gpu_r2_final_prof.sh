#!/bin/bash
# final profiles of the round: launch list of the bench command, ncu --set full of the hot kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02f}
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_ncu.log 2>&1
echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches_bench.csv | head -30
python scripts/profile_mh.py 10000 4704 80 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 70 -c 1 -f -o gpurun_out/${TAG}_mh python scripts/profile_mh.py 10000 4704 80 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu full rc=$?"; tail -1 gpurun_out/${TAG}_plain.log
