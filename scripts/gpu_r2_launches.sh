#!/bin/bash
# launch list (device time per kernel) of a few batches at given sizes: NLIVE K BATCHES TAG
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$1; K=$2; B=$3; TAG=$4
python scripts/profile_mh.py $N $K $B > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_mh.py $N $K $B > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv 2>/dev/null | head -40
