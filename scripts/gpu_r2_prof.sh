#!/bin/bash
# ncu --set full of the production MH kernel in the bulk of the bench workload, at the lane widths given
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02a}; shift
for L in "$@"; do
  python scripts/profile_mh.py 10000 4704 40 $L > gpurun_out/${TAG}_plain_l$L.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain_l$L.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 32 -c 1 -f -o gpurun_out/${TAG}_mh_l$L \
      python scripts/profile_mh.py 10000 4704 40 $L > gpurun_out/${TAG}_ncu_l$L.log 2>&1
  echo "ncu lanes $L rc=$?"; tail -2 gpurun_out/${TAG}_ncu_l$L.log
done
