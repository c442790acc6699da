#!/bin/bash
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/t_all3.log 2>&1; echo "all rc=$?"; tail -5 gpurun_out/t_all3.log)
timeout 300 python bench.py > gpurun_out/bench_v7.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_v7.log | head -c 1200
timeout 60 python scripts/profile_mh.py 10000 4704 44 > gpurun_out/plain_prof7.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 41 -c 1 -o gpurun_out/prof_mh_r01g python scripts/profile_mh.py 10000 4704 44 > gpurun_out/ncu_prof7.log 2>&1
tail -2 gpurun_out/ncu_prof7.log
