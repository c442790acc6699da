#!/bin/bash
# one GPU session: full GPU test suite, bench, ncu of the MH kernel on the bench workload, mixture throughput + ncu, mixed probe
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q > gpurun_out/t_all2.log 2>&1; echo "all rc=$?"; tail -5 gpurun_out/t_all2.log)
timeout 300 python bench.py > gpurun_out/bench_v6.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_v6.log
timeout 60 python scripts/profile_mh.py 10000 4704 44 > gpurun_out/plain_prof6.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 41 -c 1 -o gpurun_out/prof_mh_r01f python scripts/profile_mh.py 10000 4704 44 > gpurun_out/ncu_prof6.log 2>&1
tail -2 gpurun_out/ncu_prof6.log
for K in 592 1184 1776; do timeout 120 python scripts/profile_mixture.py 4000 $K 12 200; done > gpurun_out/mix_plain2.log 2>&1
cat gpurun_out/mix_plain2.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 10 -c 1 -o gpurun_out/prof_mixture_r01b python scripts/profile_mixture.py 2000 592 12 200 > gpurun_out/ncu_mix2.log 2>&1
tail -2 gpurun_out/ncu_mix2.log
timeout 200 python scripts/mixed_probe.py > gpurun_out/mixed_probe.log 2>&1; cat gpurun_out/mixed_probe.log
