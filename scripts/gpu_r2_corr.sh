#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scripts/baseline_configs.py 7 > gpurun_out/r02_config7.jsonl 2> gpurun_out/r02_config7.err; echo "config7 rc=$?"; cut -c1-400 gpurun_out/r02_config7.jsonl
python scripts/profile_corr.py 40 > gpurun_out/r02_corr_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 32 -c 1 -f -o gpurun_out/r02_corr_mh python scripts/profile_corr.py 40 > gpurun_out/r02_corr_ncu.log 2>&1
echo "ncu corr rc=$?"; cat gpurun_out/r02_corr_plain.log | tail -1
python scripts/profile_mh.py 10000 4704 40 > gpurun_out/r02c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 32 -c 1 -f -o gpurun_out/r02c_mh python scripts/profile_mh.py 10000 4704 40 > gpurun_out/r02c_ncu.log 2>&1
echo "ncu iid rc=$?"; tail -1 gpurun_out/r02c_plain.log
