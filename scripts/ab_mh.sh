#!/bin/bash
# A/B of two builds of the library on the MH kernel: identical chains (hash of a short run) and ns per chain-step.
# usage: scripts/ab_mh.sh OLD_LIB_DIR  (compares cpnest_b200/lib_exp/OLD_LIB_DIR with the in-tree build)
OLD=cpnest_b200/lib_exp/$1/libcpnest_b200.so
mkdir -p gpurun_out
{
echo "== determinism / equality of chains: old"; CPNEST_B200_LIB=$OLD python scripts/determinism.py
echo "== determinism / equality of chains: new"; python scripts/determinism.py
for cfg in ${CFGS:-"4704 16" "4704 8" "4704 32" "2352 16" "37632 8 10 10 100000" "37632 16 10 10 100000"}; do
  echo "== time_mh $cfg: old"; CPNEST_B200_LIB=$OLD python scripts/time_mh.py $cfg
  echo "== time_mh $cfg: new"; python scripts/time_mh.py $cfg
done
} > gpurun_out/ab_mh_$1.log 2>&1
tail -40 gpurun_out/ab_mh_$1.log
