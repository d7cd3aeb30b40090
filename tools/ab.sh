#!/bin/bash
# A/B timing of libtsg variants on one box: tools/ab.sh <n_envs> <steps> lib1.so lib2.so ...
n=$1; steps=$2; shift 2
for lib in "$@"; do
  for rep in 1; do
    TSG_AUTORESET=0 TSG_POOL=0 TSG_LIB=$lib timeout 300 python tools/quick_bench.py $n $steps flat $((rep==1)) 2>&1 | tail -1
  done
done
