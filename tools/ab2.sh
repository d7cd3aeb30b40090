#!/bin/bash
# A/B of libtsg variants at several env counts: tools/ab2.sh "4096 65536" steps lib...
ns=$1; steps=$2; shift 2
for lib in "$@"; do for n in $ns; do
  TSG_AUTORESET=0 TSG_POOL=0 TSG_LIB=$lib timeout 300 python tools/quick_bench.py $n $steps flat 0 2>&1 | tail -1 | cut -c1-190
done; done
