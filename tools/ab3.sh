#!/bin/bash
# A/B of libtsg variants on a model: tools/ab3.sh <xml> <n> <steps> lib...
xml=$1; n=$2; steps=$3; shift 3
for lib in "$@"; do
  TSG_AUTORESET=0 TSG_POOL=0 TSG_LIB=$lib timeout 300 python tools/quick_bench.py $n $steps $xml 1 2>&1 | tail -1 | cut -c1-170
done
