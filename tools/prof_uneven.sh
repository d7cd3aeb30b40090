#!/bin/bash
tag=${1:-uneven}
TSG_AUTORESET=0 TSG_POOL=0 timeout 300 python tools/quick_bench.py 16384 2 uneven 1 > gpurun_out/${tag}_qb.log 2>&1; cat gpurun_out/${tag}_qb.log | cut -c1-200
TSG_AUTORESET=0 TSG_POOL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tb_env_kernel -s 3 -c 1 \
  -o gpurun_out/${tag}_prof -f python tools/quick_bench.py 16384 2 uneven 0 > gpurun_out/${tag}_ncu.log 2>&1
echo done
