#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of the step kernel + SASS/source pages.
# usage: tools/gpu_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --envs 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --sweep "" > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --envs 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --sweep "" > gpurun_out/${tag}_ncu_l.log 2>&1
TSG_AUTORESET=0 TSG_POOL=0 timeout 300 python tools/quick_bench.py 16384 2 flat 0 > gpurun_out/${tag}_qb.log 2>&1 && \
TSG_AUTORESET=0 TSG_POOL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsg_env_kernel -s 3 -c 1 \
  -o gpurun_out/${tag}_prof -f python tools/quick_bench.py 16384 2 flat 0 > gpurun_out/${tag}_ncu.log 2>&1
echo done
