#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench line, ncu launch list of a short bench run, ncu --set full of one
# steady-state launch of the step kernel at the benched env count.
# usage: tools/gpu_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
S="--envs 16384 --steps 2 --warmup 3 --settle 20 --no-cpu-baseline --no-e2e --no-workloads --sweep ''"
timeout 600 python bench.py --envs 16384 --steps 2 --warmup 3 --settle 20 --no-cpu-baseline --no-e2e --no-workloads --sweep "" > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --envs 16384 --steps 2 --warmup 3 --settle 20 --no-cpu-baseline --no-e2e --no-workloads --sweep "" > gpurun_out/${tag}_ncu_l.log 2>&1
export TSG_AUTORESET=0 TSG_POOL=0 TSG_SETTLE=200
timeout 300 python tools/quick_bench.py 131072 2 flat 0 > gpurun_out/${tag}_qb.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tb_env_kernel -s 203 -c 1 \
  -o gpurun_out/${tag}_prof -f python tools/quick_bench.py 131072 2 flat 0 > gpurun_out/${tag}_ncu.log 2>&1
echo done
