#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
MSCLIP_LN_FOLD=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_fold.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "== launches fold: exit $? [$(( $(date +%s) - t0 ))s]"
MSCLIP_LN_FOLD=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_nofold.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch2.log 2>&1
echo "== launches nofold: exit $? [$(( $(date +%s) - t0 ))s]"
