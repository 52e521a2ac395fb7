#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
grep -hE "^(FAILED|ERROR)|msclip:|Error" gpurun_out/t_model.log | sort | uniq -c | sort -rn | head
timeout 400 python tools/zeroshot_bench.py > gpurun_out/zeroshot.log 2>&1; tail -2 gpurun_out/zeroshot.log
echo "[$(( $(date +%s) - t0 ))s]"
