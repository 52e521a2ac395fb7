#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "front_conv" > gpurun_out/t_ops.log 2>&1
echo "== ops: exit $? : $(tail -1 gpurun_out/t_ops.log) [$(( $(date +%s) - t0 ))s]"
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
grep -hE "^(FAILED|ERROR)|msclip:|Error" gpurun_out/t_ops.log gpurun_out/t_model.log | sort | uniq -c | sort -rn | head
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_$i.json 2> gpurun_out/b_$i.err
echo "== bench $i: $(python -c "import json;d=json.load(open('gpurun_out/b_$i.json'));print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
done
