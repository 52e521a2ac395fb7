#!/bin/bash
# Round-end style verification on one B200: full GPU test suite, smoke, bench (both arms), ncu launch list.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_final.sh'        (artefacts land in gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "== pytest -m gpu: exit $? : $(tail -1 gpurun_out/t_all.log) [$(( $(date +%s) - t0 ))s]"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke: exit $? : $(tail -1 gpurun_out/smoke.log) [$(( $(date +%s) - t0 ))s]"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "== bench reference: exit $? : $(cut -c1-160 gpurun_out/bench_ref.json) [$(( $(date +%s) - t0 ))s]"
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench: exit $? [$(( $(date +%s) - t0 ))s]"; python -c "
import json;d=json.load(open('gpurun_out/bench_n1.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'], 'roof', round(d['roofline']['achieved']), round(d['roofline']['whole_step']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],1))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "== launches: exit $? [$(( $(date +%s) - t0 ))s]"
