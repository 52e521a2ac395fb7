#!/bin/bash
# Round 2, first GPU pass: full GPU test suite, smoke, both bench arms, the global-batch-32768 configuration, launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== pytest -m gpu: exit $? : $(tail -1 gpurun_out/t_all.log) [$(( $(date +%s) - t0 ))s]"
grep -E "FAILED|ERROR" gpurun_out/t_all.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke: exit $? : $(tail -1 gpurun_out/smoke.log) [$(( $(date +%s) - t0 ))s]"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "== bench reference: exit $? : $(cut -c1-200 gpurun_out/bench_ref.json) [$(( $(date +%s) - t0 ))s]"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench: exit $? [$(( $(date +%s) - t0 ))s]"; tail -3 gpurun_out/bench_n1.err; python -c "
import json;d=json.load(open('gpurun_out/bench_n1.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'], 'roof', round(d['roofline']['achieved']), round(d['roofline']['whole_step']['frac'],3), 'cpu', d['cpu_baseline']['kind'], round(d['cpu_baseline']['value'],1)); print(json.dumps(d['comparators'])[:900])"
timeout 900 python bench.py --global-batch 32768 --steps 3 --warmup 3 --no-cpu --no-comparators > gpurun_out/bench_g32k_n1.json 2> gpurun_out/bench_g32k_n1.err
echo "== bench global 32768: exit $? [$(( $(date +%s) - t0 ))s]"; tail -3 gpurun_out/bench_g32k_n1.err; python -c "
import json;d=json.load(open('gpurun_out/bench_g32k_n1.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], 'loss', d['loss'], d['loss_expected_ln_G'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu --no-comparators > gpurun_out/ncu_launch.log 2>&1
echo "== launches: exit $? [$(( $(date +%s) - t0 ))s]"
