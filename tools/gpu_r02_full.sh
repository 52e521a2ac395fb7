#!/bin/bash
# Round-2 verification on one B200: GPU test suite, smoke, both bench arms, metric configuration (global batch 32768), B/16,
# per-kernel timing, launch list, ncu --set full captures of the attention / conv / MLP / similarity kernels, timeline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== pytest -m gpu: exit $? : $(tail -1 gpurun_out/t_all.log) [$(( $(date +%s) - t0 ))s]"; grep -E "^FAILED|^ERROR" gpurun_out/t_all.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke: exit $? : $(tail -1 gpurun_out/smoke.log) [$(( $(date +%s) - t0 ))s]"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "== bench reference: exit $? : $(cut -c1-220 gpurun_out/bench_ref.json) [$(( $(date +%s) - t0 ))s]"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench: exit $? [$(( $(date +%s) - t0 ))s]"; tail -2 gpurun_out/bench_n1.err; python -c "
import json;d=json.load(open('gpurun_out/bench_n1.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'], 'roof', round(d['roofline']['achieved']), round(d['roofline']['whole_step']['frac'],3), 'cpu', d['cpu_baseline']['kind'], round(d['cpu_baseline']['value'],1)); print({k:round(v['value']) for k,v in d['comparators']['reference_eager_b200'].items() if isinstance(v,dict)})"
timeout 900 python bench.py --global-batch 32768 --steps 3 --warmup 3 --no-cpu --no-comparators > gpurun_out/bench_g32k_n1.json 2> gpurun_out/bench_g32k_n1.err
echo "== bench global 32768: exit $? [$(( $(date +%s) - t0 ))s]"; python -c "
import json;d=json.load(open('gpurun_out/bench_g32k_n1.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], 'loss', d['loss'], d['loss_expected_ln_G'])"
timeout 900 python bench.py --patch 16 --no-cpu --no-comparators --steps 5 > gpurun_out/bench_b16_n1.json 2> gpurun_out/bench_b16_n1.err
echo "== bench B/16: exit $? [$(( $(date +%s) - t0 ))s]"; python -c "
import json;d=json.load(open('gpurun_out/bench_b16_n1.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['roofline']['whole_step'])"
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "== kernel_bench exit $? [$(( $(date +%s) - t0 ))s]"; grep -E "attention|layernorm|conv/|front|out_proj .*pair=1|fc1 .*pair=1|fc2 .*pair=1|qkv .*pair=1" gpurun_out/kernel_bench.log
python tools/timeline.py --steps 3 > gpurun_out/timeline_r02.md 2>/dev/null; sed -n 3,8p gpurun_out/timeline_r02.md
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu --no-comparators > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02.csv > gpurun_out/launches_r02_summary.md; tail -10 gpurun_out/launches_r02_summary.md
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention -c 2 -o gpurun_out/prof_att_r02 python tools/kernel_bench.py --only attention --reps 1 --warm 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -c 1 -o gpurun_out/prof_fc1_r02 python tools/kernel_bench.py --only text/fc1 --modes 1 --reps 1 --warm 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contrastive_lse -c 1 -o gpurun_out/prof_loss_r02 python -m pytest tests/test_ops_gpu.py -q -k "contrastive_lse and 4096" > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -c 12 -o gpurun_out/prof_conv_r02 python tools/kernel_bench.py --only conv/ --reps 1 --warm 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
echo "== done [$(( $(date +%s) - t0 ))s]"
