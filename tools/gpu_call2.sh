#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "conv_gemm" > gpurun_out/t_conv.log 2>&1
echo "== conv: exit $? : $(tail -1 gpurun_out/t_conv.log) [$(( $(date +%s) - t0 ))s]"
MSCLIP_CONV_LAG=deep timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "conv_gemm" > gpurun_out/t_conv_deep.log 2>&1
echo "== conv deep: exit $? : $(tail -1 gpurun_out/t_conv_deep.log) [$(( $(date +%s) - t0 ))s]"
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
grep -hE "^(FAILED|ERROR)|msclip:" gpurun_out/t_conv.log gpurun_out/t_conv_deep.log gpurun_out/t_model.log | sort | uniq -c | sort -rn | head -20
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/b_$name.json'));print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
}
run half512 MSCLIP_CONV_LAG=half
run deep512 MSCLIP_CONV_LAG=deep
run half1024 MSCLIP_CONV_LAG=half MSCLIP_CONV_CHUNK=1024
run half256 MSCLIP_CONV_LAG=half MSCLIP_CONV_CHUNK=256
timeout 300 python tools/kernel_bench.py --only conv/ --reps 10 > gpurun_out/kb_conv.log 2>&1; tail -7 gpurun_out/kb_conv.log
MSCLIP_CONV_LAG=deep timeout 300 python tools/kernel_bench.py --only conv/ --reps 10 > gpurun_out/kb_conv_deep.log 2>&1; tail -7 gpurun_out/kb_conv_deep.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "== launches: exit $? [$(( $(date +%s) - t0 ))s]"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel" -c 6 -f -o gpurun_out/prof_conv_r01e \
  python tools/kernel_bench.py --only conv/ --reps 1 --warm 0 > gpurun_out/ncu_conv.log 2>&1
echo "== ncu conv: exit $? [$(( $(date +%s) - t0 ))s]"
