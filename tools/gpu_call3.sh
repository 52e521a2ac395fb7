#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "conv_gemm or front_conv" > gpurun_out/t_conv.log 2>&1
echo "== conv+front: exit $? : $(tail -1 gpurun_out/t_conv.log) [$(( $(date +%s) - t0 ))s]"
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
grep -hE "^(FAILED|ERROR)|msclip:" gpurun_out/t_conv.log gpurun_out/t_model.log | sort | uniq -c | sort -rn | head -20
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/b_$name.json'));print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
}
run half512 MSCLIP_CONV_LAG=half
run deep512 MSCLIP_CONV_LAG=deep
run half1024 MSCLIP_CONV_LAG=half MSCLIP_CONV_CHUNK=1024
run deep1024 MSCLIP_CONV_LAG=deep MSCLIP_CONV_CHUNK=1024
timeout 300 python tools/kernel_bench.py --only conv/ --reps 10 > gpurun_out/kb_conv.log 2>&1; tail -6 gpurun_out/kb_conv.log
timeout 300 python tools/kernel_bench.py --only front --reps 10 > gpurun_out/kb_front.log 2>&1; tail -1 gpurun_out/kb_front.log
timeout 300 python tools/kernel_bench.py --only /attention --reps 10 > gpurun_out/kb_att.log 2>&1; tail -2 gpurun_out/kb_att.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"front_conv_kernel" -c 1 -f -o gpurun_out/prof_front_r01e \
  python tools/kernel_bench.py --only front --reps 1 --warm 0 > gpurun_out/ncu_front.log 2>&1
echo "== ncu front: exit $? [$(( $(date +%s) - t0 ))s]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention" -c 2 -f -o gpurun_out/prof_att_r01e \
  python tools/kernel_bench.py --only /attention --reps 1 --warm 0 > gpurun_out/ncu_att.log 2>&1
echo "== ncu attention: exit $? [$(( $(date +%s) - t0 ))s]"
