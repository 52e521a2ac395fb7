#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "attention" > gpurun_out/t_ops.log 2>&1
echo "== ops: exit $? : $(tail -1 gpurun_out/t_ops.log) [$(( $(date +%s) - t0 ))s]"
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
grep -hE "^(FAILED|ERROR)|msclip:" gpurun_out/t_ops.log gpurun_out/t_model.log | sort | uniq -c | sort -rn | head -20
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/b_$name.json'));print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
}
run c512 MSCLIP_CONV_CHUNK=512
run c1024 MSCLIP_CONV_CHUNK=1024
timeout 300 python tools/kernel_bench.py --only /attention --reps 10 > gpurun_out/kb_att.log 2>&1; tail -2 gpurun_out/kb_att.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention" -c 2 -f -o gpurun_out/prof_att_r01f \
  python tools/kernel_bench.py --only /attention --reps 1 --warm 0 > gpurun_out/ncu_att.log 2>&1
echo "== ncu attention: exit $? [$(( $(date +%s) - t0 ))s]"
