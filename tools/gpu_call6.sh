#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "gemm_ln or adapter" > gpurun_out/t_ops.log 2>&1
echo "== ops: exit $? : $(tail -1 gpurun_out/t_ops.log) [$(( $(date +%s) - t0 ))s]"
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model (fold): exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
cp gpurun_out/parity_model.json gpurun_out/parity_model_fold.json 2>/dev/null
MSCLIP_LN_FOLD=0 timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model_nofold.log 2>&1
echo "== model (no fold): exit $? : $(tail -1 gpurun_out/t_model_nofold.log) [$(( $(date +%s) - t0 ))s]"
cp gpurun_out/parity_model.json gpurun_out/parity_model_nofold.json 2>/dev/null
grep -hE "^(FAILED|ERROR)|msclip:|Error" gpurun_out/t_ops.log gpurun_out/t_model.log gpurun_out/t_model_nofold.log | sort | uniq -c | sort -rn | head -20
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/b_$name.json'));print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'], d['loss'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
}
run fold MSCLIP_LN_FOLD=1
run nofold MSCLIP_LN_FOLD=0
run fold2 MSCLIP_LN_FOLD=1
run nofold2 MSCLIP_LN_FOLD=0
