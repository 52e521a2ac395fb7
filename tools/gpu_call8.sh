#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 500 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "gemm" > gpurun_out/t_ops.log 2>&1
echo "== ops gemm (16): exit $? : $(tail -1 gpurun_out/t_ops.log) [$(( $(date +%s) - t0 ))s]"
MSCLIP_GEMM_EPI_WARPS=12 timeout 500 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "test_gemm" > gpurun_out/t_ops12.log 2>&1
echo "== ops gemm (12): exit $? : $(tail -1 gpurun_out/t_ops12.log) [$(( $(date +%s) - t0 ))s]"
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
grep -hE "^(FAILED|ERROR)|msclip:|Error" gpurun_out/t_ops.log gpurun_out/t_ops12.log gpurun_out/t_model.log | sort | uniq -c | sort -rn | head -20
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/b_$name.json'));print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'], round(d['roofline']['achieved']))" 2>&1) [$(( $(date +%s) - t0 ))s]"
}
run e8 MSCLIP_GEMM_EPI_WARPS=8
run e12 MSCLIP_GEMM_EPI_WARPS=12
run e16 MSCLIP_GEMM_EPI_WARPS=16
run e16fold MSCLIP_GEMM_EPI_WARPS=16 MSCLIP_LN_FOLD=1
for e in 8 12 16; do
MSCLIP_GEMM_EPI_WARPS=$e timeout 300 python tools/kernel_bench.py --only text/ --modes 1 --reps 10 > gpurun_out/kb_gemm$e.log 2>&1; echo "-- epi warps $e"; grep "pair=1" gpurun_out/kb_gemm$e.log
done
