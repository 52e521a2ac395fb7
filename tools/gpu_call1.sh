#!/bin/bash
# dev call: validate the fused front kernel, A/B the bench, chunk sweep, ncu of the conv kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
t0=$(date +%s)
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 150 -k "front_conv" > gpurun_out/t_front.log 2>&1
echo "== front: exit $? : $(tail -1 gpurun_out/t_front.log) [$(( $(date +%s) - t0 ))s]"
timeout 700 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/t_model.log) [$(( $(date +%s) - t0 ))s]"
grep -hE "^(FAILED|ERROR)|msclip:" gpurun_out/t_front.log gpurun_out/t_model.log | sort | uniq -c | sort -rn | head -20
MSCLIP_FRONT_FUSED=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_unfused.json 2> gpurun_out/b_unfused.err
echo "== unfused: $(python -c "import json;d=json.load(open('gpurun_out/b_unfused.json'));print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_fused.json 2> gpurun_out/b_fused.err
echo "== fused: $(python -c "import json;d=json.load(open('gpurun_out/b_fused.json'));print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
for c in 128 256 512 2048; do
  MSCLIP_CONV_CHUNK=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_chunk$c.json 2> gpurun_out/b_chunk$c.err
  echo "== chunk $c: $(python -c "import json;d=json.load(open('gpurun_out/b_chunk$c.json'));print(d['value'], d['ms_per_step'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
done
timeout 300 python tools/kernel_bench.py --only front --reps 10 > gpurun_out/kb_front.log 2>&1; tail -2 gpurun_out/kb_front.log
timeout 300 python tools/kernel_bench.py --only conv/ --reps 10 > gpurun_out/kb_conv.log 2>&1; tail -7 gpurun_out/kb_conv.log
echo "[$(( $(date +%s) - t0 ))s]"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel" -c 6 -f -o gpurun_out/prof_conv_r01d \
  python tools/kernel_bench.py --only conv/ --reps 1 --warm 0 > gpurun_out/ncu_conv.log 2>&1
echo "== ncu conv: exit $? [$(( $(date +%s) - t0 ))s]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"front_conv_kernel" -c 1 -f -o gpurun_out/prof_front_r01d \
  python tools/kernel_bench.py --only front --reps 1 --warm 0 > gpurun_out/ncu_front.log 2>&1
echo "== ncu front: exit $? [$(( $(date +%s) - t0 ))s]"
ls -la gpurun_out | head -40
