#!/bin/bash
# Round-2 evidence pass on one B200 with the FINAL build: per-kernel timing, launch lists of one forward step and one training
# step, ncu --set full captures of the attention / MLP-GEMM / similarity / convolution kernels, bench lines for the metric
# configuration and B/16.  Everything lands in gpurun_out/ and is summarised into profiles/r02_*.md afterwards.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests/test_backward_gpu.py -q --timeout 300 2>&1 | tail -3
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "== kernel_bench exit $? [$(( $(date +%s) - t0 ))s]"
timeout 600 python tools/train_bench.py --batch 4096 --reps 5 > gpurun_out/train_bench.log 2>&1; echo "== train_bench exit $? [$(( $(date +%s) - t0 ))s]"; tail -2 gpurun_out/train_bench.log | cut -c1-600
MSCLIP_TRAIN_KEEP=0 timeout 600 python tools/train_bench.py --batch 4096 --reps 5 --skip-kernels 2>&1 | tail -2 | cut -c1-600 | tee gpurun_out/train_step_nokeep.log
timeout 600 python bench.py --global-batch 32768 --steps 3 --warmup 3 --no-cpu --no-comparators --no-train > gpurun_out/bench_g32k_n1.json 2> gpurun_out/bench_g32k_n1.err; echo "== bench global 32768 exit $? [$(( $(date +%s) - t0 ))s]"
timeout 600 python bench.py --patch 16 --no-cpu --no-comparators --no-train --steps 5 > gpurun_out/bench_b16_n1.json 2> gpurun_out/bench_b16_n1.err; echo "== bench B/16 exit $? [$(( $(date +%s) - t0 ))s]"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu --no-comparators --no-train > gpurun_out/ncu_launch.log 2>&1; echo "== forward launch list exit $? [$(( $(date +%s) - t0 ))s]"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py --batch 4096 --skip-kernels --profile-step > gpurun_out/train_ncu.log 2>&1; echo "== train launch list exit $? [$(( $(date +%s) - t0 ))s]"
cap() {  # name, kernel regex, count, command...: ncu --set full capture -> markdown table; the (large) report is not kept
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 400 ncu --set full --clock-control none -k regex:$rx -c $cnt -f -o gpurun_out/prof_${name}_r02 "$@" > gpurun_out/ncu_${name}.log 2>&1
  python tools/ncu_table.py gpurun_out/prof_${name}_r02.ncu-rep > gpurun_out/ncu_${name}_table.md 2>> gpurun_out/ncu_${name}.log
  rm -f gpurun_out/prof_${name}_r02.ncu-rep
  echo "== ncu $name: $(wc -l < gpurun_out/ncu_${name}_table.md) lines [$(( $(date +%s) - t0 ))s]"
}
cap att attention 2 python tools/kernel_bench.py --only attention --reps 1 --warm 0
cap fc1 gemm_tcgen05_kernel 1 python tools/kernel_bench.py --only text/fc1 --modes 1 --reps 1 --warm 0
cap outproj gemm_tcgen05_kernel 1 python tools/kernel_bench.py --only text/out_proj --modes 1 --reps 1 --warm 0
cap qkv gemm_tcgen05_kernel 1 python tools/kernel_bench.py --only text/qkv --modes 1 --reps 1 --warm 0
cap fc2 gemm_tcgen05_kernel 1 python tools/kernel_bench.py --only text/fc2 --modes 1 --reps 1 --warm 0
cap loss contrastive_lse 1 python -m pytest tests/test_ops_gpu.py -q -k "contrastive_lse and 4096"
cap conv gemm_tcgen05_kernel 12 python tools/kernel_bench.py --only conv/ --reps 1 --warm 0
cap front front_conv 1 python tools/kernel_bench.py --only front/fused --reps 1 --warm 0
rm -f gpurun_out/*.ncu-rep gpurun_out/*_src*.csv gpurun_out/*_raw*.csv
du -sh gpurun_out
echo "== done [$(( $(date +%s) - t0 ))s]"
