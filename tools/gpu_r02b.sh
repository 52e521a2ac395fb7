#!/bin/bash
# conv path A/B: per-layer kernel timing (pair / single CTA), whole-step bench with the TMA feed vs the gather kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py --only conv/ --reps 20 > gpurun_out/kb_conv_tma_pair.log 2>&1; cat gpurun_out/kb_conv_tma_pair.log | grep conv/
echo "--- MSCLIP_CONV_PAIR=0"
MSCLIP_CONV_PAIR=0 timeout 300 python tools/kernel_bench.py --only conv/ --reps 20 2>&1 | grep conv/ | tee gpurun_out/kb_conv_tma_single.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -3
for mode in tma gather; do
  if [ $mode = gather ]; then export MSCLIP_CONV_GATHER=1; fi
  timeout 600 python bench.py --no-cpu --no-comparators --no-e2e > gpurun_out/b_conv_$mode.json 2> gpurun_out/b_conv_$mode.err
  python -c "
import json;d=json.load(open('gpurun_out/b_conv_$mode.json'));print('$mode', round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks'])"
done
unset MSCLIP_CONV_GATHER
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu --no-comparators > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02b.csv | tail -12
