#!/bin/bash
# LayerNorm warps: op parity, model parity, A/B step time, kernel timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "resid_ln" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -4
for mode in 1 0; do
  MSCLIP_LN_WARPS=$mode timeout 600 python bench.py --no-cpu --no-comparators --no-e2e > gpurun_out/b_lnw_$mode.json 2> gpurun_out/b_lnw_$mode.err
  python -c "
import json;d=json.load(open('gpurun_out/b_lnw_$mode.json'));print('ln_warps=$mode', round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks'], d['loss'])"
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu --no-comparators > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02c.csv | cut -c1-120
