#!/bin/bash
# round 2, backward pass: parity tests, kernel + whole-step timing, ncu launch list of one training step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -q --timeout 300 2>&1 | tail -15
timeout 900 python tools/train_bench.py --batch 4096 --reps 5 > gpurun_out/train_bench.log 2>&1; tail -25 gpurun_out/train_bench.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
  python tools/train_bench.py --batch 4096 --skip-kernels --profile-step > gpurun_out/train_ncu.log 2>&1; tail -3 gpurun_out/train_ncu.log
wc -l gpurun_out/train_launches.csv
