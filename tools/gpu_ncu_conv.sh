#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -c 4 -o gpurun_out/prof_convtma_r02 python tools/kernel_bench.py --only "conv/stem0" --reps 1 --warm 0 > gpurun_out/ncu_convtma.log 2>&1
tail -3 gpurun_out/ncu_convtma.log
ls -la gpurun_out/prof_convtma_r02.ncu-rep
