#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/ln_probe.py --reps 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -c 2 -o gpurun_out/prof_lnw_r02 python tools/ln_probe.py --reps 1 --only text/out_proj > gpurun_out/ncu_lnw.log 2>&1
tail -2 gpurun_out/ncu_lnw.log
