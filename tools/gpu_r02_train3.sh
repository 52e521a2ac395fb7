#!/bin/bash
# round 2, backward pass (3): fused QuickGELU-backward epilogue + chained LayerNorm-backward emissions: parity, then step time A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py -q --timeout 300 2>&1 | tail -8
timeout 600 python tools/train_bench.py --batch 4096 --reps 5 --skip-kernels 2>&1 | tail -2 | tee gpurun_out/train_step_fused.log
MSCLIP_BWD_FUSED_GELU=0 timeout 600 python tools/train_bench.py --batch 4096 --reps 5 --skip-kernels 2>&1 | tail -2 | tee gpurun_out/train_step_unfused.log
