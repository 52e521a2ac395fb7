#!/bin/bash
# round 2, backward pass (2): parity after the fp16 loss-backward operands, ncu --set full captures of the weight-gradient GEMM and
# the attention backward
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_model_gpu.py -q --timeout 300 -k "backward or training or adamw or wgrad or qgelu" 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 --launch-skip 2 -c 1 -f -o gpurun_out/prof_wgrad_fc1_r02 \
  python tools/train_bench.py --skip-step --only text/wgrad_fc1 --reps 1 > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 --launch-skip 2 -c 1 -f -o gpurun_out/prof_wgrad_qkv_r02 \
  python tools/train_bench.py --skip-step --only text/wgrad_qkv --reps 1 > gpurun_out/ncu_wgrad2.log 2>&1; tail -2 gpurun_out/ncu_wgrad2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bwd --launch-skip 2 -c 1 -f -o gpurun_out/prof_attbwd_text_r02 \
  python tools/train_bench.py --skip-step --only text/attention_bwd --reps 1 > gpurun_out/ncu_attbwd.log 2>&1; tail -2 gpurun_out/ncu_attbwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel --launch-skip 2 -c 1 -f -o gpurun_out/prof_lnbwd_text_r02 \
  python tools/train_bench.py --skip-step --only text/layernorm_bwd --reps 1 > gpurun_out/ncu_lnbwd.log 2>&1; tail -2 gpurun_out/ncu_lnbwd.log
ls -la gpurun_out/*.ncu-rep
