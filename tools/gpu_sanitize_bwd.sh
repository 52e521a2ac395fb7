#!/bin/bash
# compute-sanitizer passes over the round-2 backward / input-pipeline kernels (small shapes): weight-gradient GEMM (MN-major
# TMA + tcgen05 + split reduction), attention backward, LayerNorm / QuickGELU backward, AdamW, the image transform, and one
# whole training step of a 2-layer model.  Logs -> gpurun_out/sanitizer_bwd_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='(test_wgrad and (out_proj or proj or tiny or qkv)) or (test_attention_bwd and not 64-77) or (test_layernorm_bwd and not 20000) or (test_qgelu_bwd and not 5000) or test_adamw_matches_torch'
for tool in memcheck racecheck synccheck; do
  t0=$(date +%s)
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest tests/test_backward_gpu.py -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_bwd_$tool.log 2>&1
  rc=$?
  echo "== $tool (backward ops): exit $rc [$(( $(date +%s) - t0 ))s] $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/sanitizer_bwd_$tool.log | tr '\n' ' ')"
done
t0=$(date +%s)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 99 python -m pytest tests/test_backward_gpu.py tests/test_preprocess.py -q -x -k "text_tower_backward and 2-6 or image_tower_backward and 3-4 or fixture or checkpoint_resume" -p no:cacheprovider > gpurun_out/sanitizer_bwd_memcheck_model.log 2>&1
echo "== memcheck (towers, training step, image transform): exit $? [$(( $(date +%s) - t0 ))s] $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/sanitizer_bwd_memcheck_model.log | tr '\n' ' ')"
