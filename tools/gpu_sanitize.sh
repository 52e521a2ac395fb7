#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck, synccheck) over the kernel parity tests that exercise the mbarrier / TMEM /
# TMA pipelines and the P2P flag protocol.  Logs -> gpurun_out/sanitizer_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='(test_gemm and (qkv or out_proj or ragged)) or conv_tma_matches_conv2d or contrastive_lse or (test_attention) or resid_ln_equals and 257'
for tool in memcheck racecheck synccheck; do
  t0=$(date +%s)
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest tests/test_ops_gpu.py -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
  rc=$?
  echo "== $tool: exit $rc [$(( $(date +%s) - t0 ))s] $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
t0=$(date +%s)
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 99 python -m pytest tests/test_model_gpu.py -q -x -k "two_ranks_on_one_gpu or micro_batched" -p no:cacheprovider > gpurun_out/sanitizer_memcheck_p2p.log 2>&1
echo "== memcheck p2p/micro-batch: exit $? [$(( $(date +%s) - t0 ))s] $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/sanitizer_memcheck_p2p.log | tr '\n' ' ')"
