#!/bin/bash
# Run the GPU parity tests one group per process, so a trapped kernel in one group cannot poison the
# CUDA context of the others.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in test_gemm test_conv_gemm test_layernorm test_attention test_im2col test_patch_pool test_front_conv test_adapter test_contrastive; do
  timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 200 -k "$k" > gpurun_out/ops_$k.log 2>&1
  echo "== $k: exit $? : $(tail -1 gpurun_out/ops_$k.log)"
done
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 400 > gpurun_out/model.log 2>&1
echo "== model: exit $? : $(tail -1 gpurun_out/model.log)"
grep -hE "^(FAILED|ERROR)|Error|error|msclip:" gpurun_out/ops_*.log gpurun_out/model.log | sort | uniq -c | sort -rn | head -40
