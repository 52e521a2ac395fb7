#!/bin/bash
# A/B: one-MUFU QuickGELU (libmsclip_b200_tanh.so): parity, fc1 kernel time, step time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sfx in "" "_tanh"; do
  export MSCLIP_LIB_SUFFIX=$sfx
  echo "=== variant '$sfx'"
  timeout 600 python -m pytest tests/test_model_gpu.py -q -k "golden or fresh" 2>&1 | tail -2
  cp gpurun_out/parity_model.json gpurun_out/parity_model$sfx.json
  python tools/kernel_bench.py --only fc1 --reps 20 --modes 1 2>&1 | grep fc1
  timeout 600 python bench.py --no-cpu --no-comparators --no-e2e > gpurun_out/b_qgelu$sfx.json 2> gpurun_out/b_qgelu$sfx.err
  python -c "
import json;d=json.load(open('gpurun_out/b_qgelu$sfx.json'));print(round(d['value']), round(d['ms_per_step'],2), d['clocks'], 'roof', round(d['roofline']['achieved']))"
done
python - <<'PY'
import json
a=json.load(open('gpurun_out/parity_model.json')); b=json.load(open('gpurun_out/parity_model_tanh.json'))
for k in sorted(a):
    if 'ours_vs_fp32' in a[k] and k in b:
        oa,ob=a[k]['ours_vs_fp32'],b[k]['ours_vs_fp32']; ac=a[k].get('reference_autocast_vs_fp32',{})
        print(k, 'logits %.3e -> %.3e (autocast %.3e)  img %.3e -> %.3e  loss %.2e -> %.2e'%(oa['logits'],ob['logits'],ac.get('logits',0),oa['image_features'],ob['image_features'],oa['loss_fused_kernel'],ob['loss_fused_kernel']))
PY
