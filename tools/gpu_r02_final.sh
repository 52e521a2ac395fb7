#!/bin/bash
# round 2, full single-GPU pass: every GPU test, smoke(), the default bench line (+ reference arm), preprocessing throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['clocks'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac')}, d['roofline']['whole_step'])
print('train', d.get('train_step'))
print('cpu', d['cpu_baseline'])
PY
timeout 600 python tools/preprocess_bench.py 2>&1 | tail -30
