#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 > gpurun_out/t_mgpu.log 2>&1
echo "== multigpu tests: exit $? : $(tail -1 gpurun_out/t_mgpu.log) [$(( $(date +%s) - t0 ))s]"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "== bench n2: exit $? [$(( $(date +%s) - t0 ))s]"; python -c "
import json;d=json.load(open('gpurun_out/bench_n2.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'], d['loss'], d['loss_expected_ln_G'])"
tail -3 gpurun_out/bench_n2.err
