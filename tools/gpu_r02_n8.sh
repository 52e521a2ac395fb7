#!/bin/bash
# 8-GPU pass: weak-scaling bench (global batch 32768 = BASELINE.json configs[3]) with phases + NCCL comparator, sharded parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench weak N=$N exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_n$N.json'));print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'], 'loss', d['loss'], d['loss_expected_ln_G']); print(json.dumps(d['comparators'])); print(json.dumps(d['phases']))"
tail -n 3 gpurun_out/bench_n$N.err
timeout 400 python -m pytest tests/test_multigpu.py -q -k 96 2>&1 | tail -3
