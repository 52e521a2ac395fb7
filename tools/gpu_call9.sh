#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
run() { # name, args..., env via VAR=...
  local name=$1; shift
  env "$@" > /dev/null 2>&1
}
bench() { # name env... -- args...
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python bench.py "$@" > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/b_$name.json'));print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'], d['loss'])" 2>&1) [$(( $(date +%s) - t0 ))s]"
}
bench fold MSCLIP_LN_FOLD=1 -- --steps 10 --warmup 3 --no-e2e --no-cpu
bench nofold MSCLIP_LN_FOLD=0 -- --steps 10 --warmup 3 --no-e2e --no-cpu
bench fold2 MSCLIP_LN_FOLD=1 -- --steps 10 --warmup 3 --no-e2e --no-cpu
bench nofold2 MSCLIP_LN_FOLD=0 -- --steps 10 --warmup 3 --no-e2e --no-cpu
bench b16 X=1 -- --steps 5 --warmup 3 --no-e2e --no-cpu --patch 16
timeout 300 python tools/zeroshot_bench.py > gpurun_out/zeroshot.log 2>&1; tail -3 gpurun_out/zeroshot.log
echo "[$(( $(date +%s) - t0 ))s]"
