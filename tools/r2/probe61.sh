#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["stages_ms_sum_of_launches"])'
run() { echo -n "$1 | $2: "; env $1 GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-checks --utts 1024 $2 2>/dev/null | tail -1 | python -c "$summ"; }
run "X=1" "--wave 1024"
run "GASR_RNN_G=1" "--wave 1024"
run "GASR_RNN_G=1 GASR_CHUNK=50" "--wave 1024"
run "CUDA_DEVICE_MAX_CONNECTIONS=32" "--wave 512"
run "GASR_CHUNK=30" "--wave 1024"
run "X=1" "--wave 1024"
} > gpurun_out/probe61.log 2>&1
echo done
