#!/bin/bash
# hardware work queues: 2 lanes x 9 streams on the default 8 connections alias; CUDA_DEVICE_MAX_CONNECTIONS=32
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]))'
for mc in 8 32; do
echo "== CUDA_DEVICE_MAX_CONNECTIONS=$mc"
for a in "1024 512" "1024 1024" "2048 1024" "2048 2048" "8192 4096" "8192 2048"; do set -- $a
CUDA_DEVICE_MAX_CONNECTIONS=$mc GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts $1 --wave $2 2>/dev/null | tail -1 | python -c "$summ"
done; done
} > gpurun_out/probe39.log 2>&1
echo done
