#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q -m gpu -k "cta_pairs" 2>&1 | tail -15
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()})'
for env in "GASR_RNN_PAIR=1" "GASR_RNN_PAIR=1 GASR_RNN_G=1" "GASR_RNN_PAIR=0"; do
echo "== $env"
env $env timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
echo "== serial, pair"
GASR_RNN_PAIR=1 GASR_WAVE_SERIAL=1 timeout 600 python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
echo "== pipeline parity with pairs"
GASR_RNN_PAIR=1 timeout 600 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "cfg5_batch or job_batches" 2>&1 | tail -3
} > gpurun_out/probe12.log 2>&1
echo done
