#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), d["stages_ms_sum_of_launches"], d["pipeline"]["launches_per_stage_per_step"])'
for g in 2 1; do
GASR_RNN_G=$g GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts 1024 --wave 1024 2>/dev/null | tail -1 | python -c "$summ"
GASR_WAVE_SERIAL=1 GASR_RNN_G=$g GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-checks --utts 1024 --wave 1024 2>/dev/null | tail -1 | python -c "$summ"
done
} > gpurun_out/probe38.log 2>&1
echo done
