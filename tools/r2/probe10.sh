#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()}, d["pipeline"]["launches_per_stage_per_step"])'
for env in "GASR_WAVE_SERIAL=1" "GASR_WAVE_SERIAL=1 GASR_GEMM_BN=128" "GASR_WAVE_SERIAL=1 GASR_RNN_G=1"; do
echo "== $env (lanes 1)"
env $env timeout 600 python bench.py --steps 1 --warmup 3 --wave 2048 --lanes 1 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
} > gpurun_out/probe10.log 2>&1
echo done
