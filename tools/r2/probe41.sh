#!/bin/bash
# decoder CTA size next to the CTA-pair GEMM (148 KB of shared memory per GEMM CTA: one 8-warp decoder CTA fits beside it, or three 4-warp ones)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], d["stages_ms_sum_of_launches"])'
for w in 8 4 6 2 8; do
echo -n "GASR_CTC_WARPS=$w: "
GASR_CTC_WARPS=$w GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
done
} > gpurun_out/probe41.log 2>&1
echo done
