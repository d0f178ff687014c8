#!/bin/bash
# stream priority schemes of the wave engine (same-box A/B)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])'
for rep in 1 2; do for pm in 0 1 2 3; do
echo -n "prio $pm: "; GASR_WAVE_PRIO=$pm GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
done; done
} > gpurun_out/probe55.log 2>&1
echo done
