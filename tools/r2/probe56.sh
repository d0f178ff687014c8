#!/bin/bash
# more priority schemes; the winners at the smaller shares
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])'
for pm in 0 2 4 5 1 2; do
echo -n "prio $pm: "; GASR_WAVE_PRIO=$pm GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
done
for a in "1024 1024" "2048 2048" "4096 4096"; do set -- $a
for pm in 0 1 2; do
echo -n "prio $pm: "; GASR_WAVE_PRIO=$pm GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts $1 --wave $2 2>/dev/null | tail -1 | python -c "$summ"
done; done
} > gpurun_out/probe56.log 2>&1
echo done
