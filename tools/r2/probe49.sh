#!/bin/bash
# small shares: the decoder stream is the slowest (1024 warps, 15 us per frame and warp); CTA-per-utterance decoder instead?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["stages_ms_sum_of_launches"])'
for a in "1024 1024" "2048 2048"; do set -- $a
for env in "X=1" "GASR_CTC_KERNEL=c" "GASR_CTC_KERNEL=c GASR_RNN_G=1" "GASR_RNN_G=1"; do
echo -n "$env: "; env $env GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts $1 --wave $2 2>/dev/null | tail -1 | python -c "$summ"
done; done
} > gpurun_out/probe49.log 2>&1
echo done
