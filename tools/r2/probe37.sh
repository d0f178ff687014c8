#!/bin/bash
# chunk length vs batch size (strong scaling: 1024 utterances per GPU at N = 8)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3))'
for a in "1024 1024" "2048 2048" "8192 4096"; do set -- $a
for ch in 25 50 100; do
echo -n "chunk $ch: "
GASR_CHUNK=$ch GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts $1 --wave $2 2>/dev/null | tail -1 | python -c "$summ"
done; done
echo "1024 with groups per cluster = 1"
GASR_RNN_G=1 GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts 1024 --wave 1024 2>/dev/null | tail -1 | python -c "$summ"
GASR_RNN_G=1 GASR_CHUNK=25 GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts 1024 --wave 1024 2>/dev/null | tail -1 | python -c "$summ"
} > gpurun_out/probe37.log 2>&1
echo done
