#!/bin/bash
# rnn_wide2 with two accumulators per group: parity + bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_sizes.py -x -q -m gpu -k "pairs or cfg5 or cfg2_full or job" 2>&1 | tail -3
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["stages_ms_sum_of_launches"])'
GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts 1024 --wave 1024 2>/dev/null | tail -1 | python -c "$summ"
} > gpurun_out/probe36.log 2>&1
echo done
