#!/bin/bash
# decoder: tie counting without the per-iteration branch; 32 vs 64 probe cells
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py tests/test_gpu_lengths.py tests/test_gpu_sizes.py -x -q -m gpu -k "ctc or decoder or cfg4 or cfg5 or lengths or pipeline" 2>&1 | tail -3
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], d["stages_ms_sum_of_launches"])'
for c in 64 32 64 32; do
echo -n "GASR_CTC_CELLS=$c: "
GASR_CTC_CELLS=$c GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
done
echo "decoder alone (wave serial)"
for c in 64 32; do
GASR_WAVE_SERIAL=1 GASR_CTC_CELLS=$c GASR_WAVE_TIMEOUT_S=30 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-checks --utts 4096 --wave 4096 --lanes 1 2>/dev/null | tail -1 | python -c "$summ"
done
} > gpurun_out/probe42.log 2>&1
echo done
