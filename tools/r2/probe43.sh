#!/bin/bash
# same-box A/B: decoder before (pkgB) / after the tie-count change, 64 / 32 probe cells; interleaved
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], d["stages_ms_sum_of_launches"]["ctc_decode"])'
for rep in 1 2 3; do
echo -n "old         : "; GASR_LIB=$PWD/tools/r2/pkgB/libgasr.so GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
echo -n "new cells 64: "; GASR_CTC_CELLS=64 GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
echo -n "new cells 32: "; GASR_CTC_CELLS=32 GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
done
echo "decoder alone (serial, 4096 utterances): old / new 64 / new 32"
summ2='import json,sys
d=json.loads(sys.stdin.read()); print(d["stages_ms_sum_of_launches"])'
GASR_LIB=$PWD/tools/r2/pkgB/libgasr.so GASR_WAVE_SERIAL=1 GASR_WAVE_TIMEOUT_S=30 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-checks --utts 4096 --wave 4096 --lanes 1 2>/dev/null | tail -1 | python -c "$summ2"
GASR_WAVE_SERIAL=1 GASR_CTC_CELLS=64 GASR_WAVE_TIMEOUT_S=30 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-checks --utts 4096 --wave 4096 --lanes 1 2>/dev/null | tail -1 | python -c "$summ2"
GASR_WAVE_SERIAL=1 GASR_CTC_CELLS=32 GASR_WAVE_TIMEOUT_S=30 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-checks --utts 4096 --wave 4096 --lanes 1 2>/dev/null | tail -1 | python -c "$summ2"
} > gpurun_out/probe43.log 2>&1
echo done
