#!/bin/bash
# same-box A/B: warp decoder with the 32-iteration shuffle loops unrolled by 8 (pkgB: 5200 SASS instructions) vs fully (7008)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], d["stages_ms_sum_of_launches"]["ctc_decode"])'
for rep in 1 2 3; do
echo -n "unroll 8 (A): "; GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
echo -n "variant B   : "; GASR_LIB=$PWD/tools/r2/pkgB/libgasr.so GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
done
echo "decoder alone: A / B"
summ2='import json,sys
d=json.loads(sys.stdin.read()); print(d["stages_ms_sum_of_launches"])'
GASR_WAVE_SERIAL=1 GASR_WAVE_TIMEOUT_S=30 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-checks --utts 4096 --wave 4096 --lanes 1 2>/dev/null | tail -1 | python -c "$summ2"
GASR_LIB=$PWD/tools/r2/pkgB/libgasr.so GASR_WAVE_SERIAL=1 GASR_WAVE_TIMEOUT_S=30 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-checks --utts 4096 --wave 4096 --lanes 1 2>/dev/null | tail -1 | python -c "$summ2"
} > gpurun_out/probe53.log 2>&1
echo done
