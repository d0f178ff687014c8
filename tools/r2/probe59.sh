#!/bin/bash
# the N = 8 share (1024 utterances on one GPU): chunk length, decoder CTA size, priorities, groups per cluster
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], d["stages_ms_sum_of_launches"])'
run() { echo -n "$1: "; env $1 GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-checks --utts 1024 --wave 1024 2>/dev/null | tail -1 | python -c "$summ"; }
run "X=1"
run "GASR_CHUNK=25"
run "GASR_CHUNK=20"
run "GASR_CHUNK=25 GASR_CTC_WARPS=4"
run "GASR_CTC_WARPS=4"
run "GASR_CHUNK=25 GASR_WAVE_PRIO=1"
run "GASR_CHUNK=25 GASR_GEMM_PAIR=0"
run "GASR_CHUNK=25 GASR_CTC_KERNEL=c"
run "X=1"
} > gpurun_out/probe59.log 2>&1
echo done
