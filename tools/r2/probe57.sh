#!/bin/bash
# with equal stream priorities: hardware queues, lanes, chunk length (same-box A/B)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "lanes", d["config"]["batches_in_flight_per_gpu"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])'
run() { echo -n "$1 | $2: "; env $1 GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks $2 2>/dev/null | tail -1 | python -c "$summ"; }
run "X=1" ""
run "CUDA_DEVICE_MAX_CONNECTIONS=32" ""
run "CUDA_DEVICE_MAX_CONNECTIONS=32" "--wave 2048 --lanes 4"
run "X=1" "--wave 2048 --lanes 4"
run "GASR_CHUNK=100" ""
run "GASR_CHUNK=25" ""
run "GASR_CTC_WARPS=4" ""
run "X=1" ""
run "GASR_WAVE_PRIO=4" ""
run "GASR_WAVE_PRIO=4" "--utts 1024 --wave 1024"
run "GASR_WAVE_PRIO=4" "--utts 2048 --wave 2048"
} > gpurun_out/probe57.log 2>&1
echo done
