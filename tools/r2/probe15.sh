#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "lanes", d["config"]["batches_in_flight_per_gpu"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()})'
run() { echo "== $*"; env $1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks $2 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; tail -1 gpurun_out/tmp.json | python -c "$summ" || tail -5 gpurun_out/tmp.err; }
run "GASR_CTC_WARPS=7" ""
run "GASR_CTC_WARPS=0" ""
run "GASR_CTC_WARPS=0" "--wave 4096 --lanes 1"
run "GASR_CTC_WARPS=0" "--wave 4096 --lanes 2"
run "GASR_CTC_WARPS=0 GASR_CHUNK=100" ""
run "GASR_CTC_WARPS=0 GASR_CHUNK=25" ""
} > gpurun_out/probe15.log 2>&1
echo done
