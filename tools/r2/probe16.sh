#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py tests/test_gpu_sizes.py -x -q -m gpu -k "warp_decoder or pipeline_end_to_end or cfg5_batch or job_batches or cfg2_full or ctc" 2>&1 | tail -6
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "lanes", d["config"]["batches_in_flight_per_gpu"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()})'
run() { echo "== $*"; env $1 timeout 170 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks $2 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; tail -1 gpurun_out/tmp.json | python -c "$summ" 2>/dev/null || tail -5 gpurun_out/tmp.err; }
run "A=1" "--wave 4096 --lanes 2"
run "A=1" "--wave 4096 --lanes 1"
run "A=1" "--wave 2048 --lanes 2"
run "A=1" "--wave 2048 --lanes 3"
run "A=1" "--utts 4096 --wave 4096 --lanes 1"
run "A=1" "--utts 4096 --wave 2048 --lanes 2"
} > gpurun_out/probe16.log 2>&1
echo done
