#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py tests/test_gpu_wide.py -x -q -m gpu -k "pipeline_end_to_end or cfg5_batch or job_batches or cfg2_full or cta_pairs" 2>&1 | tail -6
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()})'
for env in "GASR_GEMM_PAIR=1" "GASR_GEMM_PAIR=0" "GASR_GEMM_PAIR=1 GASR_CTC_WARPS=7"; do
echo "== $env"
env $env timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
echo "== serial"
GASR_WAVE_SERIAL=1 timeout 600 python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
} > gpurun_out/probe13.log 2>&1
echo done
