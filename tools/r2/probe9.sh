#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide.py -x -q -m gpu -k "pipeline_end_to_end or linear or wide" 2>&1 | tail -3
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()})'
for env in "GASR_GEMM_BN=256" "GASR_GEMM_BN=128"; do
echo "== $env"
env $env timeout 600 python bench.py --steps 2 --warmup 3 --wave 2048 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
echo "== gemm alone: rows 204800"
python tools/r2/probe4.py timing 2>&1 | grep "N=2048"
} > gpurun_out/probe9.log 2>&1
echo done
