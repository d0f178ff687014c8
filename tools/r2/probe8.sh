#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline_end_to_end" 2>&1 | tail -3
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()})'
for env in "GASR_GEMM_STAGES=2 GASR_CTC_WARPS=4" "GASR_GEMM_STAGES=3 GASR_CTC_WARPS=8" "GASR_GEMM_STAGES=2 GASR_CTC_WARPS=8" "GASR_GEMM_STAGES=3 GASR_CTC_WARPS=4" "GASR_GEMM_STAGES=2 GASR_CTC_WARPS=2"; do
echo "== $env"
env $env timeout 600 python bench.py --steps 2 --warmup 3 --wave 2048 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
echo "== lanes 3, wave 1024"
timeout 600 python bench.py --steps 2 --warmup 3 --wave 1024 --lanes 3 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
echo "== trace (wave engine, T=100 N=2048)"
GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so timeout 300 python tools/r2/trace_wave.py 2>&1 | tail -5
} > gpurun_out/probe8.log 2>&1
echo done
