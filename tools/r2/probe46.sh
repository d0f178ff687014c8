#!/bin/bash
# full GPU suite, default bench line (with the stages-alone pass), round-2 ncu captures
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/probe46_tests.log 2>&1
tail -6 gpurun_out/probe46_tests.log
GASR_WAVE_TIMEOUT_S=30 timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null; echo "ref rc=$?"
bash tools/r2/capture_profiles.sh
