#!/bin/bash
# validation after the decoder code-size change + refreshed captures
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/probe54_tests.log 2>&1
tail -6 gpurun_out/probe54_tests.log
GASR_WAVE_TIMEOUT_S=30 timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_n1.err
bash tools/r2/capture_profiles.sh
