#!/bin/bash
# checkpoint: the full GPU suite + the default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/probe40_tests.log 2>&1
tail -6 gpurun_out/probe40_tests.log
GASR_WAVE_TIMEOUT_S=30 timeout 900 python bench.py > gpurun_out/bench_ckpt.json 2> gpurun_out/bench_ckpt.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_ckpt.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
