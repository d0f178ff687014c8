#!/bin/bash
# final validation: smoke, full GPU suite, default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/probe58_tests.log 2>&1
tail -5 gpurun_out/probe58_tests.log
GASR_WAVE_TIMEOUT_S=30 timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_n1.err
