#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_lengths.py tests/test_gpu_sizes.py -x -q -m gpu -k "lengths or variable or job or cfg5" 2>&1 | tail -2
