#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 300 python tools/r2/probe4.py parity
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -x -q -m gpu -k "wide or pipeline_end_to_end or rnn_recurrence" 2>&1 | tail -8
timeout 600 python tools/r2/probe4.py timing
export GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so
for g in 1 2; do
echo "== trace N=2048 G=$g"
GASR_RNN=w GASR_RNN_G=$g timeout 300 python tools/microbench.py rnn --T 100 --N 2048 --H 512 --D 512 --L 1 --iters 2 2>&1 | tail -2
done
} > gpurun_out/probe4.log 2>&1
echo done
