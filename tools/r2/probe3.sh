#!/bin/bash
# GPU probe 3 (round 2): per-step cycle breakdown of the wide recurrence (instrumented build), parity of the new release pattern
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q -m gpu 2>&1 | tail -5
export GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so
for n in 256 2048 4608; do for mc in 0 1; do for g in 1 2; do
echo "== trace N=$n mc=$mc G=$g"
GASR_RNN=w GASR_RNN_MC=$mc GASR_RNN_G=$g timeout 300 python tools/microbench.py rnn --T 100 --N $n --H 512 --D 512 --L 1 --iters 2 2>&1 | tail -3
done; done; done
unset GASR_LIB
echo "== product build timing"
for mc in 0 1; do for g in 1 2; do
GASR_RNN=w GASR_RNN_MC=$mc GASR_RNN_G=$g timeout 300 python tools/microbench.py rnn --T 200 --N 2048 --H 512 --D 512 --L 1 --iters 3 | tail -1
done; done
} > gpurun_out/probe3.log 2>&1
echo done
