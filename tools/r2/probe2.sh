#!/bin/bash
# GPU probe 2 (round 2): the wide tcgen05 recurrence -- parity, then timing against the round-1 kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q -m gpu 2>&1 | tail -15
for mc in 0 1; do for g in 1 2; do
echo "== wide mc=$mc G=$g"
GASR_RNN=w GASR_RNN_MC=$mc GASR_RNN_G=$g timeout 300 python tools/microbench.py rnn --T 200 --N 2048 --H 512 --D 512 --L 1 --iters 3
done; done
echo "== wide mc=1 G=2 N=4608 (one full wave of 18 clusters)"
GASR_RNN=w timeout 300 python tools/microbench.py rnn --T 200 --N 4608 --H 512 --D 512 --L 1 --iters 3
echo "== round-1 kernel"
GASR_RNN=m timeout 300 python tools/microbench.py rnn --T 200 --N 2048 --H 512 --D 512 --L 1 --iters 2
} > gpurun_out/probe2.log 2>&1
echo done
