#!/bin/bash
# GPU probe 1 (round 2): round-1 kernels at throughput batch sizes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== decode throughput (warp kernel = default for N > 296)"
python tools/microbench.py ctc --T 200 --N 2048 --beam 16 --iters 3 --kind flat
python tools/microbench.py ctc --T 200 --N 4096 --beam 16 --iters 2 --kind flat
echo "== cta kernel"
GASR_CTC_KERNEL=c python tools/microbench.py ctc --T 200 --N 2048 --beam 16 --iters 3 --kind flat
echo "== cta2 kernel"
GASR_CTC_KERNEL=d python tools/microbench.py ctc --T 200 --N 2048 --beam 16 --iters 3 --kind flat
echo "== cta2 kernel MW=4"
GASR_CTC_KERNEL=d GASR_CTC_MW=4 python tools/microbench.py ctc --T 200 --N 2048 --beam 16 --iters 3 --kind flat
echo "== gemm large M"
python tools/microbench.py gemm --T 100 --N 2048 --D 512 --H 512 --iters 2
python tools/microbench.py gemm --T 100 --N 2048 --D 161 --H 512 --iters 2
echo "== linear+logsoftmax"
python tools/microbench.py linear --T 100 --N 2048 --H 512 --iters 3
echo "== chunked pipeline at large N"
GASR_STREAM=0 python tools/r2/probe1.py
} > gpurun_out/probe1.log 2>&1
echo done
