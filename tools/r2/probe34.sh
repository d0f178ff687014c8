#!/bin/bash
# persistent GRU recurrence: units per CTA x row blocks per CTA: parity, cfg3 timing, per-role cycles
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp tools/r2/cfg3_time.py /tmp/cfg3_time.py
{
for cfg in "0 0" "32 1" "32 2"; do
set -- $cfg
echo "== units $1 pp $2"
GASR_GRU_UNITS=$1 GASR_GRU_PP=$2 timeout 300 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "persistent_gru" 2>&1 | tail -2
GASR_GRU_UNITS=$1 GASR_GRU_PP=$2 timeout 200 python /tmp/cfg3_time.py 3
GASR_GRU_UNITS=$1 GASR_GRU_PP=$2 GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so GASR_GS_TRACE=1 timeout 200 python /tmp/cfg3_time.py 1 2>&1 | grep "dir 0" | tail -1
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gru" 2>&1 | tail -2
} > gpurun_out/probe34.log 2>&1
echo done
