#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp tools/r2/cfg3_time.py /tmp/cfg3_time.py
{
GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so GASR_GS_TRACE=1 timeout 200 python /tmp/cfg3_time.py 1 2>&1 | grep "stamps\|dir 0" | head -26
} > gpurun_out/probe35.log 2>&1
echo done
