#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
echo "== decode only, W=7 via wave? no: direct ctc microbench uses the automatic choice"
timeout 120 python tools/microbench.py ctc --T 100 --N 2048 --beam 16 --iters 2 --kind flat 2>&1 | tail -3
echo "== wave pipeline T=200 N=2048 with W=7"
GASR_CTC_WARPS=7 timeout 120 python tools/r2/probe4.py timing 2>&1 | tail -8
} > gpurun_out/probe14.log 2>&1
echo done
