#!/bin/bash
# Captures the round's ncu evidence on a B200 (run through gpurun; outputs under gpurun_out/).
#  * the streaming pipeline's persistent kernels wait for each other, so they cannot run under ncu's kernel
#    serialisation as a pipeline; each one is captured ALONE with its dependencies preset (GASR_STREAM_DEBUG=3)
#  * the launch list of bench.py is taken in the chunked execution mode (GASR_STREAM=0), which has per-launch dependencies
set -x
OUT=gpurun_out
export GASR_STREAM_DEBUG=3 GASR_DEBUG_REC_ALONE=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rnn_stream --launch-skip 1 -c 1 -f -o $OUT/r1_rnn_stream \
    python tools/stream_one.py 1000 64 161 512 3 16 > $OUT/r1_ncu_rnn_stream.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xproj_stream -c 1 -f -o $OUT/r1_xproj_stream \
    python tools/stream_one.py 1000 64 161 512 3 16 > $OUT/r1_ncu_xproj_stream.log 2>&1
unset GASR_STREAM_DEBUG GASR_DEBUG_REC_ALONE
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ctc_beam_cta2 -c 1 -f -o $OUT/r1_ctc_cta2 \
    python tools/microbench.py ctc --T 1000 --kind flat --iters 1 > $OUT/r1_ncu_ctc.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:ctc_beam_kernel -c 1 -f -o $OUT/r1_ctc_general \
    python tools/microbench.py ctc --T 2000 --beam 128 --kind random --iters 1 > $OUT/r1_ncu_ctc_general.log 2>&1
GASR_NO_GRAPH=1 timeout 200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:gru_tc_step --launch-skip 30 -c 1 -f \
    -o $OUT/r1_gru_step python tools/microbench.py rnn --cell gru --T 60 --N 256 --H 800 --D 161 --L 1 --iters 1 > $OUT/r1_ncu_gru_step.log 2>&1
timeout 300 python tools/config_sweep.py > $OUT/r1_config_sweep.txt 2>&1
GASR_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r1_launches_chunked.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r1_ncu_launches.log 2>&1
GASR_STREAM=0 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r1_bench_chunked.json 2>/dev/null
timeout 120 python bench.py --steps 20 --warmup 3 > $OUT/r1_bench_streaming.json 2>/dev/null
tail -c 600 $OUT/r1_bench_streaming.json
