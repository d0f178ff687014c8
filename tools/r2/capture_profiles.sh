#!/bin/bash
# Round-2 ncu captures (run under gpurun, ONE GPU): launch list of a bench step + full captures of the hot kernels.
# The wave engine has no kernel that waits for another kernel, so it runs under ncu as it is.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --utts 4096 --wave 4096 --lanes 1 --steps 1 --warmup 3 --no-cpu-baseline --no-checks"
$CMD > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in rnn_wide2_kernel gemm_pair_kernel ctc_beam_warp_kernel xproj_stream_kernel; do
  $CMD > gpurun_out/r2_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 2 -o gpurun_out/r2_$k -f $CMD > gpurun_out/r2_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
# cfg3's persistent GRU recurrence (one launch = a whole layer and direction; T = 200 keeps the replays short)
GRU="python tools/r2/cfg3_time.py 1 200"
$GRU > gpurun_out/r2_plain_gru.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gru_seq_kernel -s 2 -c 1 -o gpurun_out/r2_gru_seq_kernel -f $GRU > gpurun_out/r2_ncu_gru_seq_kernel.log 2>&1
echo "gru_seq_kernel rc=$?"
ls -la gpurun_out/r2_*
