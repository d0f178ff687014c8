#!/bin/bash
# Round-2 ncu captures (run under gpurun, ONE GPU): launch list of a bench step + full captures of the three hot kernels.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --utts 2048 --wave 2048 --lanes 1 --steps 1 --warmup 3 --no-cpu-baseline --no-checks"
$CMD > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in rnn_wide2_kernel xproj_stream_kernel ctc_beam_warp_kernel; do
  $CMD > gpurun_out/r2_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 2 -o gpurun_out/r2_$k -f $CMD > gpurun_out/r2_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/r2_*
