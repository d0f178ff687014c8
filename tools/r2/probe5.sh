#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -x -q -m gpu -k "wide or pipeline_end_to_end" 2>&1 | tail -4
timeout 600 python tools/r2/probe4.py timing
echo "== wave engine with the instrumented recurrence (T=100, N=2048)"
GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so timeout 300 python - <<'PY' 2>&1 | grep "rw trace" | tail -6
import sys; sys.path.insert(0, "gpu-accelerated-speech-recognition_b200")
import gasr, synth
T, N, D, H, L, V, beam = 100, 2048, 161, 512, 3, 29, 16
ctx = gasr.Context(0)
pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
pipe.set_weights(*synth.rnn_weights(1, D, H, L), *synth.fc_weights(2, H, V))
x = synth.spectrogram_batch(3, T, N, D)
pipe.run_host(x); pipe.run_host(x)
PY
} > gpurun_out/probe5.log 2>&1
python tools/microbench.py ctc --T 60 --N 2048 --beam 16 --iters 1 --kind flat > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctc_beam_warp -c 1 -o gpurun_out/ctc_warp_r2 python tools/microbench.py ctc --T 60 --N 2048 --beam 16 --iters 1 --kind flat > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?" >> gpurun_out/probe5.log
echo done
