#!/bin/bash
# persistent GRU recurrence: row blocks per CTA (1 / 2) x k-block rotation (0 / 1): cfg3 timing + per-role cycles
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cat > /tmp/cfg3_time.py <<'PY'
import os, sys, time
sys.path.insert(0, "gpu-accelerated-speech-recognition_b200"); sys.path.insert(0, ".")
import numpy as np, gasr, synth
T, N, D, H, L, V, beam = 1000, 256, 161, 800, 5, 29, 32
w = synth.rnn_weights(7, D, H, L, cell_gates=3, bidir=True); fc = synth.fc_weights(8, 2 * H, V)
ctx = gasr.Context(0)
pipe = gasr.AsrPipeline(ctx, gasr.CELL_GRU, True, T, N, D, H, L, V, beam, 0, synth.VOCAB29, precision=gasr.PREC_BF16)
pipe.set_weights(*w, *fc)
dx = ctx.malloc(T * N * D * 4); ctx.synth_spectrogram(dx, 5, T, N, D)
for i in range(int(sys.argv[1])):
    t0 = time.time(); r = pipe.run_device(dx); dt = time.time() - t0
print("PP=%s ROT=%s" % (os.environ.get("GASR_GRU_PP", "1"), os.environ.get("GASR_GRU_ROT", "1")), "wall ms", round(dt * 1e3, 2), "stages", [round(v, 2) for v in pipe.stage_times()], flush=True)
PY
{
timeout 300 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "persistent_gru" 2>&1 | tail -3
for pp in 1 2; do for rot in 0 1; do
GASR_GRU_PP=$pp GASR_GRU_ROT=$rot timeout 200 python /tmp/cfg3_time.py 3
GASR_GRU_PP=$pp GASR_GRU_ROT=$rot GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so GASR_GS_TRACE=1 timeout 200 python /tmp/cfg3_time.py 1 2>&1 | grep "dir 0" | tail -1
done; done
GASR_GRU_PP=2 timeout 300 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "persistent_gru" 2>&1 | tail -3
} > gpurun_out/probe32.log 2>&1
echo done
