#!/bin/bash
# resident recurrence for wide hidden layers / small batches: parity, cfg1 timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_sizes.py tests/test_gpu_parity.py -x -q -m gpu -k "resident or cfg1 or rnn or deepspeech or cpp_module" 2>&1 | tail -4
timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, "gpu-accelerated-speech-recognition_b200"); sys.path.insert(0, ".")
import numpy as np, gasr, synth
T, N, D, H, L, V, beam = 200, 1, 2048, 2048, 1, 47, 100
vocab = bytes(range(1, V + 1))
x = synth.spectrogram_batch(101, T, N, D)
w = synth.rnn_weights(102, D, H, L); fc = synth.fc_weights(103, H, V)
for force in ("", "f"):
    if force: os.environ["GASR_RNN"] = force
    ctx = gasr.Context(0)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, vocab)
    pipe.set_weights(*w, *fc)
    for i in range(3):
        t0 = time.time(); r = pipe.run_host(x); dt = time.time() - t0
    print("GASR_RNN=%r" % force, "wall ms %.2f" % (dt * 1e3), "stages", [round(v, 3) for v in pipe.stage_times()], r[0][0][:16], flush=True)
    pipe.close(); ctx.close()
PY
} > gpurun_out/probe51.log 2>&1
echo done
