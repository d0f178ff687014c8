#!/bin/bash
# persistent GRU recurrence: parity tests, then cfg3 timing (config sweep of bench.py's other_configs)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "persistent_gru or cfg3" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gru" 2>&1 | tail -5
echo "== cfg3 timing: persistent vs per-step"
timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, "gpu-accelerated-speech-recognition_b200"); sys.path.insert(0, ".")
import numpy as np, gasr, synth
T, N, D, H, L, V, beam = 1000, 256, 161, 800, 5, 29, 32
w = synth.rnn_weights(7, D, H, L, cell_gates=3, bidir=True); fc = synth.fc_weights(8, 2 * H, V)
for force in ("", "t"):
    if force: os.environ["GASR_GRU"] = force
    ctx = gasr.Context(0)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_GRU, True, T, N, D, H, L, V, beam, 0, synth.VOCAB29, precision=gasr.PREC_BF16)
    pipe.set_weights(*w, *fc)
    dx = ctx.malloc(T * N * D * 4); ctx.synth_spectrogram(dx, 5, T, N, D)
    for i in range(3):
        t0 = time.time(); r = pipe.run_device(dx); dt = time.time() - t0
    print("GASR_GRU=%r" % force, "wall ms", round(dt * 1e3, 2), "stages", [round(v, 2) for v in pipe.stage_times()], r[0][0][:20], flush=True)
    if not force: ref = r
    else: print("same transcripts:", sum(a == b for a, b in zip(ref[0], r[0])), "of", N)
    pipe.close(); ctx.close()
PY
} > gpurun_out/probe30.log 2>&1
echo done
