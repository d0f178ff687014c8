import os, sys, time
sys.path.insert(0, "gpu-accelerated-speech-recognition_b200"); sys.path.insert(0, ".")
import numpy as np, gasr, synth
T, N, D, H, L, V, beam = (int(sys.argv[2]) if len(sys.argv) > 2 else 1000), 256, 161, 800, 5, 29, 32
w = synth.rnn_weights(7, D, H, L, cell_gates=3, bidir=True); fc = synth.fc_weights(8, 2 * H, V)
ctx = gasr.Context(0)
pipe = gasr.AsrPipeline(ctx, gasr.CELL_GRU, True, T, N, D, H, L, V, beam, 0, synth.VOCAB29, precision=gasr.PREC_BF16)
pipe.set_weights(*w, *fc)
dx = ctx.malloc(T * N * D * 4); ctx.synth_spectrogram(dx, 5, T, N, D)
for i in range(int(sys.argv[1])):
    t0 = time.time(); r = pipe.run_device(dx); dt = time.time() - t0
print("PP=%s ROT=%s" % (os.environ.get("GASR_GRU_PP", "1"), os.environ.get("GASR_GRU_ROT", "1")), "wall ms", round(dt * 1e3, 2), "stages", [round(v, 2) for v in pipe.stage_times()], flush=True)
