import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "gpu-accelerated-speech-recognition_b200"))
import gasr, synth
T, N, D, H, L, V, beam = 100, 2048, 161, 512, 3, 29, 16
ctx = gasr.Context(0)
pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
pipe.set_weights(*synth.rnn_weights(1, D, H, L), *synth.fc_weights(2, H, V))
x = synth.spectrogram_batch(3, T, N, D)
pipe.run_host(x); pipe.run_host(x)
print("done")
