"""Round-2 probe: how the round-1 kernels behave at throughput batch sizes (chunked pipeline, decoders, GEMM)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200")); sys.path.insert(0, ROOT)
import numpy as np, gasr, synth

def pipe_probe(T, N, D=161, H=512, L=3, beam=16, V=29):
    x = synth.spectrogram_batch(1234, T, N, D)
    w = synth.rnn_weights(4321, D, H, L)
    fc = synth.fc_weights(99, H, V)
    ctx = gasr.Context(0)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    xd = ctx.to_device(x)
    for it in range(3):
        ctx.sync(); t0 = time.perf_counter()
        pipe.run_device(xd)
        dt = time.perf_counter() - t0
    print(f"pipeline T={T} N={N}: {dt*1e3:.2f} ms -> RTFx {N*T*0.01/dt:.0f}; per 64000 frame-utts {dt*1e3*64000/(N*T):.3f} ms; "
          f"stages {['%.2f' % v for v in pipe.stage_times()]} mode {pipe.stage_launches()}", flush=True)
    pipe.close(); ctx.close()

if __name__ == "__main__":
    for T, N in ((200, 256), (200, 1024), (200, 2048)):
        pipe_probe(T, N)
