"""Round-2 probe: the wave engine -- parity of one batch against the oracle, then timing at growing batch sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200")); sys.path.insert(0, ROOT)
import numpy as np, gasr, synth

D, H, L, V, beam = 161, 512, 3, 29, 16
w = synth.rnn_weights(4321, D, H, L)
fc = synth.fc_weights(99, H, V)

def parity(T, N):
    from oracle import oracle as O
    x = synth.spectrogram_batch(1234, T, N, D)
    ctx = gasr.Context(0)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    paths, scores = pipe.run_host(x)
    logp = pipe.logprobs()
    ref = O.linear(O.rnn_forward(x, T, N, *w, nthreads=16)[-1], *fc, act="logsoftmax")
    err = float(np.abs(logp - ref).max())
    op, os_ = O.ctc_decode(logp.reshape(T, N, V), synth.VOCAB29, 0, beam, domain="log", nthreads=16)
    same = paths == op and all(np.float32(a) == np.float32(b) for a, b in zip(scores, os_))
    print(f"parity T={T} N={N}: mode {pipe.stage_launches()[1]}, max |logp - oracle| = {err:.2e}, decode bit-exact = {same}", flush=True)
    pipe.close(); ctx.close()

def timing(T, N):
    x = synth.spectrogram_batch(1234, T, N, D)
    ctx = gasr.Context(0)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    xd = ctx.to_device(x)
    for it in range(3):
        pipe.run_device(xd)
    ms = pipe.last_ms()
    pipe.profile(True); pipe.run_device(xd); st = pipe.stage_times(); pipe.profile(False)
    print(f"wave T={T} N={N}: {ms:.2f} ms -> RTFx {N*T*0.01/(ms*1e-3):.0f}; per 64000 frame-utts {ms*64000/(N*T):.3f} ms; "
          f"stage sums (profiled run) {['%.2f' % v for v in st]}", flush=True)
    pipe.close(); ctx.close()

if __name__ == "__main__":
    if sys.argv[1] == "parity":
        parity(120, 200); parity(200, 64)
    else:
        for T, N in ((200, 128), (200, 1024), (200, 2048), (1000, 64), (1000, 1024)):
            timing(T, N)
