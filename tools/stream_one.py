import sys, os
sys.path.insert(0, "gpu-accelerated-speech-recognition_b200"); sys.path.insert(0, ".")
import numpy as np, gasr, synth
T,N,D,H,L,beam = [int(v) for v in sys.argv[1:7]]
V=29
x = synth.spectrogram_batch(1234, T, N, D)
w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(4321, D, H, L)
fc_w, fc_b = synth.fc_weights(99, H, V)
ctx = gasr.Context(0)
pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
pipe.set_weights(w_ih, w_hh, b_ih, b_hh, fc_w, fc_b)
if os.environ.get("GASR_STREAM_DEBUG") == "3":
    # "alone" mode (tools/capture_profiles.sh): the library runs the persistent kernels one at a time with their
    # dependencies preset and then reports an error on purpose -- there is no result to return
    try:
        pipe.run_host(x)
    except gasr.GasrError:
        pass
    sys.exit(0)
paths, scores = pipe.run_host(x)
for _ in range(3): paths, scores = pipe.run_host(x)      # warm: the printed stage times are of the last call
print("ok", paths[0][:40], scores[0], pipe.stage_times())
if len(sys.argv) > 7:
    from oracle import oracle as O
    logp = pipe.logprobs()
    ref_logp = O.linear(O.rnn_forward(x, T, N, w_ih, w_hh, b_ih, b_hh, nthreads=8)[-1], fc_w, fc_b, act="logsoftmax")
    print("max err", np.abs(logp - ref_logp).max())
    op, os_ = O.ctc_decode(logp.reshape(T, N, V), synth.VOCAB29, 0, beam, domain="log", nthreads=8)
    print("paths equal", op == paths, "scores equal", all(np.float32(a) == np.float32(b) for a, b in zip(scores, os_)))
