"""Decode-only timing on the log-probabilities the cfg2 pipeline itself produces (the microbench's synthetic inputs prune
differently): separates 'decoder slowed by sharing SMs in the pipeline' from 'this data costs more per frame'.

    python tools/decode_alone_on_pipeline_output.py [N]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpu-accelerated-speech-recognition_b200"))
import numpy as np  # noqa: E402
import gasr  # noqa: E402
import synth  # noqa: E402

T, N, D, H, L, V, beam = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 161, 512, 3, 29, 16
ctx = gasr.Context(0)
x = synth.spectrogram_batch(1234, T, N, D)
pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
pipe.set_weights(*synth.rnn_weights(4321, D, H, L), *synth.fc_weights(99, H, V))
for _ in range(3):
    paths, scores = pipe.run_host(x)
print("pipeline stage times (ms):", ["%.3f" % v for v in pipe.stage_times()], "mode", pipe.stage_launches()[1])
logp = np.ascontiguousarray(pipe.logprobs().reshape(T * N, -1)[:, :V])
d = ctx.to_device(logp)
for it in range(3):
    ctx.sync(); ctx.timer_start()
    p2, s2 = ctx.ctc_decode(d, gasr.DOMAIN_LOG, T, N, V, V, beam, 0, synth.VOCAB29)
    ms = ctx.timer_stop()
    print(f"decode alone on the pipeline's log-probs: {ms:.3f} ms ({1e3 * ms / T:.2f} us/frame), same result: {p2 == paths}")
