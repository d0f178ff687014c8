"""Timings of the other BASELINE.json configs (cfg3: bidirectional GRU stack, cfg4: decode-only stress) on one B200.
These are parity-test shapes, not bench lines; the numbers go into DESIGN.md for orientation.

    python tools/config_sweep.py [cfg3] [cfg4]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))
import gasr  # noqa: E402
import synth  # noqa: E402


def cfg3(ctx, T=1000, N=256, D=161, H=800, L=5, V=29, beam=32):
    x = synth.spectrogram_batch(1, T, N, D)
    w = synth.rnn_weights(2, D, H, L, cell_gates=3, bidir=True)
    fc = synth.fc_weights(3, 2 * H, V)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_GRU, True, T, N, D, H, L, V, beam, 0, synth.VOCAB29, precision=gasr.PREC_BF16)
    pipe.set_weights(*w, *fc)
    xd = ctx.to_device(x)
    for it in range(2):
        ctx.sync(); t0 = time.perf_counter()
        paths, _ = pipe.run_device(xd)
        dt = time.perf_counter() - t0
        print(f"cfg3 (bi-GRU H={H} L={L}, N={N}, T={T}, beam {beam}, bf16 projection): {dt * 1e3:.1f} ms -> RTFx {N * T * 0.01 / dt:.0f}; "
              f"stages ms {['%.1f' % v for v in pipe.stage_times()]} len0={len(paths[0])}")
    pipe.close()


def cfg4(ctx, T=4000, N=64, V=29):
    for kind, gen in (("random", synth.random_logprobs), ("peaky", synth.peaky_logprobs)):
        lp = gen(1, T, N, V)
        d = ctx.to_device(lp)
        for beam in (8, 32, 128):
            best = 1e9
            for it in range(2):
                ctx.sync(); ctx.timer_start()
                p, _ = ctx.ctc_decode(d, gasr.DOMAIN_LOG, T, N, V, V, beam, 0, synth.VOCAB29)
                best = min(best, ctx.timer_stop())
            print(f"cfg4 decode-only ({kind}, T={T}, N={N}, beam {beam}): {best:.2f} ms ({1e3 * best / T:.2f} us/frame) -> RTFx {N * T * 0.01 / (best * 1e-3):.0f}")
        ctx.free(d)


if __name__ == "__main__":
    ctx = gasr.Context(0)
    which = sys.argv[1:] or ["cfg3", "cfg4"]
    if "cfg4" in which: cfg4(ctx)
    if "cfg3" in which: cfg3(ctx)
    ctx.close()
