"""Micro-benchmarks of single stages (used for ncu captures): python tools/microbench.py {ctc,rnn,linear,gemm} [...]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))
import gasr  # noqa: E402
import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stage", choices=["ctc", "rnn", "linear", "gemm"])
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--N", type=int, default=64)
    ap.add_argument("--beam", type=int, default=16)
    ap.add_argument("--H", type=int, default=512)
    ap.add_argument("--D", type=int, default=161)
    ap.add_argument("--L", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--kind", default="random", choices=["random", "peaky", "flat"])
    ap.add_argument("--cell", default="tanh", choices=["tanh", "gru"])
    ap.add_argument("--bidir", action="store_true")
    a = ap.parse_args()
    ctx = gasr.Context(0)
    V = 29
    if a.stage == "ctc":
        if a.kind == "flat":
            lp = synth.random_logprobs(1234, a.T, a.N, V, scale=0.05)   # like a random-init model's outputs
        else:
            lp = (synth.random_logprobs if a.kind == "random" else synth.peaky_logprobs)(1234, a.T, a.N, V)
        d = ctx.to_device(lp)
        for it in range(a.iters):
            ctx.sync(); ctx.timer_start()
            p, s = ctx.ctc_decode(d, gasr.DOMAIN_LOG, a.T, a.N, V, V, a.beam, 0, synth.VOCAB29)
            ms = ctx.timer_stop()
            fb, sv = ctx.ctc_last_stats()
            print(f"ctc T={a.T} N={a.N} beam={a.beam}: {ms:.3f} ms  ({1e3 * ms / a.T:.2f} us/frame)  len0={len(p[0])} "
                  f"fallback_frames={fb} mean_survivors={sv / (a.T * a.N):.1f}")
    elif a.stage == "rnn":
        x = synth.spectrogram_batch(1, a.T, a.N, a.D)
        gru = a.cell == "gru"
        w = synth.rnn_weights(2, a.D, a.H, a.L, cell_gates=3 if gru else 1, bidir=a.bidir)
        dx = ctx.to_device(x)
        dw = [[ctx.to_device(m) for m in lst] for lst in w]
        hid = [ctx.malloc(a.T * a.N * a.H * (2 if a.bidir else 1) * 4) for _ in range(a.L)]
        for it in range(a.iters):
            ctx.sync(); ctx.timer_start()
            ctx.rnn_forward(gasr.CELL_GRU if gru else gasr.CELL_TANH, a.bidir, a.T, a.N, a.D, a.H, a.L, dw[0], dw[1], dw[2], dw[3], dx, hid)
            ms = ctx.timer_stop()
            print(f"rnn {a.cell}{' bidir' if a.bidir else ''} T={a.T} N={a.N} H={a.H} L={a.L}: {ms:.3f} ms ({1e3 * ms / a.T / a.L:.2f} us/step/layer)")
    elif a.stage == "linear":
        rows = a.T * a.N
        x = np.random.default_rng(0).normal(size=(rows, a.H)).astype(np.float32)
        W, b = synth.fc_weights(3, a.H, V)
        dx, dW, db, dy = ctx.to_device(x), ctx.to_device(W), ctx.to_device(b), ctx.malloc(rows * 32 * 4)
        for it in range(a.iters):
            ctx.sync(); ctx.timer_start()
            ctx.linear(dx, a.H, dW, db, dy, 32, rows, a.H, V, gasr.ACT_LOGSOFTMAX)
            ms = ctx.timer_stop()
            print(f"linear+logsoftmax rows={rows} in={a.H}: {ms:.3f} ms  {rows * (a.H + V) * 4 / ms / 1e6:.1f} GB/s")
    else:
        rows = a.T * a.N
        x = np.random.default_rng(0).normal(size=(rows, a.D)).astype(np.float32)
        W = np.random.default_rng(1).normal(size=(a.D, a.H)).astype(np.float32)
        dx, dW, dy = ctx.to_device(x), ctx.to_device(W), ctx.malloc(rows * a.H * 4)
        for it in range(a.iters):
            ctx.sync(); ctx.timer_start()
            ctx.matmul(dx, a.D, 0, dW, a.H, 0, dy, a.H, rows, a.D, a.H)
            ms = ctx.timer_stop()
            print(f"gemm simt {rows}x{a.D}x{a.H}: {ms:.3f} ms  {2.0 * rows * a.D * a.H / ms / 1e9:.2f} TFLOP/s")
            ref = ctx.to_host(dy, (rows, a.H))
            for prec, name in ((gasr.PREC_FP32, "tc 3xbf16"), (gasr.PREC_BF16, "tc bf16")):
                ctx.sync(); ctx.timer_start()
                ctx.xproj_gemm(dx, a.D, dW, None, dy, a.H, rows, a.D, a.H, prec)
                ms = ctx.timer_stop()
                got = ctx.to_host(dy, (rows, a.H))
                err = float(np.abs(got - ref).max()) / float(np.abs(ref).max())
                print(f"gemm {name} {rows}x{a.D}x{a.H}: {ms:.3f} ms  {2.0 * rows * a.D * a.H / ms / 1e9:.2f} TFLOP/s (algorithmic, "
                      f"incl. operand split)  max rel err vs simt {err:.2e}")
    ctx.close()


if __name__ == "__main__":
    main()
