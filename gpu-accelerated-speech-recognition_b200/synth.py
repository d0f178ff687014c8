"""Synthetic workloads of BASELINE.json: counter-based RNG (splitmix64) so that every rank, the CPU
oracle and the GPU path see bit-identical inputs and weights without shipping data files."""
import numpy as np

MASK = np.uint64(0xFFFFFFFFFFFFFFFF)
VOCAB29 = bytes([1]) + b" '" + bytes(range(ord("a"), ord("z") + 1))   # blank = 0x01 (lowest char), then space, ', a-z


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & MASK
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & MASK
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & MASK
    return z ^ (z >> np.uint64(31))


def uniform01(seed, n, offset=0):
    """n floats in [0, 1): top 24 bits of splitmix64(seed-mixed counter)."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        r = splitmix64(idx ^ splitmix64(np.uint64(seed) + np.zeros(1, dtype=np.uint64)))
    return ((r >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)).astype(np.float32)


def uniform(seed, shape, lo, hi, offset=0):
    n = int(np.prod(shape))
    return (lo + (hi - lo) * uniform01(seed, n, offset)).astype(np.float32).reshape(shape)


def spectrogram_batch(seed, T, N, D, first_utt=0):
    """x[T*N, D] time-major, ~U[0,1) like baseline/main.py:39; utterance u's values depend only on (seed, u)."""
    x = np.empty((T, N, D), dtype=np.float32)
    for n in range(N):
        x[:, n, :] = uniform01(seed, T * D, offset=(first_utt + n) * T * D).reshape(T, D)
    return x.reshape(T * N, D)


def rnn_weights(seed, in_, H, L, cell_gates=1, bidir=False):
    """torch-default init U(+-1/sqrt(H)) in the reference's [in, out] layout."""
    D = 2 if bidir else 1
    k = 1.0 / np.sqrt(H)
    w_ih, w_hh, b_ih, b_hh = [], [], [], []
    off = 0
    for l in range(L):
        in_l = in_ if l == 0 else D * H
        for _ in range(D):
            for lst, shape in ((w_ih, (in_l, cell_gates * H)), (w_hh, (H, cell_gates * H)),
                               (b_ih, (cell_gates * H,)), (b_hh, (cell_gates * H,))):
                lst.append(uniform(seed, shape, -k, k, offset=off))
                off += int(np.prod(shape))
    return w_ih, w_hh, b_ih, b_hh


def fc_weights(seed, in_, V):
    k = 1.0 / np.sqrt(in_)
    return uniform(seed, (in_, V), -k, k), uniform(seed, (V,), -k, k, offset=in_ * V)


def random_logprobs(seed, T, N, V, scale=1.0):
    """log-softmax of pseudo-normal logits (sum of 4 uniforms, centred), cfg4 case (i)."""
    u = uniform01(seed, T * N * V * 4).reshape(T, N, V, 4).sum(-1) - 2.0
    logits = (u * (scale * np.sqrt(3.0))).astype(np.float32)
    m = logits.max(-1, keepdims=True)
    lse = np.log(np.exp((logits - m).astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    return ((logits - m) - lse).astype(np.float32)


def peaky_logprobs(seed, T, N, V, blank=0, period=8):
    """cfg4 case (ii): blank ~0.9 with a character spike every `period` frames."""
    r = uniform01(seed, T * N * 2).reshape(T, N, 2)
    p = np.full((T, N, V), 0.1 / (V - 1), dtype=np.float64)
    p[:, :, blank] = 0.9
    spike = (np.arange(T)[:, None] + (r[:, :, 0] * period).astype(int)) % period == 0
    ch = 1 + (r[:, :, 1] * (V - 1)).astype(int) % (V - 1)
    ch = np.where(ch >= blank + 1, ch, ch)  # labels 1..V-1 when blank = 0
    tt, nn = np.nonzero(spike)
    p[tt, nn, :] = 0.1 / (V - 1)
    p[tt, nn, ch[tt, nn]] = 0.9
    p /= p.sum(-1, keepdims=True)
    return np.log(p).astype(np.float32)
