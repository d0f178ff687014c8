"""GPU parity tests of the round-2 throughput path (-m gpu): the tcgen05 recurrence for groups of 128 utterances
(rnn_wide.cu), called through the C ABI (gasr_rnn_forward), against the CPU oracle.  Tolerance: 1e-4 absolute on fp32
acoustic-model values (BASELINE.json north_star)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))

pytestmark = pytest.mark.gpu

AM_TOL = 1e-4


@pytest.fixture(scope="module")
def gasr():
    import gasr as g
    return g


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _rnn(gasr, ctx, T, N, D, H, L, x, w_ih, w_hh, b_ih, b_hh):
    dx = ctx.to_device(x)
    dw = [[ctx.to_device(m) for m in lst] for lst in (w_ih, w_hh, b_ih, b_hh)]
    hid = [ctx.malloc(T * N * H * 4) for _ in range(L)]
    ctx.rnn_forward(gasr.CELL_TANH, False, T, N, D, H, L, dw[0], dw[1], dw[2], dw[3], dx, hid)
    out = [ctx.to_host(h, (T * N, H)) for h in hid]
    for p in [dx] + hid + [m for lst in dw for m in lst]:
        ctx.free(p)
    return out


# (multicast, groups per cluster): the options are read when the context is created
@pytest.mark.parametrize("mc,groups", [(0, 1), (0, 2), (1, 1), (1, 2)])
@pytest.mark.parametrize("H,N,T,L", [(512, 128, 24, 1), (512, 300, 17, 2), (256, 256, 20, 1), (128, 200, 12, 1), (64, 130, 9, 2)])
def test_wide_recurrence_vs_oracle(gasr, O, monkeypatch, mc, groups, H, N, T, L):
    import synth
    monkeypatch.setenv("GASR_RNN", "w")
    monkeypatch.setenv("GASR_RNN_MC", str(mc))
    monkeypatch.setenv("GASR_RNN_G", str(groups))
    ctx = gasr.Context(0)
    D = 37
    x = synth.spectrogram_batch(H + N, T, N, D)
    w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(H * 3 + 1, D, H, L)
    out = _rnn(gasr, ctx, T, N, D, H, L, x, w_ih, w_hh, b_ih, b_hh)
    ctx.close()
    ref = O.rnn_forward(x, T, N, w_ih, w_hh, b_ih, b_hh, nthreads=8)
    for l in range(L):
        err = np.abs(out[l] - ref[l]).max()
        assert err < AM_TOL, f"layer {l}: {err}"


@pytest.mark.parametrize("groups", [1, 2])
@pytest.mark.parametrize("H,N,T,L", [(512, 256, 20, 1), (512, 700, 13, 2), (256, 512, 16, 1), (128, 300, 11, 2)])
def test_wide_recurrence_cta_pairs_vs_oracle(gasr, O, monkeypatch, groups, H, N, T, L):
    """rnn_wide2.cu: tcgen05.mma.cta_group::2, groups of 256 utterances (GASR_RNN_PAIR=1)."""
    import synth
    monkeypatch.setenv("GASR_RNN", "w")
    monkeypatch.setenv("GASR_RNN_PAIR", "1")
    monkeypatch.setenv("GASR_RNN_G", str(groups))
    ctx = gasr.Context(0)
    D = 37
    x = synth.spectrogram_batch(H + N, T, N, D)
    w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(H * 3 + 1, D, H, L)
    out = _rnn(gasr, ctx, T, N, D, H, L, x, w_ih, w_hh, b_ih, b_hh)
    ctx.close()
    ref = O.rnn_forward(x, T, N, w_ih, w_hh, b_ih, b_hh, nthreads=8)
    for l in range(L):
        err = np.abs(out[l] - ref[l]).max()
        assert err < AM_TOL, f"layer {l}: {err}"


def _same(gp, gs, op, os_):
    assert gp == op
    a, b = np.array(gs, dtype=np.float32), np.array(os_, dtype=np.float32)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()


@pytest.mark.parametrize("V,beam,T,N", [(29, 16, 120, 40), (29, 32, 60, 24), (29, 8, 200, 33), (5, 3, 90, 16), (32, 32, 40, 12),
                                        (29, 1, 50, 8), (4, 9, 30, 8)])
def test_warp_decoder_vs_oracle(gasr, O, monkeypatch, V, beam, T, N):
    """The throughput decoder (one warp per utterance: what batches of > 296 utterances use) forced on small batches:
    random log-probabilities, heavily quantised ones (exact score ties every frame -> raw-string tie-breaks through the
    prefix-relation matrix) and the probability domain; transcripts and fp32 scores bit-exact against the oracle."""
    import synth
    monkeypatch.setenv("GASR_CTC_KERNEL", "w")
    ctx = gasr.Context(0)
    vocab = synth.VOCAB29 if V == 29 else bytes(range(1, V + 1))
    rng = np.random.default_rng(V * 100 + beam)
    logits = rng.normal(size=(T, N, V)).astype(np.float32) * 2.0
    lp = (logits - np.log(np.exp(logits.astype(np.float64)).sum(-1, keepdims=True))).astype(np.float32)
    cases = [("log", lp), ("log", (np.round(lp * 2.0) / 2.0).astype(np.float32))]          # second: many exact ties
    p = np.exp(lp[:24].astype(np.float64)); p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
    cases.append(("prob", p))
    for dom, x in cases:
        x = np.ascontiguousarray(x)
        gp, gs = ctx.ctc_decode_host(x, gasr.DOMAIN_LOG if dom == "log" else gasr.DOMAIN_PROB, beam, 0, vocab)
        op, os_ = O.ctc_decode(x, vocab, 0, beam, domain=dom, nthreads=8)
        _same(gp, gs, op, os_)
    ctx.close()
