"""GPU parity tests (-m gpu) of the decoder options baseline/main.py:45-46 takes from its decoder
(`decoder.decode(output, out_lens)` -> output, scores, timesteps, out_seq_len): per-utterance frame counts and per-token
timesteps, through the C ABI (gasr_ctc_decode_ex, gasr_asr_set_lengths / _enable_timesteps / _timesteps), against the CPU
oracle (CTC-REF on the first lens[n] frames; timesteps from its per-frame kept beams).  Strings, fp32 scores and frame
indices must be bit-exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gasr():
    import gasr as g
    return g


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _bits(a):
    return np.array(a, dtype=np.float32).view(np.uint32).tolist()


def _oracle(O, lp, lens, vocab, beam, nbest):
    paths, scores, stamps = [], [], []
    for n in range(lp.shape[1]):
        s = np.ascontiguousarray(lp[: lens[n], n])
        p, sc = O.ctc_decode(s[:, None, :], vocab, 0, beam, domain="log", nbest=nbest)
        tp, ts = O.ctc_timesteps(s, vocab, 0, beam, domain="log", nbest=nbest)
        if nbest == 1:
            assert tp[0] == p[0]
            paths.append(p[0]); scores.append(sc[0]); stamps.append(ts[0])
        else:
            assert tp == p[0]
            paths.append(p[0]); scores.append(sc[0]); stamps.append(ts)
    return paths, scores, stamps


# kernel: '' = by batch size (whole-sequence CTA kernel for small batches), c = 128-thread CTA kernel, w = warp per utterance;
# V = 40 / beam = 40 reach the general kernel
@pytest.mark.parametrize("kernel,V,beam,nbest", [("", 29, 16, 1), ("", 29, 8, 3), ("c", 29, 16, 1), ("w", 29, 16, 3), ("w", 29, 32, 1),
                                                 ("", 40, 6, 2), ("", 29, 40, 1)])
def test_ctc_lengths_and_timesteps_vs_oracle(gasr, O, monkeypatch, kernel, V, beam, nbest):
    import synth
    if kernel:
        monkeypatch.setenv("GASR_CTC_KERNEL", kernel)
    ctx = gasr.Context(0)
    T, N = 70, 14
    vocab = synth.VOCAB29 if V == 29 else bytes(range(1, V + 1))
    lp = synth.random_logprobs(100 + V + beam, T, N, V)
    rng = np.random.default_rng(V * beam)
    lens = rng.integers(1, T + 1, size=N).astype(np.int32)
    lens[0], lens[1], lens[2] = 1, T, 2
    gp, gs, gt = ctx.ctc_decode_host(lp, gasr.DOMAIN_LOG, beam, 0, vocab, nbest=nbest, lens=lens, timesteps=True)
    op, os_, ot = _oracle(O, lp, lens, vocab, beam, nbest)
    assert gp == op
    assert [_bits(s) for s in gs] == [_bits(s) for s in os_] if nbest > 1 else _bits(gs) == _bits(os_)
    assert gt == ot
    # lengths alone / timesteps alone agree with the combined call and with the plain decoder
    p2, s2 = ctx.ctc_decode_host(lp, gasr.DOMAIN_LOG, beam, 0, vocab, nbest=nbest, lens=lens)
    assert (p2, s2) == (gp, gs)
    full = np.full(N, T, dtype=np.int32)
    p3, s3, t3 = ctx.ctc_decode_host(lp, gasr.DOMAIN_LOG, beam, 0, vocab, nbest=nbest, timesteps=True)
    assert (p3, s3) == ctx.ctc_decode_host(lp, gasr.DOMAIN_LOG, beam, 0, vocab, nbest=nbest)
    assert (p3, s3, t3) == ctx.ctc_decode_host(lp, gasr.DOMAIN_LOG, beam, 0, vocab, nbest=nbest, lens=full, timesteps=True)
    # a length outside 1..T is an argument error
    bad = lens.copy(); bad[3] = T + 1
    with pytest.raises(gasr.GasrError) as e:
        ctx.ctc_decode_host(lp, gasr.DOMAIN_LOG, beam, 0, vocab, lens=bad)
    assert e.value.status == gasr.ERR_INVALID
    ctx.close()


@pytest.mark.parametrize("mode,N", [("wave", 300), ("chunked", 24), ("sequential", 24)])
def test_pipeline_variable_length_batch(gasr, O, monkeypatch, mode, N):
    """The fused pipeline with out_lens: every utterance's transcript equals CTC-REF on the first lens[n] frames of the pipeline's
    own log-probabilities; lengths on and next to the 50-frame chunk boundaries; switching lengths off restores the full decode."""
    import synth
    if mode != "wave":
        monkeypatch.setenv("GASR_WAVE", "0")
    if mode == "sequential":
        monkeypatch.setenv("GASR_CHUNK", "0")
    T, D, H, L, V, beam = 130, 40, 256, 2, 29, 8
    x = synth.spectrogram_batch(21, T, N, D)
    w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(22, D, H, L)
    fc_w, fc_b = synth.fc_weights(23, H, V)
    ctx = gasr.Context(0)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    pipe.set_weights(w_ih, w_hh, b_ih, b_hh, fc_w, fc_b)
    full_p, full_s = pipe.run_host(x)
    rng = np.random.default_rng(5)
    lens = rng.integers(1, T + 1, size=N).astype(np.int32)
    lens[:8] = [1, 49, 50, 51, 100, 101, T, 2]
    pipe.set_lengths(lens)
    pipe.enable_timesteps(True)
    gp, gs = pipe.run_host(x)
    gt = pipe.timesteps()
    logp = pipe.logprobs().reshape(T, N, V)
    check = list(range(8)) + list(range(8, N, max(1, N // 12)))
    for n in check:
        s = np.ascontiguousarray(logp[: lens[n], n])
        p, sc = O.ctc_decode(s[:, None, :], synth.VOCAB29, 0, beam, domain="log")
        tp, ts = O.ctc_timesteps(s, synth.VOCAB29, 0, beam, domain="log")
        assert gp[n] == p[0] and _bits([gs[n]]) == _bits(sc), n
        assert gt[n] == ts[0], n
    pipe.set_lengths(None)
    pipe.enable_timesteps(False)
    assert pipe.run_host(x) == (full_p, full_s)
    launches, m = pipe.stage_launches()
    assert {"wave": m == -2, "chunked": m > 0, "sequential": m == 0}[mode], m
    pipe.close()
    ctx.close()
