"""GPU parity tests AT THE BASELINE.json SIZES (-m gpu): cfg1 (baseline/config.json hot path: H = 2048, V = 47, beam 100),
cfg2 (64 x 1000 frames end to end), cfg3 (bidirectional GRU H = 800 at batch 256), cfg4 (T = 4000, beam 32 / 128, golden
vectors of the oracle), cfg5 (batches of 1024 utterances through the job API, device-generated inputs, 2-GPU gather).
Everything goes through the C ABI (ctypes -> libgasr.so); the CPU oracle / its committed golden vectors are the checker.
Tolerances: fp32 acoustic values 1e-4 absolute, transcripts and fp32 beam scores bit-exact (BASELINE.json north_star)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))
GOLDEN = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu

AM_TOL = 1e-4


@pytest.fixture(scope="module")
def gasr():
    import gasr as g
    return g


@pytest.fixture(scope="module")
def ctx(gasr):
    c = gasr.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def _assert_same(gp, gs, op, os_):
    assert gp == op
    assert (_bits(gs) == _bits(os_)).all(), (gs, os_)


def _run_rnn(gasr, ctx, cell, bidir, T, N, D, H, L, x, w_ih, w_hh, b_ih, b_hh, precision=0):
    Dn = 2 if bidir else 1
    dx = ctx.to_device(x)
    dw = [[ctx.to_device(a) for a in lst] for lst in (w_ih, w_hh, b_ih, b_hh)]
    hid = [ctx.malloc(T * N * Dn * H * 4) for _ in range(L)]
    ctx.rnn_forward(cell, bidir, T, N, D, H, L, dw[0], dw[1], dw[2], dw[3], dx, hid, precision)
    out = [ctx.to_host(h, (T * N, Dn * H)) for h in hid]
    for p in [dx] + sum(dw, []) + hid:
        ctx.free(p)
    return out


# ------------------------------------------------------------------------------------------------ cfg1
def test_cfg1_baseline_config_hot_path(gasr, ctx, O):
    """baseline/config.json:3-13 restricted to the hot path: one utterance, T = 200, RNN input = hidden_3
    = 2048 (baseline/model.py:30), 1-layer tanh RNN H = 2048, Linear 2048 -> 47 + log-softmax, beam 100 (> V)."""
    import synth
    T, N, D, H, L, V, beam = 200, 1, 2048, 2048, 1, 47, 100
    vocab = bytes(range(1, V + 1))
    x = synth.spectrogram_batch(101, T, N, D)
    w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(102, D, H, L)
    fc_w, fc_b = synth.fc_weights(103, H, V)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, vocab)
    pipe.set_weights(w_ih, w_hh, b_ih, b_hh, fc_w, fc_b)
    paths, scores = pipe.run_host(x)
    logp = pipe.logprobs()
    ref = O.linear(O.rnn_forward(x, T, N, w_ih, w_hh, b_ih, b_hh, nthreads=8)[-1], fc_w, fc_b, act="logsoftmax")
    assert np.abs(logp - ref).max() < AM_TOL
    op, os_ = O.ctc_decode(logp.reshape(T, N, V), vocab, 0, beam, domain="log")
    _assert_same(paths, scores, op, os_)
    pipe.close()


@pytest.mark.parametrize("T,N,D,H,L", [(40, 1, 64, 2048, 2), (25, 3, 40, 1500, 1), (30, 4, 33, 1024, 2)])
def test_wide_hidden_layer_small_batch_resident_recurrence(gasr, O, T, N, D, H, L):
    """cfg1's recurrence shape class (H = 2048, a handful of utterances): one persistent fp32 launch per layer with W_hh resident
    in the shared memory of the whole GPU (rnn_resident.cu) instead of one kernel per timestep; every layer against the oracle."""
    import synth
    x = synth.spectrogram_batch(H + N, T, N, D)
    w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(H + 1, D, H, L)
    c = gasr.Context(0)
    n0 = c.launch_count()
    out = _run_rnn(gasr, c, gasr.CELL_TANH, False, T, N, D, H, L, x, w_ih, w_hh, b_ih, b_hh)
    launches = c.launch_count() - n0
    c.close()
    assert launches < L * 12, f"{launches} launches: the per-timestep fallback ran"
    ref = O.rnn_forward(x, T, N, w_ih, w_hh, b_ih, b_hh, nthreads=8)
    for l in range(L):
        assert np.abs(out[l] - ref[l]).max() < AM_TOL, f"layer {l}"


def test_wide_linear_relu_on_tensor_cores(gasr, ctx, O):
    """The DeepSpeech FC layers (main.cpp:31-45, baseline/model.py:22-35: 2048-wide Linear + ReLU) run on the tcgen05 tile
    engine with bias + ReLU in the epilogue (Linear.cu:3-10,42-49 semantics), fp32-grade."""
    rng = np.random.default_rng(5)
    for rows, in_, out, act in ((200, 78, 2048, "relu"), (256, 2048, 2048, "relu"), (1000, 2048, 2048, "none")):
        x = rng.normal(size=(rows, in_)).astype(np.float32)
        W = (rng.normal(size=(in_, out)) / np.sqrt(in_)).astype(np.float32)
        b = rng.normal(size=(out,)).astype(np.float32)
        ref = O.linear(x, W, b, act=act)
        dx, dW, db, dy = ctx.to_device(x), ctx.to_device(W), ctx.to_device(b), ctx.malloc(rows * out * 4)
        ctx.linear(dx, in_, dW, db, dy, out, rows, in_, out, gasr.ACT_RELU if act == "relu" else gasr.ACT_NONE)
        got = ctx.to_host(dy, (rows, out))
        for q in (dx, dW, db, dy):
            ctx.free(q)
        assert np.abs(got - ref).max() < AM_TOL, (rows, in_, out, act)


# ------------------------------------------------------------------------------------------------ cfg2
def test_cfg2_full_size_pipeline_vs_oracle(gasr, ctx, O):
    """The whole cfg2 batch (64 utterances x 1000 frames, 3 layers H = 512, beam 16) end to end against the oracle: every
    log-probability within 1e-4, every transcript and score bit-exact on the GPU's own log-probabilities."""
    import synth
    T, N, D, H, L, V, beam = 1000, 64, 161, 512, 3, 29, 16
    x = synth.spectrogram_batch(1234, T, N, D)
    w = synth.rnn_weights(4321, D, H, L)
    fc = synth.fc_weights(99, H, V)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    paths, scores = pipe.run_host(x)
    logp = pipe.logprobs()
    ref = O.linear(O.rnn_forward(x, T, N, *w, nthreads=16)[-1], *fc, act="logsoftmax")
    assert np.abs(logp - ref).max() < AM_TOL
    op, os_ = O.ctc_decode(logp.reshape(T, N, V), synth.VOCAB29, 0, beam, domain="log", nthreads=16)
    _assert_same(paths, scores, op, os_)
    pipe.close()


def test_cfg2_streaming_latency_mode_vs_oracle(gasr, O, monkeypatch):
    """The opt-in streaming mode (GASR_STREAM=1: persistent kernels coupled by progress counters, 500 blocks x 3 layers of
    hand-offs) on the same full-size batch: log-probabilities of 8 utterances against the oracle, all transcripts against
    the oracle decoder."""
    import synth
    T, N, D, H, L, V, beam = 1000, 64, 161, 512, 3, 29, 16
    monkeypatch.setenv("GASR_STREAM", "1")
    c2 = gasr.Context(0)
    x = synth.spectrogram_batch(1234, T, N, D)
    w = synth.rnn_weights(4321, D, H, L)
    fc = synth.fc_weights(99, H, V)
    pipe = gasr.AsrPipeline(c2, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    paths, scores = pipe.run_host(x)
    assert pipe.stage_launches()[1] == -1
    logp = pipe.logprobs().reshape(T, N, V)
    K = 8
    xs = np.ascontiguousarray(x.reshape(T, N, D)[:, :K, :]).reshape(T * K, D)
    ref = O.linear(O.rnn_forward(xs, T, K, *w, nthreads=16)[-1], *fc, act="logsoftmax").reshape(T, K, V)
    assert np.abs(logp[:, :K, :] - ref).max() < AM_TOL
    op, os_ = O.ctc_decode(np.ascontiguousarray(logp), synth.VOCAB29, 0, beam, domain="log", nthreads=16)
    _assert_same(paths, scores, op, os_)
    pipe.close()
    c2.close()


# ------------------------------------------------------------------------------------------------ cfg3
def test_cfg3_gru_at_batch_256(gasr, ctx, O):
    """cfg3's layer shape at cfg3's batch: bidirectional GRU H = 800, N = 256 (two 128-row blocks of the tcgen05 step kernel),
    161-bin input, two layers so the second one sees the 1600-wide concatenation; every layer against the oracle."""
    import synth
    T, N, D, H, L = 12, 256, 161, 800, 2
    x = synth.spectrogram_batch(71, T, N, D)
    w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(72, D, H, L, cell_gates=3, bidir=True)
    out = _run_rnn(gasr, ctx, gasr.CELL_GRU, True, T, N, D, H, L, x, w_ih, w_hh, b_ih, b_hh)
    ref = O.gru_forward(x, T, N, H, L, True, w_ih, w_hh, b_ih, b_hh)
    for l in range(L):
        assert np.abs(out[l] - ref[l]).max() < AM_TOL, f"layer {l}"


@pytest.mark.parametrize("T,N,D,H,L", [(12, 256, 161, 800, 2), (300, 130, 40, 100, 1), (65, 33, 24, 64, 2)])
def test_cfg3_bf16_mode_persistent_gru_recurrence(gasr, O, monkeypatch, T, N, D, H, L):
    """GASR_PREC_BF16 (cfg3's mode): the recurrence of a (layer, direction) is ONE persistent launch with W_hh resident in shared
    memory as fp16 (gru_seq.cu).  Against the fp32 oracle within the mode's stated 2e-2; against the per-timestep kernel of the
    same mode (GASR_GRU=t: fp32-grade recurrence on the same bf16 projection) within 2e-3 -- what the fp16 operands cost.
    Shapes: cfg3's layer at cfg3's batch; 300 steps over ragged row blocks / a partial unit tile; a small two-layer stack."""
    import synth
    x = synth.spectrogram_batch(91, T, N, D)
    w_ih, w_hh, b_ih, b_hh = synth.rnn_weights(92, D, H, L, cell_gates=3, bidir=True)
    ref = O.gru_forward(x, T, N, H, L, True, w_ih, w_hh, b_ih, b_hh)
    outs, launches = {}, {}
    for force in ("", "t"):
        if force:
            monkeypatch.setenv("GASR_GRU", force)
        c = gasr.Context(0)
        n0 = c.launch_count()
        outs[force] = _run_rnn(gasr, c, gasr.CELL_GRU, True, T, N, D, H, L, x, w_ih, w_hh, b_ih, b_hh, precision=gasr.PREC_BF16)
        launches[force] = c.launch_count() - n0
        c.close()
    # one recurrence launch per (layer, direction) instead of T
    assert launches["t"] - launches[""] >= 2 * L * (T - 2), launches
    for l in range(L):
        assert np.abs(outs[""][l] - ref[l]).max() < 2e-2, f"layer {l} vs oracle"
        assert np.abs(outs[""][l] - outs["t"][l]).max() < 2e-3, f"layer {l} vs the per-timestep kernel"


# ------------------------------------------------------------------------------------------------ cfg4
@pytest.mark.parametrize("kind", ["random", "peaky"])
@pytest.mark.parametrize("beam", [8, 32, 128])
def test_cfg4_long_utterance_golden(gasr, ctx, kind, beam):
    """T = 4000 at beam 8 / 32 / 128 against the oracle's committed answer (tests/golden/ctc_cfg4.json, generated by
    tests/golden/make_golden_ctc.py from the same seeds): transcript and fp32 score bit-exact."""
    import synth
    g = json.load(open(os.path.join(GOLDEN, "ctc_cfg4.json")))
    case = [c for c in g["cases"] if c["kind"] == kind and c["beam"] == beam][0]
    gen = synth.random_logprobs if kind == "random" else synth.peaky_logprobs
    lp = gen(case["seed"], g["T"], 1, g["V"])
    paths, scores = ctx.ctc_decode_host(lp, gasr.DOMAIN_LOG, beam, 0, synth.VOCAB29)
    assert paths[0].hex() == case["path_hex"]
    assert int(_bits(scores)[0]) == case["score_bits"]


# ------------------------------------------------------------------------------------------------ cfg5
def test_device_generator_matches_host_generator(gasr, ctx):
    import synth
    T, N, D = 37, 5, 161
    d = ctx.malloc(T * N * D * 4)
    ctx.synth_spectrogram(d, 1234, T, N, D, first_utt=4099)
    got = ctx.to_host(d, (T * N, D))
    ctx.free(d)
    ref = synth.spectrogram_batch(1234, T, N, D, first_utt=4099)
    assert (_bits(got) == _bits(ref)).all()


def test_cfg5_batch_of_1024_vs_oracle(gasr, ctx, O):
    """One cfg5 batch (1024 utterances x 1000 frames) through the wave engine; rows never interact (RNN.cu:15-27), so the
    oracle checks a spread of utterances: log-probabilities 1e-4, transcripts and scores bit-exact."""
    import synth
    T, N, D, H, L, V, beam = 1000, 1024, 161, 512, 3, 29, 16
    w = synth.rnn_weights(4321, D, H, L)
    fc = synth.fc_weights(99, H, V)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    assert pipe.stage_launches()[1] == -2                      # wave engine
    pipe.set_weights(*w, *fc)
    d = ctx.malloc(T * N * D * 4)
    ctx.synth_spectrogram(d, 1234, T, N, D, first_utt=2048)
    paths, scores = pipe.run_device(d)
    logp = pipe.logprobs().reshape(T, N, V)
    pick = [0, 127, 128, 511, 777, 1023]
    xs = np.stack([synth.spectrogram_batch(1234, T, 1, D, first_utt=2048 + u).reshape(T, D) for u in pick], axis=1).reshape(T * len(pick), D)
    ref = O.linear(O.rnn_forward(xs, T, len(pick), *w, nthreads=16)[-1], *fc, act="logsoftmax").reshape(T, len(pick), V)
    got = np.ascontiguousarray(logp[:, pick, :])
    assert np.abs(got - ref).max() < AM_TOL
    op, os_ = O.ctc_decode(got, synth.VOCAB29, 0, beam, domain="log", nthreads=8)
    _assert_same([paths[u] for u in pick], [scores[u] for u in pick], op, os_)
    ctx.free(d)
    pipe.close()


def test_job_batches_in_flight_equal_batches_alone(gasr, ctx):
    """gasr_job_*: five batches over two lanes (device and host inputs) give exactly what each batch gives alone."""
    import synth
    T, N, D, H, L, V, beam = 120, 200, 40, 128, 2, 29, 8
    w = synth.rnn_weights(5, D, H, L)
    fc = synth.fc_weights(6, H, V)
    job = gasr.Job(0, T, N, D, H, L, V, beam, 0, synth.VOCAB29, lanes=2)
    job.set_weights(*w, *fc)
    xs = [synth.spectrogram_batch(7, T, N, D, first_utt=b * N) for b in range(5)]
    c0 = job.lane_context(0)
    dev = [c0.to_device(x) for x in xs]
    a = job.run_device(dev)
    b = job.run_host(xs)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (_bits(a[2]) == _bits(b[2])).all()
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    ml = job.cfg.max_len
    for i, x in enumerate(xs):
        paths, scores = pipe.run_host(x)
        gp, gs = gasr.unpack_results(a[0][i * N:(i + 1) * N], a[1][i * N:(i + 1) * N], a[2][i * N:(i + 1) * N], ml)
        _assert_same(gp, gs, paths, scores)
    assert job.last_ms() > 0 and job.launch_count() > 0
    for d in dev:
        c0.free(d)
    pipe.close()
    job.close()


def test_two_gpu_gather_equals_single_gpu(gasr):
    """Contiguous shards on two GPUs, host-side gather == the single-GPU result, bit for bit (skipped below 2 devices)."""
    import shard
    import synth
    if gasr.device_count() < 2:
        pytest.skip("needs two GPUs")
    T, N, D, H, L, V, beam = 150, 256, 161, 256, 2, 29, 16
    w = synth.rnn_weights(15, D, H, L)
    fc = synth.fc_weights(16, H, V)
    xs = [synth.spectrogram_batch(17, T, N, D, first_utt=b * N) for b in range(4)]
    res = []
    for world in (1, 2):
        parts = []
        for rank in range(world):
            lo, hi = shard.shard_range(len(xs), world, rank)
            job = gasr.Job(rank, T, N, D, H, L, V, beam, 0, synth.VOCAB29, lanes=2)
            job.set_weights(*w, *fc)
            parts.append(job.run_host(xs[lo:hi]))
            job.close()
        res.append(tuple(np.concatenate([p[i] for p in parts]) for i in range(3)))
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all() and (_bits(res[0][2]) == _bits(res[1][2])).all()
