"""CPU tests that pin the oracle (oracle/) against the reference's own fixtures and the committed goldens."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return json.load(open(os.path.join(GOLDEN, name)))


# ---- nn_test.cpp known answers (4-decimal comments) -----------------------------------------------------
def test_linear_nn_test_fixture():
    g = _load("nn_test.json")["linear"]
    y = O.linear(np.array(g["x"]), np.array(g["w_in_out"]), np.array(g["b"]), act="relu")
    assert np.abs(y - np.array(g["expected_4dp"])).max() < 1e-4  # comment precision: 4 decimals


def test_rnn_nn_test_fixture():
    g = _load("nn_test.json")["rnn"]
    hid = O.rnn_forward(np.array(g["x_time_major"]), g["T"], g["N"], [np.array(g["w_ih"])], [np.array(g["w_hh"])],
                        [np.array(g["b_ih"])], [np.array(g["b_hh"])])
    assert np.abs(hid[0] - np.array(g["expected_4dp"])).max() < 1e-4


# ---- main.cpp CTC vector -------------------------------------------------------------------------------
@pytest.mark.parametrize("merge", ["identity", "hash31"])
def test_ctc_main_vector(merge):
    g = _load("ctc_main.json")
    P = np.array(g["probs"], dtype=np.float32).reshape(g["T"], 1, 4)
    for beam, (path, prob) in g["expected"].items():
        p, s = O.ctc_decode(P, g["vocab"].encode(), g["blank"], int(beam), domain="prob", merge=merge)
        assert p[0].decode() == path
        assert np.float32(s[0]) == np.float32(prob)  # bit-exact fp32


def test_ctc_main_vector_per_frame_beams():
    g = _load("ctc_main.json")
    P = np.array(g["probs"], dtype=np.float32).reshape(g["T"], 4)
    frames, best, score = O.ctc_trace(P, g["vocab"].encode(), g["blank"], 2, domain="prob")
    for t, exp in enumerate(g["beam2_frames"]):
        assert len(frames[t]) == len(exp)
        for (raw, sc), (epath, esc) in zip(frames[t], exp):
            # the last frame's raw string keeps the stripped blank (CTCBeamSearch.cu:452-456)
            got = raw.decode()
            if t == g["T"] - 1 and got.endswith("$"):
                got = got[:-1]
            assert got == epath
            assert abs(sc - esc) < 5e-9
    assert best == b"cbacbc"


def test_ctc_log_mode_matches_prob_mode_labels():
    rng = np.random.default_rng(7)
    vocab = b"\x01 'abcdefghijklmnopqrstuvwxyz"
    for _ in range(20):
        T, N, V = 12, 3, len(vocab)
        logits = rng.normal(size=(T, N, V)).astype(np.float32) * 2
        P = np.exp(logits - logits.max(-1, keepdims=True))
        P = (P / P.sum(-1, keepdims=True)).astype(np.float32)
        for beam in (1, 4, 16, 40):
            pp, ps = O.ctc_decode(P, vocab, 0, beam, domain="prob")
            lp, ls = O.ctc_decode(np.log(P), vocab, 0, beam, domain="log")
            for a, b, sa, sb in zip(pp, lp, ps, ls):
                if a == b:
                    assert abs(np.log(sa) - sb) < 1e-3
            assert sum(a == b for a, b in zip(pp, lp)) >= N - 1  # near-ties may flip a label


def test_ctc_edge_cases():
    vocab = b"$ab"
    # T == 1: the rank-0 initial path is returned unstripped (blank included)
    p, s = O.ctc_decode(np.array([[[0.7, 0.2, 0.1]]], dtype=np.float32), vocab, 0, 2, domain="prob")
    assert p == [b"$"] and np.float32(s[0]) == np.float32(0.7)
    # all-blank utterance decodes to the empty string
    P = np.tile(np.array([0.98, 0.01, 0.01], dtype=np.float32), (6, 1, 1))
    p, s = O.ctc_decode(P, vocab, 0, 3, domain="prob")
    assert p == [b""]
    # repeated label separated by blank survives ("aa"), unseparated collapses ("a")
    hot = lambda i: np.eye(3, dtype=np.float32)[i] * 0.97 + 0.01
    p, _ = O.ctc_decode(np.stack([hot(1), hot(0), hot(1)])[:, None, :], vocab, 0, 3, domain="prob")
    assert p == [b"aa"]
    p, _ = O.ctc_decode(np.stack([hot(1), hot(1), hot(1)])[:, None, :], vocab, 0, 3, domain="prob")
    assert p == [b"a"]
    # beam wider than the vocabulary (the literal reference faults; intended semantics)
    p1, s1 = O.ctc_decode(P, vocab, 0, 50, domain="prob")
    assert p1 == [b""]
    # exact score ties break towards the smaller raw string
    P = np.full((3, 1, 3), 1.0 / 3, dtype=np.float32)
    p, _ = O.ctc_decode(P, vocab, 0, 2, domain="prob")
    assert isinstance(p[0], bytes)
    # invalid arguments
    with pytest.raises(ValueError):
        O.ctc_decode(np.zeros((0, 1, 3), dtype=np.float32), vocab, 0, 2)


def test_ctc_nbest_is_sorted_and_threads_agree():
    rng = np.random.default_rng(3)
    vocab = b"\x01abcdefg"
    lp = np.log(rng.dirichlet(np.ones(len(vocab)), size=(30, 5))).astype(np.float32)
    p1, s1 = O.ctc_decode(lp, vocab, 0, 8, nbest=8, nthreads=1)
    p4, s4 = O.ctc_decode(lp, vocab, 0, 8, nbest=8, nthreads=4)
    assert p1 == p4 and s1 == s4
    for sc in s1:
        assert all(a >= b for a, b in zip(sc, sc[1:]))


def test_logaddexp_against_float64():
    rng = np.random.default_rng(0)
    a = rng.uniform(-80, 0, 5000).astype(np.float32)
    b = (a + rng.uniform(-25, 25, 5000)).astype(np.float32)
    got = np.array([O.logaddexp(x, y) for x, y in zip(a, b)], dtype=np.float64)
    ref = np.logaddexp(a.astype(np.float64), b.astype(np.float64))
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    assert (np.abs(got - ref) <= 1.0 * ulp + 1e-7).all()
    assert O.logaddexp(-np.inf, -np.inf) == -np.inf
    assert O.logaddexp(-np.inf, -3.0) == -3.0
    assert O.logaddexp(-2.0, -40.0) == -2.0
    assert O.logaddexp(-1.5, -2.5) == O.logaddexp(-2.5, -1.5)


# ---- torch / baseline/model.py goldens ------------------------------------------------------------------
def test_rnn3_torch_golden():
    g = np.load(os.path.join(GOLDEN, "rnn3_torch.npz"))
    L = 3
    hid = O.rnn_forward(g["x"], int(g["T"]), int(g["N"]), [g[f"w_ih{l}"] for l in range(L)],
                        [g[f"w_hh{l}"] for l in range(L)], [g[f"b_ih{l}"] for l in range(L)],
                        [g[f"b_hh{l}"] for l in range(L)], nthreads=2)
    assert np.abs(hid[-1] - g["y"]).max() < 1e-5


def test_bigru_torch_golden():
    g = np.load(os.path.join(GOLDEN, "bigru_torch.npz"))
    L, H = int(g["L"]), int(g["H"])
    keys = [(l, d) for l in range(L) for d in range(2)]
    hid = O.gru_forward(g["x"], int(g["T"]), int(g["N"]), H, L, True, [g[f"w_ih{l}_{d}"] for l, d in keys],
                        [g[f"w_hh{l}_{d}"] for l, d in keys], [g[f"b_ih{l}_{d}"] for l, d in keys],
                        [g[f"b_hh{l}_{d}"] for l, d in keys])
    assert np.abs(hid[-1] - g["y"]).max() < 1e-5


def test_deepspeech_baseline_model_golden():
    """baseline/model.py:37-49 end to end: 3x(Linear+ReLU) -> RNN -> Linear+ReLU -> Linear -> log_softmax."""
    g = np.load(os.path.join(GOLDEN, "deepspeech_small.npz"))
    x_bt = g["x_bt"]
    B, T, D = x_bt.shape
    x = np.ascontiguousarray(x_bt.transpose(1, 0, 2)).reshape(T * B, D)  # time-major rows t*N+n
    for i in range(3):
        x = O.linear(x, g[f"fc{i}_w"], g[f"fc{i}_b"], act="relu")
    h = O.rnn_forward(x, T, B, [g["rnn_w_ih"]], [g["rnn_w_hh"]], [g["rnn_b_ih"]], [g["rnn_b_hh"]])[-1]
    h = O.linear(h, g["fc3_w"], g["fc3_b"], act="relu")
    logp = O.linear(h, g["fc4_w"], g["fc4_b"], act="logsoftmax")
    assert np.abs(logp.reshape(T, B, -1) - g["logp_tnv"]).max() < 1e-5


def test_ctc_lengths_and_timesteps_definitions():
    """Oracle-side definitions of the two decoder options of baseline/main.py:45-46 (out_lens in, timesteps out)."""
    from oracle import oracle as O
    vocab = b"$ab"
    # a a - b  ->  "ab".  Beam 1: 'a' is kept from frame 0, "ab" from frame 3; a wider beam keeps "ab" (as a low-ranked
    # state) from frame 1 on -- a timestep is the first frame at which the prefix was a kept state
    P = np.array([[0.1, 0.8, 0.1], [0.1, 0.8, 0.1], [0.8, 0.1, 0.1], [0.1, 0.1, 0.8]], dtype=np.float32)
    assert O.ctc_timesteps(P, vocab, 0, 1, domain="prob") == ([b"ab"], [[0, 3]])
    assert O.ctc_timesteps(P, vocab, 0, 3, domain="prob") == ([b"ab"], [[0, 1]])
    # T == 1: the initial path comes back as is
    paths, stamps = O.ctc_timesteps(P[:1], vocab, 0, 3, domain="prob")
    assert paths == [b"a"] and stamps == [[0]]
    # lengths: utterance n is CTC-REF on its first lens[n] frames; stamps are increasing and inside the utterance
    rng = np.random.default_rng(3)
    lp = np.log(rng.dirichlet(np.ones(5), size=(30, 4)).astype(np.float32))
    lens = [30, 1, 17, 2]
    voc = b"$abcd"
    gp, gs = O.ctc_decode_lens(lp, lens, voc, 0, 4, domain="log")
    for n, ln in enumerate(lens):
        p, s = O.ctc_decode(lp[:ln, n : n + 1], voc, 0, 4, domain="log")
        assert (gp[n], gs[n]) == (p[0], s[0])
        tp, ts = O.ctc_timesteps(lp[:ln, n], voc, 0, 4, domain="log", nbest=4)
        assert tp[0] == p[0]
        for path, st in zip(tp, ts):
            assert len(st) == len(path) and all(0 <= a < b < ln for a, b in zip(st, st[1:])) and all(0 <= a < ln for a in st)
