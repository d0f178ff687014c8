"""ctypes front-end of the CPU oracle (oracle/_build/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (gpu-accelerated-speech-recognition_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)


def build(ref=False):
    """(Re)build liboracle.so; with ref=True also oracle/_ref (needs /root/reference and nvcc)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        try:
            _LIB = ctypes.CDLL(path)
        except OSError:
            build()
            _LIB = ctypes.CDLL(path)
        _LIB.oracle_logaddexp_f32.restype = ctypes.c_float
        _LIB.oracle_logaddexp_f32.argtypes = [ctypes.c_float, ctypes.c_float]
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ptr_array(arrs):
    return (c_float_p * len(arrs))(*[_fp(a) for a in arrs])


def logaddexp(a, b):
    return float(lib().oracle_logaddexp_f32(float(a), float(b)))


def ctc_decode(scores, vocab, blank, beam, domain="log", merge="identity", nbest=1, nthreads=1, max_len=None):
    """scores: [T, N, V] float32 (probabilities for domain='prob', log-probs for 'log').
    Returns (paths, scores) -- lists of N bytes / floats -- or, for nbest > 1, per utterance lists."""
    s = _f32(scores)
    T, N, V = s.shape
    vocab = bytes(vocab)
    assert len(vocab) == V
    max_len = max_len or (T + 1)
    out_paths = np.zeros((N, nbest, max_len), dtype=np.uint8)
    out_lens = np.zeros((N, nbest), dtype=np.int32)
    out_scores = np.zeros((N, nbest), dtype=np.float32)
    out_counts = np.zeros((N,), dtype=np.int32)
    rc = lib().oracle_ctc_decode(
        _fp(s), T, N, V, V, vocab, int(blank), int(beam), 0 if domain == "prob" else 1,
        0 if merge == "identity" else 1, max_len, nbest,
        out_paths.ctypes.data_as(ctypes.c_char_p), out_lens.ctypes.data_as(c_int_p), _fp(out_scores),
        out_counts.ctypes.data_as(c_int_p), int(nthreads))
    if rc != 0:
        raise ValueError("oracle_ctc_decode: invalid arguments")
    if nbest == 1:
        paths = [bytes(out_paths[n, 0, : min(out_lens[n, 0], max_len)]) for n in range(N)]
        return paths, [float(x) for x in out_scores[:, 0]]
    res_p, res_s = [], []
    for n in range(N):
        k = min(int(out_counts[n]), nbest)
        res_p.append([bytes(out_paths[n, r, : min(out_lens[n, r], max_len)]) for r in range(k)])
        res_s.append([float(out_scores[n, r]) for r in range(k)])
    return res_p, res_s


def ctc_trace(scores, vocab, blank, beam, domain="prob", merge="identity"):
    """Single utterance [T, V]; returns (per-frame [(raw_bytes, score), ...], best_path, best_score)."""
    s = _f32(scores)
    T, V = s.shape
    max_len = T + 1
    ts = np.zeros((T, beam), dtype=np.float32)
    tl = np.zeros((T, beam), dtype=np.int32)
    tp = np.zeros((T, beam, max_len), dtype=np.uint8)
    tc = np.zeros((T,), dtype=np.int32)
    op = np.zeros((max_len,), dtype=np.uint8)
    ol = ctypes.c_int(0)
    osc = ctypes.c_float(0)
    rc = lib().oracle_ctc_trace(
        _fp(s), T, V, bytes(vocab), int(blank), int(beam), 0 if domain == "prob" else 1,
        0 if merge == "identity" else 1, max_len, _fp(ts), tl.ctypes.data_as(c_int_p),
        tp.ctypes.data_as(ctypes.c_char_p), tc.ctypes.data_as(c_int_p), op.ctypes.data_as(ctypes.c_char_p),
        ctypes.byref(ol), ctypes.byref(osc))
    if rc != 0:
        raise ValueError("oracle_ctc_trace: invalid arguments")
    frames = [[(bytes(tp[t, r, : tl[t, r]]), float(ts[t, r])) for r in range(tc[t])] for t in range(T)]
    return frames, bytes(op[: ol.value]), float(osc.value)


def ctc_decode_lens(scores, lens, vocab, blank, beam, domain="log", nbest=1):
    """Variable-length batch (baseline/main.py:45 `decoder.decode(output, out_lens)`): utterance n is decoded over its first
    lens[n] frames only, i.e. exactly CTC-REF on scores[:lens[n], n].  Returns what ctc_decode returns."""
    s = _f32(scores)
    res_p, res_s = [], []
    for n in range(s.shape[1]):
        p1, s1 = ctc_decode(s[: int(lens[n]), n : n + 1], vocab, blank, beam, domain=domain, nbest=nbest)
        res_p.append(p1[0]); res_s.append(s1[0])
    return res_p, res_s


def ctc_timesteps(scores, vocab, blank, beam, domain="log", nbest=1):
    """Per-token timesteps of one utterance [T, V] (the `timesteps` output baseline/main.py:45 takes from its decoder; ctcdecode is
    not vendored, so the definition is this package's): for output character j of a kept path, the first frame at which the label
    prefix path[0..j] was a kept beam state.  Derived from the per-frame kept beams of CTC-REF (ctc_trace).
    Returns (paths, timesteps): the nbest kept paths of the last frame, best first, and one list of frames per path."""
    vocab = bytes(vocab)
    frames, _, _ = ctc_trace(scores, vocab, blank, beam, domain=domain)
    T = len(frames)
    blank_ch = vocab[blank : blank + 1]
    first = {}
    for t, kept in enumerate(frames):
        for raw, _ in kept:
            first.setdefault(raw[:-1] if raw.endswith(blank_ch) else raw, t)
    paths, stamps = [], []
    for raw, _ in frames[-1][:nbest]:
        if T == 1:                      # the initial path is returned as is, blank included
            paths.append(raw); stamps.append([0] * len(raw))
            continue
        lab = raw[:-1] if raw.endswith(blank_ch) else raw
        paths.append(lab)
        stamps.append([first[lab[: j + 1]] for j in range(len(lab))])
    return paths, stamps


def matmul(x, y):
    x, y = _f32(x), _f32(y)
    z = np.zeros((x.shape[0], y.shape[1]), dtype=np.float32)
    lib().oracle_matmul(_fp(x), _fp(y), _fp(z), x.shape[0], x.shape[1], y.shape[1])
    return z


def linear(x, W, b, act="relu"):
    """x [rows, in], W [in, out] (reference layout), b [out]; act in none|relu|logsoftmax."""
    x, W = _f32(x), _f32(W)
    b = _f32(b) if b is not None else None
    y = np.zeros((x.shape[0], W.shape[1]), dtype=np.float32)
    lib().oracle_linear(_fp(x), x.shape[0], x.shape[1], W.shape[1], _fp(W), _fp(b) if b is not None else None,
                        {"none": 0, "relu": 1, "logsoftmax": 2}[act], _fp(y))
    return y


def log_softmax(x):
    x = _f32(x)
    y = np.zeros_like(x)
    lib().oracle_log_softmax(_fp(x), _fp(y), x.shape[0], x.shape[1])
    return y


def rnn_forward(x, T, N, w_ih, w_hh, b_ih, b_hh, nthreads=1):
    """x [T*N, in] time-major; per-layer weight lists in reference layout. Returns list of [T*N, H] per layer."""
    x = _f32(x)
    L = len(w_ih)
    H = w_hh[0].shape[0]
    w_ih = [_f32(a) for a in w_ih]
    w_hh = [_f32(a) for a in w_hh]
    b_ih = [_f32(a) for a in b_ih]
    b_hh = [_f32(a) for a in b_hh]
    hid = [np.zeros((T * N, H), dtype=np.float32) for _ in range(L)]
    lib().oracle_rnn_forward(_fp(x), T, N, x.shape[1], H, L, _ptr_array(w_ih), _ptr_array(w_hh), _ptr_array(b_ih),
                             _ptr_array(b_hh), _ptr_array(hid), int(nthreads))
    return hid


def gru_forward(x, T, N, H, L, bidir, w_ih, w_hh, b_ih, b_hh):
    """Parameter lists indexed [l * D + d]; returns list of [T*N, D*H] per layer."""
    x = _f32(x)
    D = 2 if bidir else 1
    w_ih = [_f32(a) for a in w_ih]
    w_hh = [_f32(a) for a in w_hh]
    b_ih = [_f32(a) for a in b_ih]
    b_hh = [_f32(a) for a in b_hh]
    hid = [np.zeros((T * N, D * H), dtype=np.float32) for _ in range(L)]
    lib().oracle_gru_forward(_fp(x), T, N, x.shape[1], H, L, 1 if bidir else 0, _ptr_array(w_ih), _ptr_array(w_hh),
                             _ptr_array(b_ih), _ptr_array(b_hh), _ptr_array(hid))
    return hid


# ---- the reference's own GPU build (oracle/_ref), GPU box only ------------------------------------------
def ref_lib():
    """oracle/_ref/libgasr_ref.so (the unmodified reference sources + ref_driver.cu); None if not built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libgasr_ref.so")
        if not os.path.exists(path):
            return None
        _REF = ctypes.CDLL(path)
    return _REF


def ref_ctc_decode(probs, vocab, blank, beam):
    s = _f32(probs)
    T, N, V = s.shape
    max_len = 256
    out_paths = np.zeros((N, max_len), dtype=np.uint8)
    out_lens = np.zeros((N,), dtype=np.int32)
    out_scores = np.zeros((N,), dtype=np.float32)
    rc = ref_lib().ref_ctc_decode(_fp(s), T, N, V, bytes(vocab), int(beam), int(blank), max_len,
                                  out_paths.ctypes.data_as(ctypes.c_char_p), out_lens.ctypes.data_as(c_int_p),
                                  _fp(out_scores))
    return [bytes(out_paths[n, : out_lens[n]]) for n in range(N)], [float(v) for v in out_scores], rc


def ref_linear(x, W, b):
    x, W, b = _f32(x), _f32(W), _f32(b)
    y = np.zeros((x.shape[0], W.shape[1]), dtype=np.float32)
    rc = ref_lib().ref_linear_forward(_fp(x), x.shape[0], x.shape[1], W.shape[1], _fp(W), _fp(b), _fp(y))
    return y, rc


def ref_rnn(x, T, N, w_ih, w_hh, b_ih, b_hh):
    x = _f32(x)
    L, H = len(w_ih), w_hh[0].shape[0]
    w_ih = [_f32(a) for a in w_ih]
    w_hh = [_f32(a) for a in w_hh]
    b_ih = [_f32(a) for a in b_ih]
    b_hh = [_f32(a) for a in b_hh]
    y = np.zeros((T * N, H), dtype=np.float32)
    rc = ref_lib().ref_rnn_forward(_fp(x), T, N, x.shape[1], H, L, _ptr_array(w_ih), _ptr_array(w_hh),
                                   _ptr_array(b_ih), _ptr_array(b_hh), _fp(y))
    return y, rc
