"""ctypes binding of libgasr.so (the C ABI in include/gasr.h) and a Python mirror of the reference's module API.

The classes keep the reference's names and argument order so parity tests read like the reference's own drivers:
    cuMatrix(data | rows, cols)            cuMatrix.h:18-49     toGpu / toCpu / getHost / getDev
    Linear(batch, in, out).initParams(w, b).forward(x)          Linear.h:6-27, Linear.cu:23-49
    RNN_Cell(batch, in, hid).initParams(...).forward(x, h, out) RNN_Cell.h:6-36, RNN_Cell.cu:35-74
    RNN(batch, in, hid, time_step, layers).forward(x)           RNN.h:8-35, RNN.cu:9-30
    CTCBeamSearch(vocab, vocabSize, beamWidth, blankID).decode(seqProb, timestep, batchSize)
                                                                CTCBeamSearch.h:107-131, CTCBeamSearch.cu:262-312
Everything computes on the GPU through libgasr.so; there is no CPU path here.  Importing this module without the
built library raises ImportError (run `make -C gpu-accelerated-speech-recognition_b200` or __graft_entry__.build()).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GASR_LIB") or os.path.join(_HERE, "libgasr.so")   # GASR_LIB: an instrumented build (tools only)

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_TRUNCATED = range(6)
ACT_NONE, ACT_RELU, ACT_LOGSOFTMAX = 0, 1, 2
DOMAIN_PROB, DOMAIN_LOG = 0, 1
CELL_TANH, CELL_GRU = 0, 1
PREC_FP32, PREC_BF16 = 0, 1

c_void_pp = ctypes.POINTER(ctypes.c_void_p)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)


class GasrError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"gasr status {status}: {message}")
        self.status = status


class AsrConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("cell", "bidirectional", "T", "N", "in_", "H", "L", "V", "beam", "blank", "precision", "nbest",
                 "max_len")]


if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is not built; run `make -C {_HERE}` (there is no CPU fallback)")
_lib = ctypes.CDLL(LIB_PATH)
_lib.gasr_last_error.restype = ctypes.c_char_p

# Every symbol include/gasr.h declares (tests/test_abi.py checks the header against this list and the .so).
EXPORTS = [
    "gasr_version", "gasr_last_error", "gasr_device_count", "gasr_ctx_create", "gasr_ctx_destroy", "gasr_ctx_sync",
    "gasr_ctx_sm_count", "gasr_timer_start", "gasr_timer_stop", "gasr_ctx_launch_count", "gasr_malloc_device",
    "gasr_free_device", "gasr_malloc_host", "gasr_free_host", "gasr_memcpy_h2d_on_stream", "gasr_memcpy_h2d",
    "gasr_memcpy_d2h", "gasr_memcpy_h2d_async", "gasr_memcpy_d2h_async", "gasr_memset_device", "gasr_memory_stats",
    "gasr_matmul", "gasr_matadd", "gasr_xproj_gemm", "gasr_linear_forward", "gasr_log_softmax", "gasr_rnn_cell_forward",
    "gasr_rnn_forward", "gasr_ctc_decode", "gasr_ctc_decode_ex", "gasr_ctc_last_stats", "gasr_ctc_decode_host", "gasr_asr_create", "gasr_asr_destroy",
    "gasr_asr_set_weights", "gasr_asr_run_host", "gasr_asr_run_device", "gasr_asr_logprobs", "gasr_asr_stage_times", "gasr_asr_stage_launches",
    "gasr_asr_submit_host", "gasr_asr_submit_device", "gasr_asr_collect", "gasr_asr_set_lengths", "gasr_asr_enable_timesteps",
    "gasr_asr_timesteps", "gasr_asr_profile", "gasr_asr_last_ms",
    "gasr_job_create", "gasr_job_destroy", "gasr_job_set_weights", "gasr_job_run_host", "gasr_job_run_device", "gasr_job_last_ms",
    "gasr_job_launch_count", "gasr_job_lane", "gasr_job_profile", "gasr_job_stage_times", "gasr_synth_spectrogram",
]


def _check(status):
    if status != OK:
        raise GasrError(status, _lib.gasr_last_error().decode(errors="replace"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def device_count():
    n = ctypes.c_int(0)
    _check(_lib.gasr_device_count(ctypes.byref(n)))
    return n.value


class Context:
    """gasr_ctx: one device + one stream + scratch workspaces."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        _check(_lib.gasr_ctx_create(int(device), ctypes.byref(self._h)))
        self.device = device

    _borrowed = False

    def close(self):
        if self._h and not self._borrowed:
            _lib.gasr_ctx_destroy(self._h)
        self._h = ctypes.c_void_p()

    def synth_spectrogram(self, dptr, seed, T, N, D, first_utt=0):
        """Fill dptr[T*N, D] on the device with the same values synth.spectrogram_batch produces on the host."""
        _check(_lib.gasr_synth_spectrogram(self._h, dptr, int(seed), T, N, D, int(first_utt)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- memory
    def malloc(self, nbytes):
        p = ctypes.c_void_p()
        _check(_lib.gasr_malloc_device(self._h, ctypes.c_size_t(int(nbytes)), ctypes.byref(p)))
        return p

    def free(self, p):
        _check(_lib.gasr_free_device(self._h, p))

    def pinned(self, shape, dtype=np.float32):
        """numpy view of a pinned host block (cuMatrix::mallocHost -> cudaHostAlloc, MemoryMonitor.cpp:9-18)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = ctypes.c_void_p()
        _check(_lib.gasr_malloc_host(self._h, ctypes.c_size_t(max(n, 1)), ctypes.byref(p)))
        buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(arr.nbytes)
        _check(_lib.gasr_memcpy_h2d(self._h, p, arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(arr.nbytes)))
        return p

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        _check(_lib.gasr_memcpy_h2d(self._h, dptr, arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(arr.nbytes)))

    def h2d_async(self, dptr, arr):
        _check(_lib.gasr_memcpy_h2d_async(self._h, dptr, arr.ctypes.data_as(ctypes.c_void_p),
                                          ctypes.c_size_t(arr.nbytes)))

    def d2h_into(self, arr, dptr):
        """Blocking device -> host copy into an existing C-contiguous array (e.g. a pinned block)."""
        _check(_lib.gasr_memcpy_d2h(self._h, arr.ctypes.data_as(ctypes.c_void_p), dptr, ctypes.c_size_t(arr.nbytes)))

    def to_host(self, dptr, shape, dtype=np.float32):
        out = np.empty(shape, dtype=dtype)
        _check(_lib.gasr_memcpy_d2h(self._h, out.ctypes.data_as(ctypes.c_void_p), dptr, ctypes.c_size_t(out.nbytes)))
        return out

    def memset(self, dptr, value, nbytes):
        _check(_lib.gasr_memset_device(self._h, dptr, int(value), ctypes.c_size_t(int(nbytes))))

    def sync(self):
        _check(_lib.gasr_ctx_sync(self._h))

    def sm_count(self):
        n = ctypes.c_int(0)
        _check(_lib.gasr_ctx_sm_count(self._h, ctypes.byref(n)))
        return n.value

    def timer_start(self):
        _check(_lib.gasr_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float(0)
        _check(_lib.gasr_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def ctc_last_stats(self):
        f, s = ctypes.c_longlong(0), ctypes.c_longlong(0)
        _check(_lib.gasr_ctc_last_stats(self._h, ctypes.byref(f), ctypes.byref(s)))
        return f.value, s.value

    def launch_count(self):
        n = ctypes.c_longlong(0)
        _check(_lib.gasr_ctx_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def memory_stats(self):
        d, h = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _check(_lib.gasr_memory_stats(self._h, ctypes.byref(d), ctypes.byref(h)))
        return d.value, h.value

    # -- dense math on device pointers
    def matmul(self, x, ldx, tx, y, ldy, ty, z, ldz, m, k, n):
        _check(_lib.gasr_matmul(self._h, x, ldx, tx, y, ldy, ty, z, ldz, m, k, n))

    def matadd(self, x, ldx, y, ldy, z, ldz, rows, cols, lam):
        _check(_lib.gasr_matadd(self._h, x, ldx, y, ldy, z, ldz, rows, cols, ctypes.c_float(lam)))

    def xproj_gemm(self, x, ldx, W, bias, y, ldy, rows, in_, out, precision=PREC_FP32):
        _check(_lib.gasr_xproj_gemm(self._h, x, ldx, W, bias, y, ldy, rows, in_, out, precision))

    def linear(self, x, ldx, W, b, y, ldy, rows, in_, out, act):
        _check(_lib.gasr_linear_forward(self._h, x, ldx, W, b, y, ldy, rows, in_, out, act))

    def log_softmax(self, x, ldx, y, ldy, rows, cols):
        _check(_lib.gasr_log_softmax(self._h, x, ldx, y, ldy, rows, cols))

    def ctc_decode(self, scores_dev, domain, T, N, V, ld, beam, blank, vocab, max_len=None, nbest=1, lens=None, timesteps=False):
        """lens: frames per utterance (baseline/main.py:45 out_lens); timesteps=True also returns, per utterance (and per kept
        path when nbest > 1), the frame at which each output character's prefix first entered the beam."""
        max_len = T + 1 if max_len is None else max_len
        paths = np.zeros((N, nbest, max(max_len, 1)), dtype=np.uint8)
        out_lens = np.zeros((N, nbest), dtype=np.int32)
        scores = np.zeros((N, nbest), dtype=np.float32)
        counts = np.zeros((N,), dtype=np.int32)
        if lens is None and not timesteps:
            _check(_lib.gasr_ctc_decode(self._h, scores_dev, domain, T, N, V, ld, beam, blank, bytes(vocab), max_len,
                                        nbest, paths.ctypes.data_as(ctypes.c_char_p), out_lens.ctypes.data_as(c_int_p),
                                        _fp(scores), counts.ctypes.data_as(c_int_p)))
            return _unpack(paths, out_lens, scores, counts, nbest, max_len)
        lens_a = None if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
        ts = np.zeros((N, nbest, max(max_len, 1)), dtype=np.int32) if timesteps else None
        _check(_lib.gasr_ctc_decode_ex(self._h, scores_dev, domain, T, N, V, ld, beam, blank, bytes(vocab), max_len, nbest,
                                       None if lens_a is None else lens_a.ctypes.data_as(c_int_p),
                                       paths.ctypes.data_as(ctypes.c_char_p), out_lens.ctypes.data_as(c_int_p), _fp(scores),
                                       counts.ctypes.data_as(c_int_p), None if ts is None else ts.ctypes.data_as(c_int_p)))
        res = _unpack(paths, out_lens, scores, counts, nbest, max_len)
        if not timesteps:
            return res
        return res + (_unpack_ts(ts, out_lens, counts, nbest, max_len),)

    def ctc_decode_host(self, scores, domain, beam, blank, vocab, max_len=None, nbest=1, lens=None, timesteps=False):
        s = _f32(scores)
        T, N, V = s.shape
        if lens is not None or timesteps:
            d = self.to_device(s)
            try:
                return self.ctc_decode(d, domain, T, N, V, V, beam, blank, vocab, max_len, nbest, lens, timesteps)
            finally:
                self.free(d)
        max_len = T + 1 if max_len is None else max_len
        paths = np.zeros((N, nbest, max(max_len, 1)), dtype=np.uint8)
        lens = np.zeros((N, nbest), dtype=np.int32)
        sc = np.zeros((N, nbest), dtype=np.float32)
        counts = np.zeros((N,), dtype=np.int32)
        _check(_lib.gasr_ctc_decode_host(self._h, _fp(s), domain, T, N, V, beam, blank, bytes(vocab), max_len, nbest,
                                         paths.ctypes.data_as(ctypes.c_char_p), lens.ctypes.data_as(c_int_p), _fp(sc),
                                         counts.ctypes.data_as(c_int_p)))
        return _unpack(paths, lens, sc, counts, nbest, max_len)

    def rnn_forward(self, cell, bidir, T, N, in_, H, L, w_ih, w_hh, b_ih, b_hh, x, hiddens, precision=PREC_FP32):
        arr = lambda ps: (ctypes.c_void_p * len(ps))(*[p.value if isinstance(p, ctypes.c_void_p) else p for p in ps])
        _check(_lib.gasr_rnn_forward(self._h, cell, int(bidir), T, N, in_, H, L, arr(w_ih), arr(w_hh), arr(b_ih),
                                     arr(b_hh), x, arr(hiddens), precision))


def _unpack(paths, lens, scores, counts, nbest, max_len):
    N = paths.shape[0]
    if nbest == 1:
        return ([bytes(paths[n, 0, : min(int(lens[n, 0]), max_len)]) for n in range(N)],
                [float(scores[n, 0]) for n in range(N)])
    out_p, out_s = [], []
    for n in range(N):
        k = min(int(counts[n]), nbest)
        out_p.append([bytes(paths[n, r, : min(int(lens[n, r]), max_len)]) for r in range(k)])
        out_s.append([float(scores[n, r]) for r in range(k)])
    return out_p, out_s


def _unpack_ts(ts, lens, counts, nbest, max_len):
    N = ts.shape[0]
    if nbest == 1:
        return [[int(v) for v in ts[n, 0, : min(int(lens[n, 0]), max_len)]] for n in range(N)]
    return [[[int(v) for v in ts[n, r, : min(int(lens[n, r]), max_len)]] for r in range(min(int(counts[n]), nbest))] for n in range(N)]


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


# ---- mirror of the reference's module API ------------------------------------------------------------------
class cuMatrix:
    """Row-major [rows, cols] fp32 matrix with a host (numpy) and a device copy (cuMatrix.h:13-229)."""

    def __init__(self, data_or_rows, cols=None, channels=1, ctx=None):
        self.ctx = ctx or default_context()
        if cols is None or isinstance(data_or_rows, np.ndarray):
            host = _f32(data_or_rows)
            if cols is not None:
                host = host.reshape(-1, cols)
            self.host = host.copy()
        else:
            self.host = np.zeros((int(data_or_rows), int(cols)), dtype=np.float32)
        self.rows, self.cols, self.channels = self.host.shape[0], self.host.shape[1], channels
        self._dev = None

    def getDev(self):
        if self._dev is None:
            self._dev = self.ctx.malloc(max(self.host.nbytes, 16))   # zero-filled like mallocDev
        return self._dev

    def getHost(self):
        return self.host

    def toGpu(self):
        self.ctx.h2d(self.getDev(), self.host)
        return self

    def toCpu(self):
        self.host = self.ctx.to_host(self.getDev(), self.host.shape)
        return self

    def getLen(self):
        return self.rows * self.cols * self.channels

    def getRows(self):
        return self.rows

    def getCols(self):
        return self.cols

    def free(self):
        if self._dev is not None:
            self.ctx.free(self._dev)
            self._dev = None


def matrixMul(x, y, z):
    """z = x * y (cuMatrix.cpp:33-70); shape mismatch raises instead of exit(0)."""
    if x.cols != y.rows or z.rows != x.rows or z.cols != y.cols:
        raise GasrError(ERR_INVALID, "matrix mul dimension mismatch")
    x.ctx.matmul(x.getDev(), x.cols, 0, y.getDev(), y.cols, 0, z.getDev(), z.cols, x.rows, x.cols, y.cols)


def matrixMulTA(x, y, z):
    if x.rows != y.rows or z.rows != x.cols or z.cols != y.cols:
        raise GasrError(ERR_INVALID, "matrix mul dimension mismatch")
    x.ctx.matmul(x.getDev(), x.cols, 1, y.getDev(), y.cols, 0, z.getDev(), z.cols, x.cols, x.rows, y.cols)


def matrixMulTB(x, y, z):
    if x.cols != y.cols or z.rows != x.rows or z.cols != y.rows:
        raise GasrError(ERR_INVALID, "matrix mul dimension mismatch")
    x.ctx.matmul(x.getDev(), x.cols, 0, y.getDev(), y.cols, 1, z.getDev(), z.cols, x.rows, x.cols, y.rows)


def matrixAdd(x, y, z, lam):
    x.ctx.matadd(x.getDev(), x.cols, y.getDev(), y.cols, z.getDev(), z.cols, x.rows, x.cols, lam)


class Linear:
    """ReLU(x*W + b) (Linear.cu:42-49); `act` widens it to none / log-softmax for the output layer."""

    def __init__(self, batch_size, input_size, output_size, act=ACT_RELU, ctx=None):
        self.ctx = ctx or default_context()
        self.batch_size, self.input_size, self.output_size, self.act = batch_size, input_size, output_size, act
        self.w = cuMatrix(input_size, output_size, ctx=self.ctx)
        self.b = cuMatrix(output_size, 1, ctx=self.ctx)
        self.outputs = cuMatrix(batch_size, output_size, ctx=self.ctx)
        self.initRandom()

    def initRandom(self, rng=None):
        rng = rng or np.random.default_rng(0)
        self.w.host[...] = rng.uniform(-1, 1, self.w.host.shape)   # U[-1,1], bias 0 (Linear.cu:12-21)
        self.b.host[...] = 0
        self.w.toGpu(); self.b.toGpu()

    def initParams(self, weight, bias):
        self.w.host[...] = _f32(weight).reshape(self.w.host.shape)
        self.b.host[...] = _f32(bias).reshape(self.b.host.shape)
        self.w.toGpu(); self.b.toGpu()
        return self

    def forward(self, inputs):
        if inputs.rows != self.batch_size or inputs.cols != self.input_size:
            raise GasrError(ERR_INVALID, "matrix mul dimension mismatch")
        self.ctx.linear(inputs.getDev(), inputs.cols, self.w.getDev(), self.b.getDev(), self.outputs.getDev(),
                        self.output_size, self.batch_size, self.input_size, self.output_size, self.act)
        return self.outputs


class RNN_Cell:
    def __init__(self, batch_size, input_size, hidden_size, ctx=None):
        self.ctx = ctx or default_context()
        self.batch_size, self.input_size, self.hidden_size = batch_size, input_size, hidden_size
        self.w_ih = cuMatrix(input_size, hidden_size, ctx=self.ctx)
        self.w_hh = cuMatrix(hidden_size, hidden_size, ctx=self.ctx)
        self.b_ih = cuMatrix(hidden_size, 1, ctx=self.ctx)
        self.b_hh = cuMatrix(hidden_size, 1, ctx=self.ctx)
        self.initRandom()

    def initRandom(self, rng=None):
        rng = rng or np.random.default_rng(0)
        self.w_ih.host[...] = rng.uniform(-1, 1, self.w_ih.host.shape)   # RNN_Cell.cu:15-33
        self.w_hh.host[...] = rng.uniform(-1, 1, self.w_hh.host.shape)
        self.b_ih.host[...] = 0
        self.b_hh.host[...] = 0
        for m in (self.w_ih, self.w_hh, self.b_ih, self.b_hh):
            m.toGpu()

    def initParams(self, w_ih, w_hh, b_ih, b_hh):
        for m, v in ((self.w_ih, w_ih), (self.w_hh, w_hh), (self.b_ih, b_ih), (self.b_hh, b_hh)):
            m.host[...] = _f32(v).reshape(m.host.shape)
            m.toGpu()
        return self

    def forward(self, inputs, pre_hidden, outputs):
        _check(_lib.gasr_rnn_cell_forward(self.ctx._h, inputs.getDev(), pre_hidden.getDev(), self.w_ih.getDev(),
                                          self.w_hh.getDev(), self.b_ih.getDev(), self.b_hh.getDev(), outputs.getDev(),
                                          self.batch_size, self.input_size, self.hidden_size))
        return outputs


class RNN:
    """L stacked tanh cells over a time-major [T*N, in] input; returns the last layer's [T*N, H] (RNN.cu:9-30)."""

    def __init__(self, batch_size, input_size, hidden_size, time_step, num_layers, ctx=None):
        self.ctx = ctx or default_context()
        self.batch_size, self.input_size, self.hidden_size = batch_size, input_size, hidden_size
        self.time_step, self.num_layers = time_step, num_layers
        self.rnn_cell = [RNN_Cell(batch_size, input_size if i == 0 else hidden_size, hidden_size, ctx=self.ctx)
                         for i in range(num_layers)]
        self.hiddens = [cuMatrix(time_step * batch_size, hidden_size, ctx=self.ctx) for _ in range(num_layers)]

    def forward(self, inputs, precision=PREC_FP32):
        if inputs.rows != self.time_step * self.batch_size or inputs.cols != self.input_size:
            raise GasrError(ERR_INVALID, "RNN input shape mismatch")
        c = self.rnn_cell
        self.ctx.rnn_forward(CELL_TANH, False, self.time_step, self.batch_size, self.input_size, self.hidden_size,
                             self.num_layers, [m.w_ih.getDev() for m in c], [m.w_hh.getDev() for m in c],
                             [m.b_ih.getDev() for m in c], [m.b_hh.getDev() for m in c], inputs.getDev(),
                             [h.getDev() for h in self.hiddens], precision)
        return self.hiddens[-1]


class CTCBeamSearch:
    def __init__(self, vocab, vocabSize, beamWidth, blankID, domain=DOMAIN_PROB, ctx=None):
        self.ctx = ctx or default_context()
        self.vocab = bytes(vocab)[:vocabSize]
        self.vocabSize, self.beamWidth, self.blankID, self.domain = vocabSize, beamWidth, blankID, domain

    def decode(self, seqProb, timestep, batchSize, nbest=1):
        """seqProb: cuMatrix [timestep*batchSize, vocabSize]; returns [(string, score)] per utterance."""
        if seqProb.getCols() != self.vocabSize:
            raise GasrError(ERR_INVALID, "inconsistent vocabulary size in CTC decoder")   # CTCBeamSearch.cu:267-270
        paths, scores = self.ctx.ctc_decode(seqProb.getDev(), self.domain, timestep, batchSize, self.vocabSize,
                                            seqProb.cols, self.beamWidth, self.blankID, self.vocab, nbest=nbest)
        if nbest == 1:
            return list(zip(paths, scores))
        return [list(zip(p, s)) for p, s in zip(paths, scores)]


class AsrPipeline:
    """gasr_asr: RNN stack -> Linear -> log-softmax -> CTC beam search as one call."""

    def __init__(self, ctx, cell, bidirectional, T, N, in_, H, L, V, beam, blank, vocab, precision=PREC_FP32, nbest=1,
                 max_len=None):
        self.ctx = ctx
        self.cfg = AsrConfig(cell, int(bidirectional), T, N, in_, H, L, V, beam, blank, precision, nbest,
                             T + 1 if max_len is None else max_len)
        self._h = ctypes.c_void_p()
        _check(_lib.gasr_asr_create(ctx._h, ctypes.byref(self.cfg), bytes(vocab), ctypes.byref(self._h)))
        n = N * nbest
        self._paths = np.zeros((n, max(self.cfg.max_len, 1)), dtype=np.uint8)
        self._lens = np.zeros((n,), dtype=np.int32)
        self._scores = np.zeros((n,), dtype=np.float32)

    def close(self):
        if self._h:
            _lib.gasr_asr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def set_weights(self, w_ih, w_hh, b_ih, b_hh, fc_w, fc_b):
        keep = [[_f32(a) for a in lst] for lst in (w_ih, w_hh, b_ih, b_hh)]
        arrs = [(c_float_p * len(lst))(*[_fp(a) for a in lst]) for lst in keep]
        fw, fb = _f32(fc_w), _f32(fc_b)
        _check(_lib.gasr_asr_set_weights(self._h, arrs[0], arrs[1], arrs[2], arrs[3], _fp(fw), _fp(fb)))

    def _result(self):
        ml = self.cfg.max_len
        return ([bytes(self._paths[i, : min(int(self._lens[i]), ml)]) for i in range(self._paths.shape[0])],
                [float(s) for s in self._scores])

    def run_host(self, x_host):
        assert x_host.dtype == np.float32 and x_host.flags["C_CONTIGUOUS"]
        _check(_lib.gasr_asr_run_host(self._h, _fp(x_host), self._paths.ctypes.data_as(ctypes.c_char_p),
                                      self._lens.ctypes.data_as(c_int_p), _fp(self._scores)))
        return self._result()

    def run_device(self, x_dev):
        _check(_lib.gasr_asr_run_device(self._h, x_dev, self._paths.ctypes.data_as(ctypes.c_char_p),
                                        self._lens.ctypes.data_as(c_int_p), _fp(self._scores)))
        return self._result()

    def submit_host(self, x_host):
        assert x_host.dtype == np.float32 and x_host.flags["C_CONTIGUOUS"]
        self._keep = x_host
        _check(_lib.gasr_asr_submit_host(self._h, _fp(x_host)))

    def submit_device(self, x_dev):
        _check(_lib.gasr_asr_submit_device(self._h, x_dev))

    def collect(self):
        _check(_lib.gasr_asr_collect(self._h, self._paths.ctypes.data_as(ctypes.c_char_p),
                                     self._lens.ctypes.data_as(c_int_p), _fp(self._scores)))
        return self._result()

    def set_lengths(self, lens):
        """Frames per utterance of the batches that follow (baseline/main.py:45 out_lens); None = all T."""
        if lens is None:
            _check(_lib.gasr_asr_set_lengths(self._h, None))
        else:
            a = np.ascontiguousarray(lens, dtype=np.int32)
            assert a.shape == (self.cfg.N,)
            _check(_lib.gasr_asr_set_lengths(self._h, a.ctypes.data_as(c_int_p)))

    def enable_timesteps(self, on=True):
        _check(_lib.gasr_asr_enable_timesteps(self._h, int(on)))

    def timesteps(self):
        """Per-token timesteps of the last run: [N][len] (nbest = 1) or [N][nbest][len]."""
        c = self.cfg
        ts = np.zeros((c.N, c.nbest, max(c.max_len, 1)), dtype=np.int32)
        _check(_lib.gasr_asr_timesteps(self._h, ts.ctypes.data_as(c_int_p)))
        counts = np.full((c.N,), c.nbest, dtype=np.int32)
        return _unpack_ts(ts, self._lens.reshape(c.N, c.nbest), counts, c.nbest, c.max_len)

    def profile(self, on=True):
        _check(_lib.gasr_asr_profile(self._h, int(on)))

    def last_ms(self):
        ms = ctypes.c_float(0)
        _check(_lib.gasr_asr_last_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def logprobs(self):
        p, ld = ctypes.c_void_p(), ctypes.c_int(0)
        _check(_lib.gasr_asr_logprobs(self._h, ctypes.byref(p), ctypes.byref(ld)))
        rows = self.cfg.T * self.cfg.N
        return self.ctx.to_host(p, (rows, ld.value))[:, : self.cfg.V]

    def stage_times(self):
        ms = (ctypes.c_float * 4)()
        _check(_lib.gasr_asr_stage_times(self._h, ms))
        return list(ms)

    def stage_launches(self):
        n = (ctypes.c_int * 4)()
        chunk = ctypes.c_int(0)
        _check(_lib.gasr_asr_stage_launches(self._h, n, ctypes.byref(chunk)))
        return list(n), chunk.value


class Job:
    """gasr_job: many batches of cfg.N utterances on one GPU, `lanes` of them in flight (BASELINE.json cfg5)."""

    def __init__(self, device, T, N, in_, H, L, V, beam, blank, vocab, lanes=2, precision=PREC_FP32, nbest=1, max_len=None):
        self.cfg = AsrConfig(CELL_TANH, 0, T, N, in_, H, L, V, beam, blank, precision, nbest, T + 1 if max_len is None else max_len)
        self._h = ctypes.c_void_p()
        _check(_lib.gasr_job_create(device, ctypes.byref(self.cfg), bytes(vocab), lanes, ctypes.byref(self._h)))
        self.lanes = lanes

    def close(self):
        if self._h:
            _lib.gasr_job_destroy(self._h)
            self._h = ctypes.c_void_p()

    def set_weights(self, w_ih, w_hh, b_ih, b_hh, fc_w, fc_b):
        keep = [[_f32(a) for a in lst] for lst in (w_ih, w_hh, b_ih, b_hh)]
        arrs = [(c_float_p * len(lst))(*[_fp(a) for a in lst]) for lst in keep]
        fw, fb = _f32(fc_w), _f32(fc_b)
        _check(_lib.gasr_job_set_weights(self._h, arrs[0], arrs[1], arrs[2], arrs[3], _fp(fw), _fp(fb)))

    def lane_context(self, lane=0):
        """A borrowed Context view of a lane (allocation / copies on the job's device); do not close it."""
        c = ctypes.c_void_p()
        _check(_lib.gasr_job_lane(self._h, lane, ctypes.byref(c), None))
        ctx = Context.__new__(Context)
        ctx._h = c
        ctx._borrowed = True
        return ctx

    def _run(self, fn, ptrs):
        n = len(ptrs)
        per = self.cfg.N * self.cfg.nbest
        ml = max(self.cfg.max_len, 1)
        paths = np.zeros((n * per, ml), dtype=np.uint8)
        lens = np.zeros((n * per,), dtype=np.int32)
        scores = np.zeros((n * per,), dtype=np.float32)
        arr = (ctypes.c_void_p * n)(*ptrs)
        _check(fn(self._h, arr, n, paths.ctypes.data_as(ctypes.c_char_p), lens.ctypes.data_as(c_int_p), _fp(scores)))
        return paths, lens, scores

    def run_host(self, batches):
        """batches: list of C-contiguous float32 arrays [T*N, in] (or raw host addresses)."""
        ptrs = [b if isinstance(b, int) else b.ctypes.data for b in batches]
        return self._run(_lib.gasr_job_run_host, ptrs)

    def run_device(self, dev_ptrs):
        ptrs = [p.value if isinstance(p, ctypes.c_void_p) else int(p) for p in dev_ptrs]
        return self._run(_lib.gasr_job_run_device, ptrs)

    def last_ms(self):
        ms = ctypes.c_float(0)
        _check(_lib.gasr_job_last_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        n = ctypes.c_longlong(0)
        _check(_lib.gasr_job_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def profile(self, on=True):
        _check(_lib.gasr_job_profile(self._h, int(on)))

    def stage_times(self):
        ms, n = (ctypes.c_float * 4)(), (ctypes.c_int * 4)()
        _check(_lib.gasr_job_stage_times(self._h, ms, n))
        return list(ms), list(n)

    def lane_logprobs(self, lane=0):
        """Log-probabilities [T*N, V] of the last batch that ran on a lane."""
        c, a = ctypes.c_void_p(), ctypes.c_void_p()
        _check(_lib.gasr_job_lane(self._h, lane, ctypes.byref(c), ctypes.byref(a)))
        p, ld = ctypes.c_void_p(), ctypes.c_int(0)
        _check(_lib.gasr_asr_logprobs(a, ctypes.byref(p), ctypes.byref(ld)))
        return self.lane_context(lane).to_host(p, (self.cfg.T * self.cfg.N, ld.value))[:, : self.cfg.V]


def unpack_results(paths, lens, scores, max_len):
    return ([bytes(paths[i, : min(int(lens[i]), max_len)]) for i in range(paths.shape[0])], [float(s) for s in scores])


def _declare():
    vp, sz, ci, cf = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
    L = _lib
    L.gasr_ctx_create.argtypes = [ci, c_void_pp]
    L.gasr_ctx_destroy.argtypes = [vp]
    L.gasr_ctx_sync.argtypes = [vp]
    L.gasr_ctx_sm_count.argtypes = [vp, c_int_p]
    L.gasr_timer_start.argtypes = [vp]
    L.gasr_timer_stop.argtypes = [vp, c_float_p]
    L.gasr_ctx_launch_count.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
    L.gasr_ctc_last_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    L.gasr_malloc_device.argtypes = [vp, sz, c_void_pp]
    L.gasr_free_device.argtypes = [vp, vp]
    L.gasr_malloc_host.argtypes = [vp, sz, c_void_pp]
    L.gasr_free_host.argtypes = [vp, vp]
    L.gasr_memcpy_h2d_on_stream.argtypes = [vp, vp, vp, sz, vp]
    for f in (L.gasr_memcpy_h2d, L.gasr_memcpy_d2h, L.gasr_memcpy_h2d_async, L.gasr_memcpy_d2h_async):
        f.argtypes = [vp, vp, vp, sz]
    L.gasr_memset_device.argtypes = [vp, vp, ci, sz]
    L.gasr_memory_stats.argtypes = [vp, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    L.gasr_matmul.argtypes = [vp, vp, ci, ci, vp, ci, ci, vp, ci, ci, ci, ci]
    L.gasr_matadd.argtypes = [vp, vp, ci, vp, ci, vp, ci, ci, ci, cf]
    L.gasr_linear_forward.argtypes = [vp, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci]
    L.gasr_xproj_gemm.argtypes = [vp, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci]
    L.gasr_log_softmax.argtypes = [vp, vp, ci, vp, ci, ci, ci]
    L.gasr_rnn_cell_forward.argtypes = [vp] + [vp] * 7 + [ci, ci, ci]
    L.gasr_rnn_forward.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, ci]
    L.gasr_ctc_decode.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, ctypes.c_char_p, ci, ci, ctypes.c_char_p,
                                  c_int_p, c_float_p, c_int_p]
    L.gasr_ctc_decode_host.argtypes = [vp, c_float_p, ci, ci, ci, ci, ci, ci, ctypes.c_char_p, ci, ci,
                                       ctypes.c_char_p, c_int_p, c_float_p, c_int_p]
    L.gasr_asr_create.argtypes = [vp, ctypes.POINTER(AsrConfig), ctypes.c_char_p, c_void_pp]
    L.gasr_asr_destroy.argtypes = [vp]
    L.gasr_asr_set_weights.argtypes = [vp, vp, vp, vp, vp, c_float_p, c_float_p]
    L.gasr_asr_run_host.argtypes = [vp, c_float_p, ctypes.c_char_p, c_int_p, c_float_p]
    L.gasr_asr_run_device.argtypes = [vp, vp, ctypes.c_char_p, c_int_p, c_float_p]
    L.gasr_asr_logprobs.argtypes = [vp, c_void_pp, c_int_p]
    L.gasr_asr_stage_times.argtypes = [vp, c_float_p]
    L.gasr_asr_stage_launches.argtypes = [vp, c_int_p, c_int_p]
    L.gasr_asr_submit_host.argtypes = [vp, c_float_p]
    L.gasr_asr_submit_device.argtypes = [vp, vp]
    L.gasr_asr_collect.argtypes = [vp, ctypes.c_char_p, c_int_p, c_float_p]
    L.gasr_asr_profile.argtypes = [vp, ci]
    L.gasr_asr_last_ms.argtypes = [vp, c_float_p]
    L.gasr_job_create.argtypes = [ci, ctypes.POINTER(AsrConfig), ctypes.c_char_p, ci, c_void_pp]
    L.gasr_job_destroy.argtypes = [vp]
    L.gasr_job_set_weights.argtypes = [vp, vp, vp, vp, vp, c_float_p, c_float_p]
    L.gasr_job_run_host.argtypes = [vp, c_void_pp, ci, ctypes.c_char_p, c_int_p, c_float_p]
    L.gasr_job_run_device.argtypes = [vp, c_void_pp, ci, ctypes.c_char_p, c_int_p, c_float_p]
    L.gasr_job_last_ms.argtypes = [vp, c_float_p]
    L.gasr_job_launch_count.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
    L.gasr_job_lane.argtypes = [vp, ci, c_void_pp, c_void_pp]
    L.gasr_job_profile.argtypes = [vp, ci]
    L.gasr_job_stage_times.argtypes = [vp, c_float_p, c_int_p]
    L.gasr_synth_spectrogram.argtypes = [vp, vp, ctypes.c_ulonglong, ci, ci, ci, ctypes.c_longlong]


_declare()
