// capi.cu -- the extern "C" boundary declared in include/gasr.h: context, memory, and the module entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <chrono>
#include <vector>

#include "asr.cuh"
#include "common.cuh"
#include "rnn_stream.cuh"
#include "stream.cuh"
#include "tc_common.cuh"

namespace gasr {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ws_reserve(gasr_ctx *ctx, Workspace &ws, size_t bytes) {
    if (bytes <= ws.bytes) return GASR_OK;
    if (ws.ptr) {
        GASR_CUDA(cudaStreamSynchronize(ctx->stream));
        GASR_CUDA(cudaFree(ws.ptr));
        ctx->device_bytes -= ws.bytes;
        ws.ptr = nullptr; ws.bytes = 0;
    }
    bytes = align_up(bytes, 1 << 20);
    cudaError_t e = cudaMalloc(&ws.ptr, bytes);
    if (e != cudaSuccess) {
        set_error("workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        ws.ptr = nullptr;
        return GASR_ERR_NOMEM;
    }
    ws.bytes = bytes;
    ctx->device_bytes += bytes;
    return GASR_OK;
}

int pinned_reserve(gasr_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_out_bytes) return GASR_OK;
    if (ctx->pinned_out) {
        GASR_CUDA(cudaStreamSynchronize(ctx->stream));
        GASR_CUDA(cudaFreeHost(ctx->pinned_out));
        ctx->host_bytes -= ctx->pinned_out_bytes;
        ctx->pinned_out = nullptr; ctx->pinned_out_bytes = 0;
    }
    bytes = align_up(bytes, 1 << 16);
    cudaError_t e = cudaHostAlloc(&ctx->pinned_out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        set_error("pinned staging allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        ctx->pinned_out = nullptr;
        return GASR_ERR_NOMEM;
    }
    ctx->pinned_out_bytes = bytes;
    ctx->host_bytes += bytes;
    return GASR_OK;
}

}  // namespace gasr

using namespace gasr;

extern "C" {

int gasr_version(void) { return 100; }
const char *gasr_last_error(void) { return g_err; }

int gasr_device_count(int *count) {
    GASR_CHECK(count != nullptr, "gasr_device_count: null output");
    *count = 0;
    GASR_CUDA(cudaGetDeviceCount(count));
    return GASR_OK;
}

int gasr_ctx_create(int device, gasr_ctx **out) {
    GASR_CHECK(out != nullptr, "gasr_ctx_create: null output");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return GASR_ERR_CUDA;
    }
    GASR_CHECK(device >= 0 && device < n, "gasr_ctx_create: device %d out of range (0..%d)", device, n - 1);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    GASR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return GASR_ERR_UNSUPPORTED;
    }
    gasr_ctx *ctx = new gasr_ctx();
    {   // environment switches are read here and nowhere else
        gasr_options &o = ctx->opt;
        auto chr = [](const char *n) -> char { const char *e = getenv(n); return e ? e[0] : (char)0; };
        auto num = [](const char *n, int dflt) -> int { const char *e = getenv(n); return e ? atoi(e) : dflt; };
        o.rnn = chr("GASR_RNN"); o.rnn_mc = num("GASR_RNN_MC", 1); o.rnn_groups = num("GASR_RNN_G", 0); o.rnn_pair = num("GASR_RNN_PAIR", 1);
        o.ctc_kernel = chr("GASR_CTC_KERNEL"); o.ctc_mw = num("GASR_CTC_MW", 8); o.ctc_cells = num("GASR_CTC_CELLS", 0); o.ctc_pad = num("GASR_CTC_PAD", 0);
        o.gru = chr("GASR_GRU"); o.gru_pp = num("GASR_GRU_PP", 0); o.gru_units = num("GASR_GRU_UNITS", 0); o.gru_no_pdl = getenv("GASR_GRU_NO_PDL") != nullptr; o.no_graph = getenv("GASR_NO_GRAPH") != nullptr;
        o.bidir_serial = getenv("GASR_BIDIR_SERIAL") != nullptr; o.linear_simt = getenv("GASR_LINEAR_SIMT") != nullptr;
        o.xproj = chr("GASR_XPROJ"); o.chunk = num("GASR_CHUNK", -1); o.stream = num("GASR_STREAM", -1); o.wave = num("GASR_WAVE", -1);
        o.stream_gemm_ctas = num("GASR_STREAM_GEMM_CTAS", 24);
        o.gemm_stages = num("GASR_GEMM_STAGES", 3); o.ctc_warps = num("GASR_CTC_WARPS", 8); o.gemm_bn = num("GASR_GEMM_BN", 256); o.gemm_pair = num("GASR_GEMM_PAIR", 1);
        o.wave_serial = getenv("GASR_WAVE_SERIAL") != nullptr; o.wave_prio = num("GASR_WAVE_PRIO", -1); o.wave_timeout_s = num("GASR_WAVE_TIMEOUT_S", 60);
    }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    int cl = 0;
    cudaDeviceGetAttribute(&cl, cudaDevAttrClusterLaunch, device);
    ctx->cluster_ok = cl;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_start) != cudaSuccess || cudaEventCreate(&ctx->ev_stop) != cudaSuccess) {
        set_error("gasr_ctx_create: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return GASR_ERR_CUDA;
    }
    for (int i = 0; i < 6; i++) cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking);
    *out = ctx;
    return GASR_OK;
}

int gasr_ctx_destroy(gasr_ctx *ctx) {
    GASR_ENTER(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->dev_blocks) cudaFree(kv.first);
    for (auto &kv : ctx->host_blocks) cudaFreeHost(kv.first);
    Workspace *wss[] = {&ctx->ws_ctc, &ctx->ws_rnn, &ctx->ws_misc, &ctx->ws_out, &ctx->ws_gru, &ctx->ws_lin, &ctx->ws_lens, &ctx->ws_rnn_b, &ctx->ws_misc_b, &ctx->ws_gru_b, &ctx->ws_wide};
    for (cudaEvent_t e : ctx->ev_bi) if (e) cudaEventDestroy(e);
    for (auto &g : ctx->step_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (Workspace *w : wss) if (w->ptr) cudaFree(w->ptr);
    if (ctx->pinned_out) cudaFreeHost(ctx->pinned_out);
    for (int i = 0; i < 6; i++) if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    cudaEventDestroy(ctx->ev_start);
    cudaEventDestroy(ctx->ev_stop);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GASR_OK;
}

int gasr_ctx_sync(gasr_ctx *ctx) {
    GASR_ENTER(ctx);
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    return GASR_OK;
}

int gasr_ctx_sm_count(gasr_ctx *ctx, int *sms) {
    GASR_ENTER(ctx);
    GASR_CHECK(sms != nullptr, "null output");
    *sms = ctx->sm_count;
    return GASR_OK;
}

int gasr_timer_start(gasr_ctx *ctx) {
    GASR_ENTER(ctx);
    GASR_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
    return GASR_OK;
}

int gasr_timer_stop(gasr_ctx *ctx, float *ms) {
    GASR_ENTER(ctx);
    GASR_CHECK(ms != nullptr, "null output");
    GASR_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
    GASR_CUDA(cudaEventSynchronize(ctx->ev_stop));
    GASR_CUDA(cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
    return GASR_OK;
}

int gasr_ctc_last_stats(gasr_ctx *ctx, long long *fallback_frames, long long *survivors) {
    GASR_ENTER(ctx);
    if (fallback_frames) *fallback_frames = ctx->ctc_fallback_frames;
    if (survivors) *survivors = ctx->ctc_survivors;
    return GASR_OK;
}

int gasr_ctx_launch_count(gasr_ctx *ctx, long long *launches) {
    GASR_ENTER(ctx);
    GASR_CHECK(launches != nullptr, "null output");
    *launches = ctx->launches;
    return GASR_OK;
}

/* ---- memory ---------------------------------------------------------------------------------------- */
int gasr_malloc_device(gasr_ctx *ctx, size_t bytes, void **ptr) {
    GASR_ENTER(ctx);
    GASR_CHECK(ptr != nullptr, "gasr_malloc_device: null output");
    *ptr = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) {
        set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        *ptr = nullptr;
        return GASR_ERR_NOMEM;
    }
    GASR_CUDA(cudaMemsetAsync(*ptr, 0, bytes, ctx->stream));
    ctx->dev_blocks[*ptr] = bytes;
    ctx->device_bytes += bytes;
    return GASR_OK;
}

int gasr_free_device(gasr_ctx *ctx, void *ptr) {
    GASR_ENTER(ctx);
    auto it = ctx->dev_blocks.find(ptr);
    if (it == ctx->dev_blocks.end()) return GASR_OK;   // MemoryMonitor::freeGpuMemory ignores unknown pointers
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    GASR_CUDA(cudaFree(ptr));
    ctx->device_bytes -= it->second;
    ctx->dev_blocks.erase(it);
    return GASR_OK;
}

int gasr_malloc_host(gasr_ctx *ctx, size_t bytes, void **ptr) {
    GASR_ENTER(ctx);
    GASR_CHECK(ptr != nullptr, "gasr_malloc_host: null output");
    *ptr = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);   // MemoryMonitor.cpp:12
    if (e != cudaSuccess) {
        set_error("pinned host allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        *ptr = nullptr;
        return GASR_ERR_NOMEM;
    }
    memset(*ptr, 0, bytes);
    ctx->host_blocks[*ptr] = bytes;
    ctx->host_bytes += bytes;
    return GASR_OK;
}

int gasr_free_host(gasr_ctx *ctx, void *ptr) {
    GASR_ENTER(ctx);
    auto it = ctx->host_blocks.find(ptr);
    if (it == ctx->host_blocks.end()) return GASR_OK;
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    GASR_CUDA(cudaFreeHost(ptr));
    ctx->host_bytes -= it->second;
    ctx->host_blocks.erase(it);
    return GASR_OK;
}

int gasr_memcpy_h2d(gasr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    GASR_ENTER(ctx);
    GASR_CHECK(bytes == 0 || (dst && src), "gasr_memcpy_h2d: null pointer");
    if (bytes == 0) return GASR_OK;
    GASR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    return GASR_OK;
}

int gasr_memcpy_d2h(gasr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    GASR_ENTER(ctx);
    GASR_CHECK(bytes == 0 || (dst && src), "gasr_memcpy_d2h: null pointer");
    if (bytes == 0) return GASR_OK;
    GASR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    return GASR_OK;
}

int gasr_memcpy_h2d_async(gasr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    GASR_ENTER(ctx);
    GASR_CHECK(bytes == 0 || (dst && src), "gasr_memcpy_h2d_async: null pointer");
    if (bytes == 0) return GASR_OK;
    GASR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return GASR_OK;
}

int gasr_memcpy_h2d_on_stream(gasr_ctx *ctx, void *dst, const void *src, size_t bytes, void *cuda_stream) {
    GASR_ENTER(ctx);
    GASR_CHECK(bytes == 0 || (dst && src), "gasr_memcpy_h2d_on_stream: null pointer");
    if (bytes == 0) return GASR_OK;
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    GASR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return GASR_OK;
}

int gasr_memcpy_d2h_async(gasr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    GASR_ENTER(ctx);
    GASR_CHECK(bytes == 0 || (dst && src), "gasr_memcpy_d2h_async: null pointer");
    if (bytes == 0) return GASR_OK;
    GASR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return GASR_OK;
}

int gasr_memset_device(gasr_ctx *ctx, void *dst, int value, size_t bytes) {
    GASR_ENTER(ctx);
    GASR_CHECK(bytes == 0 || dst, "gasr_memset_device: null pointer");
    if (bytes == 0) return GASR_OK;
    GASR_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return GASR_OK;
}

int gasr_memory_stats(gasr_ctx *ctx, size_t *device_bytes, size_t *host_bytes) {
    GASR_ENTER(ctx);
    if (device_bytes) *device_bytes = ctx->device_bytes;
    if (host_bytes) *host_bytes = ctx->host_bytes;
    return GASR_OK;
}

/* ---- dense math -------------------------------------------------------------------------------------- */
int gasr_matmul(gasr_ctx *ctx, const float *x, int ldx, int trans_x, const float *y, int ldy, int trans_y, float *z,
                int ldz, int m, int k, int n) {
    GASR_ENTER(ctx);
    return launch_matmul(ctx, x, ldx, trans_x, y, ldy, trans_y, z, ldz, m, k, n, nullptr, ctx->stream);
}

int gasr_matadd(gasr_ctx *ctx, const float *x, int ldx, const float *y, int ldy, float *z, int ldz, int rows, int cols,
                float lambda) {
    GASR_ENTER(ctx);
    return launch_matadd(ctx, x, ldx, y, ldy, z, ldz, rows, cols, lambda, ctx->stream);
}

int gasr_xproj_gemm(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *bias, float *y, int ldy,
                    int rows, int in, int out, int precision) {
    GASR_ENTER(ctx);
    GASR_CHECK(x && W && y && rows >= 0 && in >= 1 && out >= 1 && ldx >= in && ldy >= out, "xproj_gemm: bad arguments");
    GASR_CHECK(precision == GASR_PREC_FP32 || precision == GASR_PREC_BF16, "xproj_gemm: unknown precision");
    if (rows == 0) return GASR_OK;
    if (!xproj_tc_supported(rows, in, out) || ldy % 4 != 0)
        return launch_matmul(ctx, x, ldx, 0, W, out, 0, y, ldy, rows, in, out, bias, ctx->stream);
    const size_t wb = xproj_tc_w_bytes(in, out), ab = xproj_tc_a_bytes(rows, in);
    GASR_TRY(ws_reserve(ctx, ctx->ws_misc, wb + ab + 2048));
    unsigned char *base = static_cast<unsigned char *>(ctx->ws_misc.ptr);
    GASR_TRY(xproj_tc_prepare_weights(ctx, W, in, out, base, ctx->stream));
    return launch_xproj_tc(ctx, x, ldx, rows, in, out, base, base + align_up(wb, 1024), bias, y, ldy, precision, ctx->stream);
}

int gasr_linear_forward(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *b, float *y, int ldy,
                        int rows, int in, int out, int act) {
    GASR_ENTER(ctx);
    return launch_linear(ctx, x, ldx, W, b, y, ldy, rows, in, out, act, ctx->stream);
}

int gasr_log_softmax(gasr_ctx *ctx, const float *x, int ldx, float *y, int ldy, int rows, int cols) {
    GASR_ENTER(ctx);
    return launch_log_softmax(ctx, x, ldx, y, ldy, rows, cols, ctx->stream);
}

int gasr_rnn_cell_forward(gasr_ctx *ctx, const float *x, const float *h_prev, const float *w_ih, const float *w_hh,
                          const float *b_ih, const float *b_hh, float *out, int batch, int in, int hidden) {
    GASR_ENTER(ctx);
    return launch_rnn_cell(ctx, x, h_prev, w_ih, w_hh, b_ih, b_hh, out, batch, in, hidden, ctx->stream);
}

}  // extern "C"

namespace gasr {



// One layer, one direction: xproj = src*W_ih + bias (all timesteps), then the persistent recurrence.
static int rnn_layer_direction(gasr_ctx *ctx, int cell, int T, int N, int in_l, int H, const float *src, int ld_src,
                               const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh, int reverse,
                               float *out, int ldo, int col0, int precision, float *xproj, float *bias,
                               cudaStream_t st, StageEvents *prof, bool concurrent = false) {
    const int G = cell == GASR_CELL_GRU ? 3 : 1;
    if (cell == GASR_CELL_TANH) {
        GASR_TRY(launch_matadd(ctx, b_ih, H, b_hh, H, bias, H, 1, H, 1.0f, st));   // (b_hh + b_ih), RNN_Cell.cu:10
    } else {
        GASR_CUDA(cudaMemcpyAsync(bias, b_ih, sizeof(float) * G * H, cudaMemcpyDeviceToDevice, st));
    }
    const bool use_tc = ctx->opt.xproj != 's' && ld_src == in_l && xproj_tc_supported(T * N, in_l, G * H);
    if (use_tc) {
        // tensor-core path: W^T and A are split into bf16 hi/lo planes in the misc workspace
        const size_t wb = xproj_tc_w_bytes(in_l, G * H), ab = xproj_tc_a_bytes(T * N, in_l);
        Workspace &wsm = ctx->ws_sel ? ctx->ws_misc_b : ctx->ws_misc;
        GASR_TRY(ws_reserve(ctx, wsm, wb + ab + 2048));
        unsigned char *base = static_cast<unsigned char *>(wsm.ptr);
        GASR_TRY(xproj_tc_prepare_weights(ctx, w_ih, in_l, G * H, base, st));
        GASR_TRY(launch_xproj_tc(ctx, src, ld_src, T * N, in_l, G * H, base, base + align_up(wb, 1024), bias, xproj, G * H,
                                 precision, st));
    } else {
        GASR_TRY(launch_matmul(ctx, src, ld_src, 0, w_ih, G * H, 0, xproj, G * H, T * N, in_l, G * H, bias, st));
    }
    if (prof) GASR_TRY(prof->mark(0, st));
    RnnLayerArgs a;
    a.cell = cell; a.T = T; a.N = N; a.H = H; a.reverse = reverse;
    a.xproj = xproj; a.ldxp = G * H; a.w_hh = w_hh; a.b_hh = b_hh; a.out = out; a.ldo = ldo; a.col0 = col0;
    a.precision = precision; a.concurrent = concurrent;
    GASR_TRY(launch_rnn_recurrence(ctx, a, st));
    if (prof) GASR_TRY(prof->mark(1, st));
    return GASR_OK;
}

int rnn_forward_impl(gasr_ctx *ctx, int cell, int bidir, int T, int N, int in, int H, int L, const float *const *w_ih,
                     const float *const *w_hh, const float *const *b_ih, const float *const *b_hh, const float *x,
                     float *const *hiddens, int precision, cudaStream_t st, StageEvents *prof) {
    GASR_CHECK(cell == GASR_CELL_TANH || cell == GASR_CELL_GRU, "rnn_forward: unknown cell type %d", cell);
    GASR_CHECK(T >= 0 && N >= 0 && in >= 1 && H >= 1 && L >= 1, "rnn_forward: bad shape");
    GASR_CHECK(w_ih && w_hh && b_ih && b_hh && x && hiddens, "rnn_forward: null argument");
    GASR_CHECK((long long)T * N < (1ll << 31) / 4, "rnn_forward: T*N too large");
    const int D = bidir ? 2 : 1, G = cell == GASR_CELL_GRU ? 3 : 1;
    if (T == 0 || N == 0) return GASR_OK;
    const size_t xp_bytes = align_up(sizeof(float) * (size_t)T * N * G * H, 256);
    GASR_TRY(ws_reserve(ctx, ctx->ws_rnn, xp_bytes + align_up(sizeof(float) * G * H, 256)));
    // the two directions of a bidirectional layer are independent (same input, disjoint output columns): the backward one
    // runs on a side stream with its own workspaces
    const bool fork = D == 2 && st == ctx->stream && !ctx->opt.bidir_serial;
    if (fork) {
        GASR_TRY(ws_reserve(ctx, ctx->ws_rnn_b, xp_bytes + align_up(sizeof(float) * G * H, 256)));
        for (cudaEvent_t &e : ctx->ev_bi)
            if (e == nullptr) GASR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t st_b = fork ? ctx->side[5] : st;
    int rc = GASR_OK;
    for (int l = 0; l < L && rc == GASR_OK; l++) {
        const int in_l = l == 0 ? in : D * H;
        const float *src = l == 0 ? x : hiddens[l - 1];
        if (fork) {
            GASR_CUDA(cudaEventRecord(ctx->ev_bi[0], st));
            GASR_CUDA(cudaStreamWaitEvent(st_b, ctx->ev_bi[0], 0));
        }
        for (int d = 0; d < D && rc == GASR_OK; d++) {
            const int i = l * D + d;
            if (!(w_ih[i] && w_hh[i] && b_ih[i] && b_hh[i] && hiddens[l])) { set_error("rnn_forward: null layer parameter"); rc = GASR_ERR_INVALID; break; }
            const bool side = fork && d == 1;
            Workspace &wr = side ? ctx->ws_rnn_b : ctx->ws_rnn;
            float *xproj = static_cast<float *>(wr.ptr);
            float *bias = reinterpret_cast<float *>(static_cast<unsigned char *>(wr.ptr) + xp_bytes);
            ctx->ws_sel = side ? 1 : 0;
            rc = rnn_layer_direction(ctx, cell, T, N, in_l, H, src, in_l, w_ih[i], w_hh[i], b_ih[i], b_hh[i], d,
                                     hiddens[l], D * H, d * H, precision, xproj, bias, side ? st_b : st, side ? nullptr : prof, fork);
            ctx->ws_sel = 0;
        }
        if (fork && rc == GASR_OK) {
            GASR_CUDA(cudaEventRecord(ctx->ev_bi[1], st_b));
            GASR_CUDA(cudaStreamWaitEvent(st, ctx->ev_bi[1], 0));
        }
    }
    if (rc != GASR_OK && fork) {
        // a failed layer must not leave the other direction's work running on the side stream behind the caller's back
        cudaStreamSynchronize(st_b);
        cudaStreamSynchronize(st);
        cudaGetLastError();
    }
    return rc;
}

}  // namespace gasr

extern "C" {

int gasr_rnn_forward(gasr_ctx *ctx, int cell, int bidirectional, int T, int N, int in, int H, int L,
                     const float *const *w_ih, const float *const *w_hh, const float *const *b_ih,
                     const float *const *b_hh, const float *x, float *const *hiddens, int precision) {
    GASR_ENTER(ctx);
    return rnn_forward_impl(ctx, cell, bidirectional, T, N, in, H, L, w_ih, w_hh, b_ih, b_hh, x, hiddens, precision,
                            ctx->stream);
}

int gasr_ctc_decode_ex(gasr_ctx *ctx, const float *scores, int domain, int T, int N, int V, int ld, int beam, int blank,
                       const char *vocab, int max_len, int nbest, const int *lens_host, char *out_paths, int *out_lens,
                       float *out_scores, int *out_counts, int *out_timesteps) {
    GASR_ENTER(ctx);
    GASR_CHECK(out_paths && out_lens && out_scores, "ctc_decode: null output buffer");
    CtcArgs a = {scores, domain, T, N, V, ld, beam, blank, vocab, max_len, nbest, out_paths, out_lens, out_scores, out_counts};
    a.out_timesteps = out_timesteps;
    if (lens_host != nullptr && N > 0) {
        for (int n = 0; n < N; n++)
            GASR_CHECK(lens_host[n] >= 1 && lens_host[n] <= T, "ctc_decode: length %d of utterance %d outside 1..T=%d", lens_host[n], n, T);
        GASR_TRY(ws_reserve(ctx, ctx->ws_lens, sizeof(int) * (size_t)N));
        GASR_CUDA(cudaMemcpyAsync(ctx->ws_lens.ptr, lens_host, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
        a.lens_dev = static_cast<const int *>(ctx->ws_lens.ptr);
    }
    GASR_TRY(ctc_decode_launch(ctx, a, ctx->stream));
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    return ctc_decode_finish(ctx, a);
}

int gasr_ctc_decode(gasr_ctx *ctx, const float *scores, int domain, int T, int N, int V, int ld, int beam, int blank,
                    const char *vocab, int max_len, int nbest, char *out_paths, int *out_lens, float *out_scores,
                    int *out_counts) {
    return gasr_ctc_decode_ex(ctx, scores, domain, T, N, V, ld, beam, blank, vocab, max_len, nbest, nullptr, out_paths, out_lens,
                              out_scores, out_counts, nullptr);
}

int gasr_ctc_decode_host(gasr_ctx *ctx, const float *scores_host, int domain, int T, int N, int V, int beam, int blank,
                         const char *vocab, int max_len, int nbest, char *out_paths, int *out_lens, float *out_scores,
                         int *out_counts) {
    GASR_ENTER(ctx);
    GASR_CHECK(scores_host != nullptr && T >= 1 && N >= 0 && V >= 1, "ctc_decode_host: bad arguments");
    const size_t bytes = sizeof(float) * (size_t)T * N * V;
    GASR_TRY(ws_reserve(ctx, ctx->ws_misc, bytes + 256));
    GASR_CUDA(cudaMemcpyAsync(ctx->ws_misc.ptr, scores_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return gasr_ctc_decode(ctx, static_cast<const float *>(ctx->ws_misc.ptr), domain, T, N, V, V, beam, blank, vocab,
                           max_len, nbest, out_paths, out_lens, out_scores, out_counts);
}

}  // extern "C"
