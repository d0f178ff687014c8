// asr.cu -- the fused pipeline object behind gasr_asr_* (include/gasr.h): creation and mode selection, weights, and the three
// older execution modes (sequential, time-chunked, streaming); the throughput engine is asr_wave.cu, many batches per GPU job.cu.
// Stands behind the reference's drivers (main.cpp:31-72, baseline/main.py:36-52): acoustic model forward + decode in one call.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "asr.cuh"
#include "common.cuh"
#include "rnn_stream.cuh"
#include "stream.cuh"
#include "tc_common.cuh"

using namespace gasr;

/* ---- fused pipeline ------------------------------------------------------------------------------------ */


extern "C" {

int gasr_asr_create(gasr_ctx *ctx, const gasr_asr_config *cfg, const char *vocab, gasr_asr **out) {
    GASR_ENTER(ctx);
    GASR_CHECK(cfg && vocab && out, "gasr_asr_create: null argument");
    GASR_CHECK(cfg->T >= 1 && cfg->N >= 1 && cfg->in >= 1 && cfg->H >= 1 && cfg->L >= 1 && cfg->V >= 1 &&
                   cfg->beam >= 1 && cfg->blank >= 0 && cfg->blank < cfg->V && cfg->nbest >= 1 && cfg->max_len >= 0,
               "gasr_asr_create: bad configuration");
    GASR_CHECK(cfg->cell == GASR_CELL_TANH || cfg->cell == GASR_CELL_GRU, "gasr_asr_create: unknown cell");
    gasr_asr *a = new gasr_asr();
    a->ctx = ctx; a->cfg = *cfg;
    a->vocab.assign(vocab, vocab + cfg->V);
    a->D = cfg->bidirectional ? 2 : 1;
    a->G = cfg->cell == GASR_CELL_GRU ? 3 : 1;
    a->ldp = (cfg->V + 3) / 4 * 4;
    const int D = a->D, G = a->G, H = cfg->H;
    const size_t rows = (size_t)cfg->T * cfg->N;
    int st = GASR_OK;
    auto alloc = [&](float **p, size_t n) { if (st == GASR_OK) st = gasr_malloc_device(ctx, sizeof(float) * n, (void **)p); };
    a->w_ih.assign(cfg->L * D, nullptr); a->w_hh.assign(cfg->L * D, nullptr);
    a->b_ih.assign(cfg->L * D, nullptr); a->b_hh.assign(cfg->L * D, nullptr);
    a->hiddens.assign(cfg->L, nullptr);
    // Execution mode.  Default: the wave engine (asr_wave.cu) -- stream-ordered time chunks over groups of 128 utterances,
    // no kernel ever waits for another kernel.  GASR_STREAM=1 opts into the round-1 latency mode (three persistent kernels
    // coupled by progress counters: needs a whole idle GPU); GASR_WAVE=0 selects the older chunked path.
    const bool want_wave = ctx->opt.wave != 0 && ctx->opt.stream != 1 && wave_supported(ctx, *cfg);
    for (int l = 0; l < cfg->L; l++) {
        const int in_l = l == 0 ? cfg->in : D * H;
        for (int d = 0; d < D; d++) {
            alloc(&a->w_ih[l * D + d], (size_t)in_l * G * H);
            alloc(&a->w_hh[l * D + d], (size_t)H * G * H);
            alloc(&a->b_ih[l * D + d], (size_t)G * H);
            alloc(&a->b_hh[l * D + d], (size_t)G * H);
        }
        if (!want_wave) alloc(&a->hiddens[l], rows * D * H);
    }
    alloc(&a->fc_w, (size_t)D * H * cfg->V);
    alloc(&a->fc_b, (size_t)cfg->V);
    alloc(&a->x_dev, rows * cfg->in);
    if (want_wave) {
        a->ldp = 32;
        if (st == GASR_OK) st = wave_create(a);
        if (st != GASR_OK) { gasr_asr_destroy(a); return st; }
        *out = a;
        return GASR_OK;
    }
    alloc(&a->logp, rows * a->ldp);
    {
        // time-chunked pipelining needs the chunk-resumable kernels: cluster recurrence + warp decoder
        const int want = ctx->opt.chunk >= 0 ? ctx->opt.chunk : 50;
        const bool h_ok = (H == 64 || H == 128 || H == 256 || H == 512) && ctx->cluster_ok;
        if (want > 0 && cfg->cell == GASR_CELL_TANH && !cfg->bidirectional && h_ok && cfg->beam <= 32 && cfg->V <= 32 &&
            cfg->T >= 2 * want) {
            a->chunk = want;
            alloc(&a->xproj_all, (size_t)cfg->L * rows * H);
            alloc(&a->bias_all, (size_t)cfg->L * H);
            a->use_tc = ctx->opt.xproj != 's' && xproj_tc_supported((int)rows, cfg->in, H);
            if (a->use_tc) {
                a->tc_abuf.assign(cfg->L, nullptr); a->tc_wbuf.assign(cfg->L, nullptr);
                for (int l = 0; l < cfg->L && st == GASR_OK; l++) {
                    const int in_l = l == 0 ? cfg->in : H;
                    const int nchunks = ceil_div(cfg->T, want);
                    st = gasr_malloc_device(ctx, align_up(xproj_tc_a_bytes(want * cfg->N, in_l), 1024) * (size_t)nchunks + 1024,
                                            &a->tc_abuf[l]);
                    if (st == GASR_OK) st = gasr_malloc_device(ctx, xproj_tc_w_bytes(in_l, H) + 1024, &a->tc_wbuf[l]);
                }
            }
        }
    }
    if (st == GASR_OK && a->chunk > 0 && a->use_tc) {
        // streaming envelope: whole blocks of 128 rows, the persistent recurrence's cluster budget, one output tile
        const int want_stream = ctx->opt.stream == 1;          // opt-in: persistent kernels that wait for each other
        const int N = cfg->N, L = cfg->L;
        // Residency is computed, not assumed: every cluster of the recurrence kernel (whole SMs), the GEMM CTAs and one decoder
        // CTA per utterance must be co-resident, or the kernels would wait for each other until their watchdogs fire.
        bool fits = false;
        if (want_stream && (H == 512 || H == 256 || H == 128) && ctx->cluster_ok) {
            int max_clusters = 0, cs = 1;
            if (rnn_stream_max_clusters(ctx, H, &max_clusters, &cs) == GASR_OK) {
                const int rec_clusters = ceil_div(N, 16) * L;
                const int other_sms = ctx->sm_count - rec_clusters * cs;             // SMs left for the GEMM and decoder CTAs
                fits = rec_clusters <= max_clusters && other_sms >= ctx->opt.stream_gemm_ctas &&
                       N <= 2 * (other_sms - ctx->opt.stream_gemm_ctas) + ctx->opt.stream_gemm_ctas;   // two decoder CTAs per free SM, one next to a GEMM CTA
            }
            cudaGetLastError();
        }
        if (want_stream && fits && cfg->beam <= 32 && cfg->V <= 32 && N % 16 == 0 && 128 % N == 0 && L + 1 <= XS_MAX_TARGETS && L <= RS_MAX_LAYERS &&
            rnn_stream_supported(ctx, H, N, L) && H % 128 == 0 && ((size_t)cfg->T * N) % 128 == 0) {
            a->stream_fpb = 128 / N;
            a->stream_blocks = (int)(rows / 128);
            a->stream_gemm_ctas = ctx->opt.stream_gemm_ctas;
            auto allocv = [&](void **p, size_t bytes) { if (st == GASR_OK) st = gasr_malloc_device(ctx, bytes, p); };
            allocv(&a->x_planes, xproj_tc_a_bytes((int)rows, cfg->in) + 1024);
            a->h_planes.assign(L, nullptr);
            for (int l = 0; l < L; l++) allocv(&a->h_planes[l], xproj_tc_a_bytes((int)rows, H) + 1024);
            allocv(&a->fc_wbuf, xproj_tc_w_bytes(H, 32) + 1024);
            alloc(&a->fc_b_pad, 32);
            a->flags_bytes = sizeof(unsigned) * ((size_t)(2 * L + 2) * a->stream_blocks + 16);
            allocv((void **)&a->flags, a->flags_bytes);
            if (st == GASR_OK && cudaHostAlloc((void **)&a->host_words, 64, cudaHostAllocMapped) != cudaSuccess) st = GASR_ERR_CUDA;
            if (st == GASR_OK && cudaHostGetDevicePointer((void **)&a->host_words_dev, a->host_words, 0) != cudaSuccess) st = GASR_ERR_CUDA;
            if (st == GASR_OK) {
                a->host_words[0] = 0; a->host_words[1] = 0;
                for (cudaEvent_t *e : {&a->ev_r0, &a->ev_r1, &a->ev_g0, &a->ev_g1, &a->ev_d0, &a->ev_d1})
                    if (cudaEventCreate(e) != cudaSuccess) st = GASR_ERR_CUDA;
                if (cudaEventCreateWithFlags(&a->ev_go, cudaEventDisableTiming) != cudaSuccess) st = GASR_ERR_CUDA;
                if (cudaEventCreateWithFlags(&a->ev_cp, cudaEventDisableTiming) != cudaSuccess) st = GASR_ERR_CUDA;
            }
            a->stream_ok = st == GASR_OK;
        }
    }
    if (st != GASR_OK) { gasr_asr_destroy(a); return st; }
    *out = a;
    return GASR_OK;
}

int gasr_asr_destroy(gasr_asr *a) {
    if (a == nullptr) return GASR_OK;
    gasr_ctx *ctx = a->ctx;
    GASR_ENTER(ctx);
    cudaStreamSynchronize(ctx->stream);
    wave_destroy(a);
    for (auto v : {&a->w_ih, &a->w_hh, &a->b_ih, &a->b_hh, &a->hiddens})
        for (float *p : *v) if (p) gasr_free_device(ctx, p);
    for (float *p : {a->fc_w, a->fc_b, a->x_dev, a->logp, a->xproj_all, a->bias_all}) if (p) gasr_free_device(ctx, p);
    if (a->lens_dev) gasr_free_device(ctx, a->lens_dev);
    for (void *p : a->tc_abuf) if (p) gasr_free_device(ctx, p);
    for (void *p : a->tc_wbuf) if (p) gasr_free_device(ctx, p);
    for (void *p : a->h_planes) if (p) gasr_free_device(ctx, p);
    for (void *p : {a->x_planes, a->fc_wbuf, (void *)a->fc_b_pad, (void *)a->flags}) if (p) gasr_free_device(ctx, p);
    if (a->host_words) cudaFreeHost(a->host_words);
    for (cudaEvent_t e : {a->ev_cp, a->ev_go, a->ev_r0, a->ev_r1, a->ev_g0, a->ev_g1, a->ev_d0, a->ev_d1}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : a->sync_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : a->t0_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : a->t1_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : a->prof.pool) cudaEventDestroy(e);
    delete a;
    return GASR_OK;
}

int gasr_asr_set_weights(gasr_asr *a, const float *const *w_ih, const float *const *w_hh, const float *const *b_ih,
                         const float *const *b_hh, const float *fc_w, const float *fc_b) {
    GASR_CHECK(a != nullptr, "null gasr_asr");
    gasr_ctx *ctx = a->ctx;
    GASR_ENTER(ctx);
    GASR_CHECK(w_ih && w_hh && b_ih && b_hh && fc_w && fc_b, "gasr_asr_set_weights: null argument");
    const gasr_asr_config &c = a->cfg;
    const int D = a->D, G = a->G, H = c.H;
    for (int l = 0; l < c.L; l++) {
        const int in_l = l == 0 ? c.in : D * H;
        for (int d = 0; d < D; d++) {
            const int i = l * D + d;
            GASR_CHECK(w_ih[i] && w_hh[i] && b_ih[i] && b_hh[i], "gasr_asr_set_weights: null layer parameter %d", i);
            GASR_TRY(gasr_memcpy_h2d(ctx, a->w_ih[i], w_ih[i], sizeof(float) * in_l * G * H));
            GASR_TRY(gasr_memcpy_h2d(ctx, a->w_hh[i], w_hh[i], sizeof(float) * H * G * H));
            GASR_TRY(gasr_memcpy_h2d(ctx, a->b_ih[i], b_ih[i], sizeof(float) * G * H));
            GASR_TRY(gasr_memcpy_h2d(ctx, a->b_hh[i], b_hh[i], sizeof(float) * G * H));
        }
    }
    if (a->use_tc)
        for (int l = 0; l < c.L; l++)
            GASR_TRY(xproj_tc_prepare_weights(ctx, a->w_ih[l], l == 0 ? c.in : H, H, a->tc_wbuf[l], ctx->stream));
    GASR_TRY(gasr_memcpy_h2d(ctx, a->fc_w, fc_w, sizeof(float) * D * H * c.V));
    GASR_TRY(gasr_memcpy_h2d(ctx, a->fc_b, fc_b, sizeof(float) * c.V));
    if (a->wave) {
        GASR_TRY(wave_set_weights(a, fc_w, fc_b));
        a->have_weights = true;
        return GASR_OK;
    }
    if (a->stream_ok) {
        // output layer as a 32-column GEMM target: W_fc padded to [H, 32] -> W^T hi/lo planes; bias padded with zeros
        std::vector<float> wpad((size_t)H * 32, 0.0f), bpad(32, 0.0f);
        for (int k = 0; k < H; k++) for (int v = 0; v < c.V; v++) wpad[(size_t)k * 32 + v] = fc_w[(size_t)k * c.V + v];
        for (int v = 0; v < c.V; v++) bpad[v] = fc_b[v];
        GASR_TRY(ws_reserve(ctx, ctx->ws_misc, sizeof(float) * (size_t)H * 32 + 256));
        GASR_TRY(gasr_memcpy_h2d(ctx, ctx->ws_misc.ptr, wpad.data(), sizeof(float) * (size_t)H * 32));
        GASR_TRY(xproj_tc_prepare_weights(ctx, static_cast<const float *>(ctx->ws_misc.ptr), H, 32, a->fc_wbuf, ctx->stream));
        GASR_TRY(gasr_memcpy_h2d(ctx, a->fc_b_pad, bpad.data(), sizeof(float) * 32));
        for (int l = 0; l < c.L; l++)   // (b_hh + b_ih), RNN_Cell.cu:10
            GASR_TRY(launch_matadd(ctx, a->b_ih[l], H, a->b_hh[l], H, a->bias_all + (size_t)l * H, H, 1, H, 1.0f, ctx->stream));
        GASR_CUDA(cudaStreamSynchronize(ctx->stream));
        // TMA descriptors: target l < L = projection of layer l (A = planes of layer l-1 / of x), target L = output layer
        const int rows = c.T * c.N;
        for (int tg = 0; tg <= c.L; tg++) {
            const int K = tg == 0 ? c.in : H, Kp = ceil_div(K, TC_BK) * TC_BK;
            unsigned char *ab = static_cast<unsigned char *>(tg == 0 ? a->x_planes : a->h_planes[tg - 1]);
            unsigned char *wb = static_cast<unsigned char *>(tg < c.L ? a->tc_wbuf[tg] : a->fc_wbuf);
            const int nout = tg < c.L ? H : 32;
            GASR_TRY(tc_make_map(&a->xs_maps.m[4 * tg + 0], ab, rows, Kp, TC_BM));
            GASR_TRY(tc_make_map(&a->xs_maps.m[4 * tg + 1], ab + xproj_tc_a_bytes(rows, K) / 2, rows, Kp, TC_BM));
            GASR_TRY(tc_make_map(&a->xs_maps.m[4 * tg + 2], wb, nout, Kp, tg < c.L ? TC_BN : 32));
            GASR_TRY(tc_make_map(&a->xs_maps.m[4 * tg + 3], wb + xproj_tc_w_bytes(K, nout) / 2, nout, Kp, tg < c.L ? TC_BN : 32));
        }
    }
    a->have_weights = true;
    return GASR_OK;
}

static int asr_run_sequential(gasr_asr *a, const float *x_dev, char *out_paths, int *out_lens, float *out_scores) {
    gasr_ctx *ctx = a->ctx;
    const gasr_asr_config &c = a->cfg;
    cudaStream_t st = ctx->stream;
    const int rows = c.T * c.N;
    a->prof.used = 0;
    GASR_TRY(a->prof.mark(-1, st));
    GASR_TRY(rnn_forward_impl(ctx, c.cell, c.bidirectional, c.T, c.N, c.in, c.H, c.L, a->w_ih.data(), a->w_hh.data(),
                              a->b_ih.data(), a->b_hh.data(), x_dev, a->hiddens.data(), c.precision, st, &a->prof));
    GASR_TRY(launch_linear(ctx, a->hiddens[c.L - 1], a->D * c.H, a->fc_w, a->fc_b, a->logp, a->ldp, rows, a->D * c.H,
                           c.V, GASR_ACT_LOGSOFTMAX, st));
    GASR_TRY(a->prof.mark(2, st));
    CtcArgs ca = {a->logp, GASR_DOMAIN_LOG, c.T, c.N, c.V, a->ldp, c.beam, c.blank, a->vocab.data(), c.max_len,
                  c.nbest, out_paths, out_lens, out_scores, nullptr};
    a->decode_extras(ca);
    GASR_TRY(ctc_decode_launch(ctx, ca, st));
    GASR_TRY(a->prof.mark(3, st));
    GASR_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 4; i++) { a->stage_ms[i] = 0.0f; a->stage_launches[i] = 0; }
    for (size_t i = 1; i < a->prof.used; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, a->prof.pool[i - 1], a->prof.pool[i]);
        if (a->prof.tag[i] >= 0) { a->stage_ms[a->prof.tag[i]] += ms; a->stage_launches[a->prof.tag[i]] += 1; }
    }
    return ctc_decode_finish(ctx, ca);
}

// Pipelined path (unidirectional tanh stacks the cluster kernel supports, beam <= 32, vocab <= 32): the sequence is
// cut into chunks of `chunk` frames; layer l works on chunk c while layer l+1 works on chunk c-1 and the decoder on
// an even earlier one.  Every stage is sequential in time, so each owns a stream; cross-stage edges are events.
static int asr_run_pipelined(gasr_asr *a, const float *x_dev, char *out_paths, int *out_lens, float *out_scores) {
    gasr_ctx *ctx = a->ctx;
    const gasr_asr_config &c = a->cfg;
    const int T = c.T, N = c.N, H = c.H, L = c.L, Tc = a->chunk;
    const int C = ceil_div(T, Tc);
    cudaStream_t main_st = ctx->stream, dec_st = ctx->side[3], lin_st = ctx->side[4];
    auto layer_stream = [&](int l) { return ctx->side[l % 3]; };
    const size_t need_ev = (size_t)(L + 1) * C + 4;
    while (a->sync_ev.size() < need_ev) {
        cudaEvent_t e;
        GASR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        a->sync_ev.push_back(e);
    }
    a->n_timed = 0;
    auto timed_begin = [&](int tag, cudaStream_t st) -> int {
        if (a->n_timed == a->t0_ev.size()) {
            cudaEvent_t e0, e1;
            if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return GASR_ERR_CUDA;
            a->t0_ev.push_back(e0); a->t1_ev.push_back(e1); a->t_tag.push_back(tag);
        }
        a->t_tag[a->n_timed] = tag;
        return cudaEventRecord(a->t0_ev[a->n_timed], st) == cudaSuccess ? GASR_OK : GASR_ERR_CUDA;
    };
    auto timed_end = [&](cudaStream_t st) -> int {
        return cudaEventRecord(a->t1_ev[a->n_timed++], st) == cudaSuccess ? GASR_OK : GASR_ERR_CUDA;
    };
    cudaEvent_t ev_start = a->sync_ev[(size_t)(L + 1) * C], ev_dec_done = a->sync_ev[(size_t)(L + 1) * C + 1];
    GASR_CUDA(cudaEventRecord(ev_start, main_st));
    for (int l = 0; l < L && l < 3; l++) GASR_CUDA(cudaStreamWaitEvent(layer_stream(l), ev_start, 0));
    GASR_CUDA(cudaStreamWaitEvent(dec_st, ev_start, 0));
    GASR_CUDA(cudaStreamWaitEvent(lin_st, ev_start, 0));
    for (int l = 0; l < L; l++)   // (b_hh + b_ih), RNN_Cell.cu:10
        GASR_TRY(launch_matadd(ctx, a->b_ih[l], H, a->b_hh[l], H, a->bias_all + (size_t)l * H, H, 1, H, 1.0f, layer_stream(l)));
    CtcArgs ca = {a->logp, GASR_DOMAIN_LOG, T, N, c.V, a->ldp, c.beam, c.blank, a->vocab.data(), c.max_len,
                  c.nbest, out_paths, out_lens, out_scores, nullptr};
    a->decode_extras(ca);
    for (int ci = 0; ci < C; ci++) {
        const int f0 = ci * Tc, f1 = (ci + 1) * Tc < T ? (ci + 1) * Tc : T;
        const size_t row0 = (size_t)f0 * N;
        const int rows = (f1 - f0) * N;
        for (int l = 0; l < L; l++) {
            cudaStream_t st = layer_stream(l);
            const int in_l = l == 0 ? c.in : H;
            const float *src = l == 0 ? x_dev : a->hiddens[l - 1];
            float *xp = a->xproj_all + (size_t)l * T * N * H;
            if (l > 0) GASR_CUDA(cudaStreamWaitEvent(st, a->sync_ev[(size_t)(l - 1) * C + ci], 0));
            GASR_TRY(timed_begin(0, st));
            if (a->use_tc) {
                // each chunk owns a disjoint slice of the layer's bf16 scratch planes
                const size_t a_off = align_up(xproj_tc_a_bytes(Tc * N, in_l), 1024) * (size_t)ci;
                GASR_TRY(launch_xproj_tc(ctx, src + row0 * in_l, in_l, rows, in_l, H, a->tc_wbuf[l],
                                         static_cast<unsigned char *>(a->tc_abuf[l]) + a_off, a->bias_all + (size_t)l * H,
                                         xp + row0 * H, H, c.precision, st));
            } else {
                GASR_TRY(launch_matmul(ctx, src + row0 * in_l, in_l, 0, a->w_ih[l], H, 0, xp + row0 * H, H, rows, in_l, H,
                                       a->bias_all + (size_t)l * H, st));
            }
            GASR_TRY(timed_end(st));
            RnnLayerArgs ra;
            ra.cell = c.cell; ra.T = T; ra.N = N; ra.H = H; ra.reverse = 0;
            ra.xproj = xp; ra.ldxp = H; ra.w_hh = a->w_hh[l]; ra.b_hh = a->b_hh[l];
            ra.out = a->hiddens[l]; ra.ldo = H; ra.col0 = 0; ra.s0 = f0; ra.s1 = f1;
            GASR_TRY(timed_begin(1, st));
            GASR_TRY(launch_rnn_recurrence(ctx, ra, st));
            GASR_TRY(timed_end(st));
            GASR_CUDA(cudaEventRecord(a->sync_ev[(size_t)l * C + ci], st));
        }
        GASR_CUDA(cudaStreamWaitEvent(lin_st, a->sync_ev[(size_t)(L - 1) * C + ci], 0));
        GASR_TRY(timed_begin(2, lin_st));
        GASR_TRY(launch_linear(ctx, a->hiddens[L - 1] + row0 * H, H, a->fc_w, a->fc_b, a->logp + row0 * a->ldp, a->ldp,
                               rows, H, c.V, GASR_ACT_LOGSOFTMAX, lin_st));
        GASR_TRY(timed_end(lin_st));
        GASR_CUDA(cudaEventRecord(a->sync_ev[(size_t)L * C + ci], lin_st));
        GASR_CUDA(cudaStreamWaitEvent(dec_st, a->sync_ev[(size_t)L * C + ci], 0));
        ca.t0 = f0; ca.t1 = f1;
        GASR_TRY(timed_begin(3, dec_st));
        GASR_TRY(ctc_decode_launch(ctx, ca, dec_st));
        GASR_TRY(timed_end(dec_st));
    }
    GASR_CUDA(cudaEventRecord(ev_dec_done, dec_st));
    GASR_CUDA(cudaStreamWaitEvent(main_st, ev_dec_done, 0));
    GASR_CUDA(cudaStreamSynchronize(main_st));
    for (int i = 0; i < 4; i++) { a->stage_ms[i] = 0.0f; a->stage_launches[i] = 0; }
    for (size_t i = 0; i < a->n_timed; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, a->t0_ev[i], a->t1_ev[i]);
        a->stage_ms[a->t_tag[i]] += ms;
        a->stage_launches[a->t_tag[i]] += 1;
    }
    ca.t0 = 0; ca.t1 = 0;
    return ctc_decode_finish(ctx, ca);
}

// Streaming path: three persistent kernels -- the layer stack's recurrence (rnn_stream.cu), the projection / output
// layer GEMM (xproj_stream.cu) and the decoder (ctc_beam.cu, CTA kernel) -- run concurrently for the whole sequence and
// hand 128-row blocks (a few frames of the batch) to each other through counters in HBM.
static int asr_run_streaming(gasr_asr *a, const float *x_dev, const float *x_host, char *out_paths, int *out_lens,
                             float *out_scores) {
    gasr_ctx *ctx = a->ctx;
    const gasr_asr_config &c = a->cfg;
    const int T = c.T, N = c.N, H = c.H, L = c.L, nb = a->stream_blocks, rows = T * N;
    cudaStream_t main_st = ctx->stream, rec_st = ctx->side[0], gemm_st = ctx->side[1], dec_st = ctx->side[3];
    unsigned *h_done = a->flags, *xp_ready = a->flags + (size_t)L * nb, *lp_ready = a->flags + (size_t)2 * L * nb;
    unsigned *x_ready = lp_ready + nb, *misc = x_ready + nb;
    const int rec_nsub = 0;
    const int rec_groups = ceil_div(N, 16);
    const int rec_ctas_per_layer = rec_groups * (H / 64);
    a->epoch += 1;
    a->host_words[1] = 0;
    // Any early return below (a failed launch, a CUDA error) must not leave persistent kernels running behind the caller's
    // back: the guard raises the pipeline's abort word (every in-kernel wait polls it) and joins the pipeline's streams.
    struct Join {
        gasr_ctx *ctx; unsigned *abort_word; bool done = false;
        ~Join() {
            if (done) return;
            const unsigned one = 1u;
            cudaMemcpyAsync(abort_word, &one, sizeof(one), cudaMemcpyHostToDevice, ctx->side[4]);
            for (cudaStream_t s : {ctx->side[4], ctx->side[0], ctx->side[1], ctx->side[2], ctx->side[3], ctx->stream}) cudaStreamSynchronize(s);
            cudaGetLastError();
        }
    } join{ctx, misc + 1};
    // everything that may synchronise the device (allocations, function attributes) happens before the first persistent
    // kernel starts: once they run they wait for each other, not for the host
    CtcArgs ca = {a->logp, GASR_DOMAIN_LOG, T, N, c.V, a->ldp, c.beam, c.blank, a->vocab.data(), c.max_len,
                  c.nbest, out_paths, out_lens, out_scores, nullptr};
    a->decode_extras(ca);
    GASR_TRY(ctc_decode_reserve(ctx, ca));
    {
        XsParams prep = {};
        prep.n_targets = 1; prep.abort = misc + 1;
        XsTarget &t = prep.target[0];
        t.kind = XS_KIND_XPROJ; t.C = a->xproj_all; t.ldc = H; t.src_done = x_ready; t.dst_ready = xp_ready; t.kblocks = 1; t.bn = TC_BN; t.terms = 3; t.n_tiles = 1;
        GASR_TRY(launch_xproj_stream(ctx, a->xs_maps, prep, 0, gemm_st));
    }
    GASR_CUDA(cudaMemsetAsync(a->flags, 0, a->flags_bytes, main_st));
    GASR_TRY(ctc_decode_upload_vocab(ctx, ca, main_st));       // before the input copy occupies the copy engine
    ca.vocab_resident = true;
    if (x_host == nullptr) {
        GASR_CUDA(cudaMemsetAsync(x_ready, 0xff, sizeof(unsigned) * nb, main_st));
        GASR_TRY(xproj_tc_split_rows(ctx, x_dev, c.in, rows, c.in, a->x_planes, main_st));
        GASR_CUDA(cudaEventRecord(a->ev_go, main_st));
    } else {
        // host input: the batch is copied in slices on the copy stream; each slice is split into its bf16 planes and
        // then marked ready, so the pipeline starts after the first slice and the rest of the copy hides behind it
        GASR_CUDA(cudaEventRecord(a->ev_go, main_st));
        GASR_CUDA(cudaStreamWaitEvent(ctx->side[2], a->ev_go, 0));
    }
    // Only the first slices are queued before the kernels are launched (enough to keep the copy engine busy meanwhile);
    // queueing all 16 first cost ~50 API calls of host time before the recurrence kernel could start.
    const int slices = x_host ? (nb >= 16 ? 16 : 1) : 0;
    int slices_first = slices < 4 ? slices : 4;
    if (x_host) {
        // Pageable host memory: the driver stages such copies, and staging them while the persistent kernels hold the
        // GPU never completed (the watchdogs fired).  Only pinned / registered buffers are copied behind the launches.
        cudaPointerAttributes at = {};
        if (cudaPointerGetAttributes(&at, x_host) != cudaSuccess || at.type != cudaMemoryTypeHost) { cudaGetLastError(); slices_first = slices; }
    }
    auto issue_slices = [&](int s_begin, int s_end) -> int {
        cudaStream_t cp_st = ctx->side[2];
        for (int sidx = s_begin; sidx < s_end; sidx++) {
            const int b0 = (int)((long long)nb * sidx / slices), b1 = (int)((long long)nb * (sidx + 1) / slices);
            if (b1 == b0) continue;
            const size_t r0 = (size_t)b0 * 128, nr = (size_t)(b1 - b0) * 128;
            GASR_CUDA(cudaMemcpyAsync(a->x_dev + r0 * c.in, x_host + r0 * c.in, sizeof(float) * nr * c.in, cudaMemcpyHostToDevice, cp_st));
            GASR_TRY(xproj_tc_split_rows_range(ctx, a->x_dev, c.in, rows, (int)r0, (int)nr, c.in, a->x_planes, cp_st));
            GASR_CUDA(cudaMemsetAsync(x_ready + b0, 0xff, sizeof(unsigned) * (size_t)(b1 - b0), cp_st));
        }
        if (s_end == slices && slices > 0) GASR_CUDA(cudaEventRecord(a->ev_cp, cp_st));
        return GASR_OK;
    };
    GASR_TRY(issue_slices(0, slices_first));
#ifdef GASR_STREAM_HOOKS
    const bool dbg = getenv("GASR_STREAM_DEBUG") != nullptr;   // instrumented build only (make TRACE=1): stage-by-stage diagnosis
#else
    constexpr bool dbg = false;
#endif
    if (dbg) { GASR_CUDA(cudaStreamSynchronize(main_st)); fprintf(stderr, "[stream] prep ok\n"); }

    // ---- recurrence: all layers, one launch ------------------------------------------------------------------------
    RnnStreamParams rp = {};
    rp.T = T; rp.N = N; rp.L = L; rp.groups = rec_groups; rp.nsub = rec_nsub; rp.frames_per_block = a->stream_fpb;
    rp.xp_need = XS_EPI_WARPS * (H / TC_BN);
    rp.error = a->host_words_dev + 1; rp.abort = misc + 1; rp.started = misc; rp.host_go = a->host_words_dev; rp.epoch = a->epoch;
    for (int l = 0; l < L; l++) {
        RnnStreamLayer &y = rp.layer[l];
        y.xproj = a->xproj_all + (size_t)l * rows * H; y.ldxp = H; y.w_hh = a->w_hh[l];
        y.out = nullptr; y.ldo = H;
        y.out_hi = static_cast<__nv_bfloat16 *>(a->h_planes[l]);
        y.out_lo = reinterpret_cast<__nv_bfloat16 *>(static_cast<unsigned char *>(a->h_planes[l]) + xproj_tc_a_bytes(rows, H) / 2);
        y.ldp = H;
        y.xp_ready = xp_ready + (size_t)l * nb;
        y.h_done = h_done + (size_t)l * nb;
    }
    GASR_CUDA(cudaStreamWaitEvent(rec_st, a->ev_go, 0));
    GASR_CUDA(cudaEventRecord(a->ev_r0, rec_st));
    GASR_TRY(launch_rnn_stream(ctx, rp, H, rec_st));
    GASR_CUDA(cudaEventRecord(a->ev_r1, rec_st));
    {   // the recurrence needs whole SMs in cluster-sized groups: everything else is launched once it is resident
        const auto t0 = std::chrono::steady_clock::now();
        volatile int *go = a->host_words;
        while (*go != a->epoch) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 10.0) {
                cudaStreamSynchronize(rec_st);
                set_error("streaming pipeline: the recurrence kernel did not become resident");
                return GASR_ERR_CUDA;
            }
        }
    }
    if (dbg) {
        cudaError_t q = cudaStreamQuery(rec_st);
        fprintf(stderr, "[stream] recurrence resident, query: %s\n", cudaGetErrorString(q));
        if (q != cudaSuccess && q != cudaErrorNotReady) { set_error("recurrence kernel failed: %s", cudaGetErrorString(q)); return GASR_ERR_CUDA; }
#ifdef GASR_STREAM_HOOKS
        if (getenv("GASR_STREAM_DEBUG")[0] == '2') {
            q = cudaStreamSynchronize(rec_st);
            fprintf(stderr, "[stream] recurrence alone: %s, error word %d\n", cudaGetErrorString(q), a->host_words[1]);
            return GASR_ERR_CUDA;
        }
#endif
    }
    // ---- projection + output-layer GEMM -------------------------------------------------------------------------------
    XsParams xp = {};
    xp.M = rows; xp.n_blocks = nb; xp.n_targets = L + 1; xp.error = a->host_words_dev + 1; xp.abort = misc + 1;
    // CTAs per target in proportion to its operand traffic per block (the tile engine is L2->smem bound)
    int gemm_ctas = 0;
    {
        double w[XS_MAX_TARGETS], wsum = 0.0;
        for (int tg = 0; tg <= L; tg++) {
            const int K = tg == 0 ? c.in : H, bn = tg < L ? TC_BN : 32, nt = tg < L ? H / TC_BN : 1;
            w[tg] = (double)nt * ceil_div(K, TC_BK) * (128 + bn);
            wsum += w[tg];
        }
        for (int tg = 0; tg <= L; tg++) {
            int n = (int)(a->stream_gemm_ctas * w[tg] / wsum + 0.5);
            if (n < 1) n = 1;
            if (tg == L && n < 2) n = 2;
            xp.target[tg].cta0 = gemm_ctas; xp.target[tg].nctas = n;
            gemm_ctas += n;
        }
    }
    for (int tg = 0; tg <= L; tg++) {
        XsTarget &t = xp.target[tg];
        const int K = tg == 0 ? c.in : H;
        t.kblocks = ceil_div(K, TC_BK); t.terms = c.precision == GASR_PREC_BF16 ? 1 : 3;
        t.src_done = tg == 0 ? x_ready : h_done + (size_t)(tg - 1) * nb;
        t.src_need = tg == 0 ? 1 : rec_ctas_per_layer;
        if (tg < L) {
            t.kind = XS_KIND_XPROJ; t.n_tiles = H / TC_BN; t.bn = TC_BN; t.V = 0;
            t.C = a->xproj_all + (size_t)tg * rows * H; t.ldc = H; t.bias = a->bias_all + (size_t)tg * H;
            t.dst_ready = xp_ready + (size_t)tg * nb;
        } else {
            t.kind = XS_KIND_LOGSOFTMAX; t.n_tiles = 1; t.bn = 32; t.V = c.V; t.terms = 3;
            t.C = a->logp; t.ldc = a->ldp; t.bias = a->fc_b_pad; t.dst_ready = lp_ready;
        }
    }
#ifdef GASR_STREAM_HOOKS   // ablation / alone-timing paths of the instrumented build (tools/capture_profiles.sh)
    if (dbg && getenv("GASR_DEBUG_TERMS")) for (int tg = 0; tg <= L; tg++) xp.target[tg].terms = atoi(getenv("GASR_DEBUG_TERMS"));
    if (dbg && getenv("GASR_DEBUG_SAMEMAP")) for (int tg = 0; tg <= L; tg++) xp.target[tg].kind |= 32;
    if (dbg && getenv("GASR_DEBUG_PRINT")) for (int tg = 0; tg <= L; tg++) xp.target[tg].kind |= 64;
    if (dbg && getenv("GASR_DEBUG_NOEPI")) for (int tg = 0; tg <= L; tg++) xp.target[tg].kind |= 16;
    if (dbg && getenv("GASR_STREAM_DEBUG")[0] == '3') {
        // tile-engine throughput: every dependency preset, the GEMM kernel alone
        cudaStreamSynchronize(rec_st);
        GASR_CUDA(cudaMemsetAsync(a->flags, 0xff, sizeof(unsigned) * (size_t)L * nb, gemm_st));
        GASR_CUDA(cudaMemsetAsync(misc, 0, 64, gemm_st));
        a->host_words[1] = 0;
        GASR_CUDA(cudaEventRecord(a->ev_g0, gemm_st));
        GASR_TRY(launch_xproj_stream(ctx, a->xs_maps, xp, gemm_ctas, gemm_st));
        GASR_CUDA(cudaEventRecord(a->ev_g1, gemm_st));
        GASR_CUDA(cudaStreamSynchronize(gemm_st));
        float ms = 0; cudaEventElapsedTime(&ms, a->ev_g0, a->ev_g1);
        fprintf(stderr, "[stream] gemm alone: %.3f ms, %d CTAs:", ms, gemm_ctas);
        for (int tg = 0; tg <= L; tg++) fprintf(stderr, " target %d -> %d CTAs", tg, xp.target[tg].nctas);
        fprintf(stderr, " (error word %d)\n", a->host_words[1]);
        if (getenv("GASR_DEBUG_REC_ALONE")) {
            // the recurrence alone: every projection block is already counted complete
            a->epoch += 1; rp.epoch = a->epoch;
            const int variant = atoi(getenv("GASR_DEBUG_REC_ALONE"));
            for (int l = 0; l < L; l++) {
                if (variant & 2) rp.layer[l].h_done = nullptr;
                if (variant & 4) rp.layer[l].xp_ready = nullptr;
                if (variant & 8) { rp.layer[l].out_hi = nullptr; rp.layer[l].out_lo = nullptr; rp.layer[l].out = a->hiddens[l]; }
            }
            GASR_CUDA(cudaMemsetAsync(misc, 0, 64, rec_st));
            GASR_CUDA(cudaMemsetAsync(h_done, 0, sizeof(unsigned) * (size_t)L * nb, rec_st));
            GASR_CUDA(cudaEventRecord(a->ev_r0, rec_st));
            GASR_TRY(launch_rnn_stream(ctx, rp, H, rec_st));
            GASR_CUDA(cudaEventRecord(a->ev_r1, rec_st));
            GASR_CUDA(cudaStreamSynchronize(rec_st));
            cudaEventElapsedTime(&ms, a->ev_r0, a->ev_r1);
            fprintf(stderr, "[stream] recurrence alone (all layers, inputs ready): %.3f ms (error word %d)\n", ms, a->host_words[1]);
        }
        return GASR_ERR_CUDA;
    }
#endif
    GASR_CUDA(cudaStreamWaitEvent(gemm_st, a->ev_go, 0));
    GASR_CUDA(cudaEventRecord(a->ev_g0, gemm_st));
#ifdef GASR_STREAM_HOOKS
    if (!getenv("GASR_STREAM_INJECT_LOST_PRODUCER"))     // (test hook of the instrumented build: the GEMM kernel never starts -> watchdogs -> fallback)
#endif
        GASR_TRY(launch_xproj_stream(ctx, a->xs_maps, xp, gemm_ctas, gemm_st));
    GASR_CUDA(cudaEventRecord(a->ev_g1, gemm_st));
    if (dbg) {
        cudaError_t q = cudaStreamSynchronize(gemm_st);
        fprintf(stderr, "[stream] gemm done: %s, error word %d\n", cudaGetErrorString(q), a->host_words[1]);
        q = cudaStreamSynchronize(rec_st);
        fprintf(stderr, "[stream] recurrence done: %s, error word %d\n", cudaGetErrorString(q), a->host_words[1]);
    }
    // ---- decoder ----------------------------------------------------------------------------------------------------------
    ca.lp_ready = lp_ready; ca.lp_need = XS_EPI_WARPS; ca.lp_fpb = a->stream_fpb; ca.error = a->host_words_dev + 1; ca.abort = misc + 1;
    GASR_CUDA(cudaStreamWaitEvent(dec_st, a->ev_go, 0));
    GASR_CUDA(cudaEventRecord(a->ev_d0, dec_st));
    GASR_TRY(ctc_decode_launch(ctx, ca, dec_st));
    GASR_CUDA(cudaEventRecord(a->ev_d1, dec_st));
    GASR_TRY(issue_slices(slices_first, slices));              // the rest of the input copy, behind the running pipeline
    GASR_CUDA(cudaStreamWaitEvent(main_st, a->ev_r1, 0));
    GASR_CUDA(cudaStreamWaitEvent(main_st, a->ev_g1, 0));
    GASR_CUDA(cudaStreamWaitEvent(main_st, a->ev_d1, 0));
    if (x_host != nullptr) GASR_CUDA(cudaStreamWaitEvent(main_st, a->ev_cp, 0));
    GASR_CUDA(cudaStreamSynchronize(main_st));
    join.done = true;                                           // everything has completed (main_st was synchronised)
    if (a->host_words[1] != 0) {
        set_error("streaming pipeline: watchdog fired (code %d): a persistent kernel waited too long for its producer", a->host_words[1]);
        return GASR_ERR_CUDA;
    }
    for (int i = 0; i < 4; i++) { a->stage_ms[i] = 0.0f; a->stage_launches[i] = 0; }
    cudaEventElapsedTime(&a->stage_ms[0], a->ev_g0, a->ev_g1); a->stage_launches[0] = 1;
    cudaEventElapsedTime(&a->stage_ms[1], a->ev_r0, a->ev_r1); a->stage_launches[1] = 1;
    cudaEventElapsedTime(&a->stage_ms[3], a->ev_d0, a->ev_d1); a->stage_launches[3] = 1;
    ca.lp_ready = nullptr;
    join.done = true;
    return ctc_decode_finish(ctx, ca);
}

// The streaming mode needs every CTA of its three kernels resident at once (a whole, otherwise idle B200).  If its
// residency handshake or a watchdog failed without a CUDA fault (e.g. the GPU is shared), the pipeline object drops to
// the time-chunked mode -- still the same kernels' siblings on the GPU, never a CPU path -- and says so once.
static bool asr_stream_recover(gasr_asr *a, int rc) {
    if (rc != GASR_ERR_CUDA || cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return false;
    a->stream_ok = false;            // visible to the caller: gasr_asr_stage_launches reports the chunked mode from now on
    return true;
}

int gasr_asr_run_device(gasr_asr *a, const float *x_dev, char *out_paths, int *out_lens, float *out_scores) {
    GASR_CHECK(a != nullptr, "null gasr_asr");
    gasr_ctx *ctx = a->ctx;
    GASR_ENTER(ctx);
    GASR_CHECK(a->have_weights, "gasr_asr_run: weights not set");
    GASR_CHECK(x_dev && out_paths && out_lens && out_scores, "gasr_asr_run: null argument");
    if (a->wave) {
        GASR_TRY(wave_submit(a, x_dev, nullptr));
        return wave_collect(a, out_paths, out_lens, out_scores);
    }
    if (a->stream_ok) {
        const int rc = asr_run_streaming(a, x_dev, nullptr, out_paths, out_lens, out_scores);
        if (rc == GASR_OK) return rc;
#ifdef GASR_STREAM_HOOKS
        if (getenv("GASR_STREAM_DEBUG")) return rc;
#endif
        if (!asr_stream_recover(a, rc)) return rc;
    }
    if (a->chunk > 0) return asr_run_pipelined(a, x_dev, out_paths, out_lens, out_scores);
    return asr_run_sequential(a, x_dev, out_paths, out_lens, out_scores);
}

int gasr_asr_run_host(gasr_asr *a, const float *x_host, char *out_paths, int *out_lens, float *out_scores) {
    GASR_CHECK(a != nullptr, "null gasr_asr");
    gasr_ctx *ctx = a->ctx;
    GASR_ENTER(ctx);
    GASR_CHECK(x_host != nullptr, "gasr_asr_run_host: null input");
    const gasr_asr_config &c = a->cfg;
    GASR_CHECK(a->have_weights, "gasr_asr_run: weights not set");
    GASR_CHECK(out_paths && out_lens && out_scores, "gasr_asr_run: null argument");
    if (a->wave) {
        GASR_TRY(wave_submit(a, nullptr, x_host));
        return wave_collect(a, out_paths, out_lens, out_scores);
    }
    if (a->stream_ok) {
        const int rc = asr_run_streaming(a, a->x_dev, x_host, out_paths, out_lens, out_scores);
        if (rc == GASR_OK) return rc;
#ifdef GASR_STREAM_HOOKS
        if (getenv("GASR_STREAM_DEBUG")) return rc;
#endif
        if (!asr_stream_recover(a, rc)) return rc;
    }
    GASR_CUDA(cudaMemcpyAsync(a->x_dev, x_host, sizeof(float) * (size_t)c.T * c.N * c.in, cudaMemcpyHostToDevice,
                              ctx->stream));
    return gasr_asr_run_device(a, a->x_dev, out_paths, out_lens, out_scores);
}

int gasr_asr_logprobs(gasr_asr *a, const float **logp_dev, int *ldp) {
    GASR_CHECK(a && logp_dev && ldp, "gasr_asr_logprobs: null argument");
    if (a->wave) { DeviceGuard guard(a->ctx->device); return wave_logprobs(a, logp_dev, ldp); }
    *logp_dev = a->logp;
    *ldp = a->ldp;
    return GASR_OK;
}

int gasr_asr_stage_times(gasr_asr *a, float *ms4) {
    GASR_CHECK(a && ms4, "gasr_asr_stage_times: null argument");
    for (int i = 0; i < 4; i++) ms4[i] = a->stage_ms[i];
    return GASR_OK;
}

int gasr_asr_stage_launches(gasr_asr *a, int *n4, int *chunk_frames) {
    GASR_CHECK(a && n4, "gasr_asr_stage_launches: null argument");
    for (int i = 0; i < 4; i++) n4[i] = a->stage_launches[i];
    if (chunk_frames) *chunk_frames = a->wave ? -2 : (a->stream_ok ? -1 : a->chunk);
    return GASR_OK;
}

/* ---- asynchronous form + profiling (wave engine) -------------------------------------------------------------------- */
int gasr_asr_submit_host(gasr_asr *a, const float *x_host) {
    GASR_CHECK(a != nullptr && x_host != nullptr, "gasr_asr_submit_host: null argument");
    GASR_ENTER(a->ctx);
    GASR_CHECK(a->have_weights, "gasr_asr_submit: weights not set");
    if (!a->wave) { set_error("gasr_asr_submit: this configuration runs outside the wave engine; use gasr_asr_run_host"); return GASR_ERR_UNSUPPORTED; }
    return wave_submit(a, nullptr, x_host);
}

int gasr_asr_submit_device(gasr_asr *a, const float *x_dev) {
    GASR_CHECK(a != nullptr && x_dev != nullptr, "gasr_asr_submit_device: null argument");
    GASR_ENTER(a->ctx);
    GASR_CHECK(a->have_weights, "gasr_asr_submit: weights not set");
    if (!a->wave) { set_error("gasr_asr_submit: this configuration runs outside the wave engine; use gasr_asr_run_device"); return GASR_ERR_UNSUPPORTED; }
    return wave_submit(a, x_dev, nullptr);
}

int gasr_asr_collect(gasr_asr *a, char *out_paths, int *out_lens, float *out_scores) {
    GASR_CHECK(a != nullptr && out_paths && out_lens && out_scores, "gasr_asr_collect: null argument");
    GASR_ENTER(a->ctx);
    if (!a->wave) { set_error("gasr_asr_collect: nothing was submitted"); return GASR_ERR_INVALID; }
    return wave_collect(a, out_paths, out_lens, out_scores);
}

int gasr_asr_set_lengths(gasr_asr *a, const int *lens_host) {
    GASR_CHECK(a != nullptr, "null gasr_asr");
    gasr_ctx *ctx = a->ctx;
    GASR_ENTER(ctx);
    const gasr_asr_config &c = a->cfg;
    GASR_CHECK(!wave_pending(a), "gasr_asr_set_lengths: a batch is in flight (collect it first)");
    if (lens_host == nullptr) {
        if (a->lens_dev) { GASR_CUDA(cudaStreamSynchronize(ctx->stream)); gasr_free_device(ctx, a->lens_dev); a->lens_dev = nullptr; }
        return GASR_OK;
    }
    GASR_CHECK(!c.bidirectional, "gasr_asr_set_lengths: a bidirectional stack would read the padding frames (pack the batch by length instead)");
    for (int n = 0; n < c.N; n++)
        GASR_CHECK(lens_host[n] >= 1 && lens_host[n] <= c.T, "gasr_asr_set_lengths: length %d of utterance %d outside 1..T=%d", lens_host[n], n, c.T);
    if (!a->lens_dev) GASR_TRY(gasr_malloc_device(ctx, sizeof(int) * (size_t)c.N, (void **)&a->lens_dev));
    GASR_CUDA(cudaMemcpyAsync(a->lens_dev, lens_host, sizeof(int) * (size_t)c.N, cudaMemcpyHostToDevice, ctx->stream));
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    return GASR_OK;
}

int gasr_asr_enable_timesteps(gasr_asr *a, int on) {
    GASR_CHECK(a != nullptr, "null gasr_asr");
    gasr_ctx *ctx = a->ctx;
    GASR_ENTER(ctx);
    const gasr_asr_config &c = a->cfg;
    a->want_ts = on != 0;
    if (a->want_ts) a->ts_host.assign((size_t)c.N * c.nbest * c.max_len, 0);
    if (a->wave) GASR_TRY(wave_refresh_decoder(a));
    return GASR_OK;
}

int gasr_asr_timesteps(gasr_asr *a, int *out_timesteps) {
    GASR_CHECK(a != nullptr && out_timesteps != nullptr, "gasr_asr_timesteps: null argument");
    GASR_CHECK(a->want_ts, "gasr_asr_timesteps: call gasr_asr_enable_timesteps(asr, 1) before the run");
    memcpy(out_timesteps, a->ts_host.data(), sizeof(int) * a->ts_host.size());
    return GASR_OK;
}

int gasr_asr_profile(gasr_asr *a, int on) {
    GASR_CHECK(a != nullptr, "null gasr_asr");
    a->profile = on != 0;
    return GASR_OK;
}

int gasr_asr_last_ms(gasr_asr *a, float *ms) {
    GASR_CHECK(a != nullptr && ms != nullptr, "gasr_asr_last_ms: null argument");
    *ms = wave_last_ms(a);
    return GASR_OK;
}

}  // extern "C"
