// asr_wave.cu -- the throughput engine behind gasr_asr_run_* / gasr_job_*: RNN stack -> Linear + log-softmax -> CTC beam search
// for MANY utterances per GPU (BASELINE.json cfg5: 8192 utterances).
//
// The reference runs one (t, layer) cell call at a time with a host sync after every BLAS call (RNN.cu:15-27,
// cuMatrix.cpp:61) and one decode step per frame from the host (CTCBeamSearch.cu:278-281).  Every utterance is independent
// of every other (RNN.cu:15-27 batches rows; CTCBeamSearch.cu:416 segments by utterance), so throughput comes from width:
//
//   * the batch is padded to whole GROUPS OF 128 utterances (the M of one tcgen05.mma); all activations are time-major
//     with Npad rows per frame: x planes [T*Npad, Kp] (bf16 hi/lo), per layer xproj [T*Npad, H] fp32 and the hidden sequence
//     as bf16 hi/lo planes [T*Npad, H] (written by the recurrence, read back by it through TMA and by the next GEMM),
//     log-probabilities [T*Npad, 32];
//   * time is cut into chunks of Tc frames; chunk c flows  split -> GEMM_0 -> REC_0 -> GEMM_1 -> REC_1 ... -> FC+log-softmax
//     -> decode,  every stage on its own stream, every edge a CUDA event (NO kernel ever waits for another kernel: the
//     persistent-kernel streaming mode of round 1 stays available as an opt-in latency mode only);
//   * GEMMs: the persistent tcgen05 tile engine (xproj_stream.cu) with its dependencies preset, reading the bf16 planes
//     directly; recurrence: rnn_wide.cu (W_hh resident in shared memory, clusters of H/64 CTAs per 128/256 utterances);
//     decoder: one warp / CTA per utterance, beam parked in HBM between chunks (ctc_beam.cu).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "asr.cuh"
#include "common.cuh"
#include "gemm_pair.cuh"
#include "rnn_wide.cuh"
#include "stream.cuh"
#include "tc_common.cuh"

namespace gasr {

struct WaveState {
    int Npad = 0, Tc = 0, C = 0, Kp0 = 0, xp_slots = 1, rec_groups = 2;
    size_t rows_p = 0;                                   // T * Npad
    void *x_planes = nullptr;                            // [rows_p, Kp0] hi, then lo
    std::vector<void *> h_planes, wih, whh;              // per layer: hidden planes (hi, lo), W_ih^T planes, W_hh^T planes
    std::vector<float *> xp;                             // per layer: ring of xp_slots chunks, [xp_slots * Tc * Npad, H] fp32
    void *fc_wbuf = nullptr;
    float *fc_b_pad = nullptr, *bias_all = nullptr, *logp = nullptr, *logp_dense = nullptr;
    unsigned *flags = nullptr;                           // [ones: max blocks][sink: max blocks][misc 16]
    int max_blocks = 0;
    std::vector<RnnWidePlan> rec;
    std::vector<XsMaps> maps;                            // per layer + output layer
    std::vector<XsMaps> pmaps;                           // per layer: the same operands with 128-row W^T boxes (CTA-pair GEMM)
    cudaStream_t st_in = nullptr, st_fc = nullptr, st_dec = nullptr;
    std::vector<cudaStream_t> st_g, st_r;
    std::vector<cudaEvent_t> ev_in, ev_fc;
    std::vector<std::vector<cudaEvent_t>> ev_g, ev_r;
    cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    std::vector<cudaEvent_t> t0, t1;                     // per-launch timing pairs (profile mode)
    std::vector<int> tag;
    size_t n_timed = 0;
    CtcArgs ca = {};
    bool pending = false, serial = false, pair = false;
    float last_ms = 0.0f;
};

// fp32 rows [frames * N, K] (row f * N + n) -> bf16 hi / lo planes rows f * Npad + n of [*, Kp] (columns >= K zero; the pad
// rows n >= N stay zero from the allocation)
__global__ void wave_split_kernel(const float *__restrict__ x, int ldx, int N, int Npad, int frames, int K, int Kp,
                                  __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
    const int half = Kp / 2;
    const size_t total = (size_t)frames * N * half;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / half;
        const int c = (int)(i - r * half) * 2;
        const int f = (int)(r / N), n = (int)(r - (size_t)f * N);
        const float a = c < K ? x[r * ldx + c] : 0.0f, b = c + 1 < K ? x[r * ldx + c + 1] : 0.0f;
        const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
        const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
        const size_t o = ((size_t)f * Npad + n) * half + c / 2;
        reinterpret_cast<__nv_bfloat162 *>(hi)[o] = __halves2bfloat162(ah, bh);
        reinterpret_cast<__nv_bfloat162 *>(lo)[o] = __halves2bfloat162(al, bl);
    }
}

bool wave_supported(const gasr_ctx *ctx, const gasr_asr_config &c) {
    return c.cell == GASR_CELL_TANH && !c.bidirectional && (c.H == 128 || c.H == 256 || c.H == 512) && c.beam <= 32 && c.V <= 32 &&
           ctx->cluster_ok && rnn_wide_supported(ctx, c.H) &&
           (long long)c.T * (ceil_div(c.N, 256) * 256) < (1ll << 31) / 64;
}

static size_t planes_bytes(size_t rows, int Kp) { return 2 * align_up(rows * (size_t)Kp * 2, 1024); }

int wave_create(gasr_asr *a) {
    gasr_ctx *ctx = a->ctx;
    const gasr_asr_config &c = a->cfg;
    WaveState *w = new WaveState();
    a->wave = w;
    const int L = c.L, H = c.H;
    w->pair = ctx->opt.rnn_pair && rnn_wide2_supported(ctx, c.H) && c.N > 128;   // CTA-pair recurrence: groups of 256 utterances
    w->Npad = w->pair ? ceil_div(c.N, 256) * 256 : ceil_div(c.N, 128) * 128;
    {
        // Groups per cluster: two (ping-pong: TMA / MMA of one group overlap epilogue / exchange of the other).  One group per
        // cluster gives small batches more clusters but measured slower even at 1024 utterances (22.4 vs 18.9 ms).
        w->rec_groups = ctx->opt.rnn_groups > 0 ? ctx->opt.rnn_groups : 2;
    }
    // 50-frame time chunks; 25 for batches of 257..1024 utterances, where the pipeline is bound by its dependency chains and the
    // shorter fill outweighs the extra launches (1024 utterances: 19.5 -> 18.2 ms together with equal stream priorities)
    w->Tc = ctx->opt.chunk > 0 ? ctx->opt.chunk : ((c.N > 256 && c.N <= 1024) ? 25 : 50);
    if (w->Tc > c.T) w->Tc = c.T;
    w->C = ceil_div(c.T, w->Tc);
    w->Kp0 = ceil_div(c.in, TC_BK) * TC_BK;
    w->rows_p = (size_t)c.T * w->Npad;
    w->max_blocks = ceil_div(w->Tc * w->Npad, TC_BM);
    // xproj only lives from a chunk's GEMM to the same chunk's recurrence: a ring of chunk slots instead of the whole sequence
    w->xp_slots = w->C < 4 ? w->C : 4;
    int st = GASR_OK;
    auto allocv = [&](void **p, size_t bytes) { if (st == GASR_OK) st = gasr_malloc_device(ctx, bytes, p); };
    allocv(&w->x_planes, planes_bytes(w->rows_p, w->Kp0));
    w->h_planes.assign(L, nullptr); w->wih.assign(L, nullptr); w->whh.assign(L, nullptr); w->xp.assign(L, nullptr);
    for (int l = 0; l < L; l++) {
        allocv(&w->h_planes[l], 2 * rnn_wide_plane_bytes(c.T, w->Npad, H));
        allocv(&w->wih[l], xproj_tc_w_bytes(l == 0 ? c.in : H, H) + 1024);
        allocv(&w->whh[l], xproj_tc_w_bytes(H, H) + 1024);
        allocv((void **)&w->xp[l], sizeof(float) * (size_t)w->xp_slots * w->Tc * w->Npad * H);
    }
    allocv(&w->fc_wbuf, xproj_tc_w_bytes(H, 32) + 1024);
    allocv((void **)&w->fc_b_pad, sizeof(float) * 32);
    allocv((void **)&w->bias_all, sizeof(float) * (size_t)L * H);
    allocv((void **)&w->logp, sizeof(float) * w->rows_p * 32);
    allocv((void **)&w->flags, sizeof(unsigned) * (2 * (size_t)w->max_blocks + 16));
    if (st != GASR_OK) return st;
    GASR_CUDA(cudaMemsetAsync(w->flags, 0xff, sizeof(unsigned) * (size_t)w->max_blocks, ctx->stream));   // every source block complete
    // streams: recurrences first in line for SMs (the longest dependency chain), then GEMMs, then the decoder
    int prio_lo = 0, prio_hi = 0;
    GASR_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const int prio_mid = prio_hi < prio_lo ? prio_hi + 1 : prio_lo;
    auto mk = [&](cudaStream_t *s, int prio) { if (st == GASR_OK && cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, prio) != cudaSuccess) st = GASR_ERR_CUDA; };
    w->st_g.assign(L, nullptr); w->st_r.assign(L, nullptr);
    if (ctx->opt.wave_serial) {
        // diagnostic: every stage on ONE stream, nothing overlaps -- the profiled stage sums are then the stages' costs alone
        mk(&w->st_in, prio_mid);
        w->st_fc = w->st_dec = w->st_in;
        for (int l = 0; l < L; l++) w->st_g[l] = w->st_r[l] = w->st_in;
        w->serial = true;
    } else {
        // Stream priorities.  Favouring the recurrences (the longest dependency chain) is kept for small batches, where the chain
        // sets the time (1024 utterances with 50-frame chunks: 18.5 ms vs 18.9 with equal priorities; with 25-frame chunks equal
        // priorities win, 18.2 vs 18.6); with more utterances the recurrence launches
        // (64 SMs each, mostly waiting on TMA latency) then take SMs from the stages that use them better: equal priorities give
        // 2048 utterances 32.4 -> 27.5 ms, 4096 58.6 -> 55.7 ms, the 8192-utterance step 116.5 -> 110.7 ms (same-box A/B).
        // GASR_WAVE_PRIO: 0 = recurrences > GEMMs > decoder, 1 = all equal, 2 = decoder > GEMMs > recurrences (24.6 ms at 1024),
        // 3 = GEMMs > recurrences > decoder (117 ms), 4 = decoder > recurrences > GEMMs, 5 = decoder = GEMMs > recurrences
        const int pm = ctx->opt.wave_prio >= 0 ? ctx->opt.wave_prio : (c.N <= 256 ? 0 : 1);
        const int p_rec = pm == 1 ? prio_lo : pm == 2 ? prio_lo : pm == 3 ? prio_mid : pm == 4 ? prio_mid : pm == 5 ? prio_lo : prio_hi;
        const int p_gemm = pm == 1 ? prio_lo : pm == 3 ? prio_hi : pm == 4 ? prio_lo : pm == 5 ? prio_hi : prio_mid;
        const int p_dec = (pm == 2 || pm == 4 || pm == 5) ? prio_hi : prio_lo;
        mk(&w->st_in, p_gemm); mk(&w->st_fc, p_gemm); mk(&w->st_dec, p_dec);
        for (int l = 0; l < L; l++) { mk(&w->st_g[l], p_gemm); mk(&w->st_r[l], p_rec); }
    }
    auto mkev = [&](cudaEvent_t *e, bool timing) {
        if (st == GASR_OK && cudaEventCreateWithFlags(e, timing ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) st = GASR_ERR_CUDA;
    };
    w->ev_in.assign(w->C, nullptr); w->ev_fc.assign(w->C, nullptr);
    w->ev_g.assign(L, std::vector<cudaEvent_t>(w->C, nullptr)); w->ev_r.assign(L, std::vector<cudaEvent_t>(w->C, nullptr));
    for (int ci = 0; ci < w->C; ci++) {
        mkev(&w->ev_in[ci], false); mkev(&w->ev_fc[ci], false);
        for (int l = 0; l < L; l++) { mkev(&w->ev_g[l][ci], false); mkev(&w->ev_r[l][ci], false); }
    }
    mkev(&w->ev_start, false); mkev(&w->ev_done, false); mkev(&w->ev_t0, true); mkev(&w->ev_t1, true);
    if (st != GASR_OK) { set_error("wave engine: stream / event creation failed"); return st; }
    // function attributes and the decoder's workspaces now (both may synchronise the device), never between launches
    GASR_TRY(rnn_wide_prepare(ctx, H));
    if (w->pair) GASR_TRY(rnn_wide2_prepare(ctx));
    GASR_TRY(gemm_pair_prepare(ctx));
    {
        XsParams prep = {};
        XsMaps none;
        memset(&none, 0, sizeof(none));
        prep.n_targets = 1; prep.abort = w->flags + 2 * (size_t)w->max_blocks;
        XsTarget &t = prep.target[0];
        t.kind = XS_KIND_XPROJ; t.C = w->xp[0]; t.ldc = H; t.src_done = w->flags; t.dst_ready = w->flags + w->max_blocks;
        t.kblocks = 1; t.bn = TC_BN; t.terms = 3; t.n_tiles = 1;
        GASR_TRY(launch_xproj_stream(ctx, none, prep, 0, ctx->stream));
    }
    CtcArgs &ca = w->ca;
    ca.scores = w->logp; ca.domain = GASR_DOMAIN_LOG; ca.T = c.T; ca.N = c.N; ca.V = c.V; ca.ld = 32; ca.beam = c.beam; ca.blank = c.blank;
    ca.vocab_host = a->vocab.data(); ca.max_len = c.max_len; ca.nbest = c.nbest; ca.frame_rows = w->Npad;
    ca.warps_per_cta = ctx->opt.ctc_warps;
    GASR_TRY(ctc_decode_reserve(ctx, ca));
    GASR_TRY(ctc_decode_upload_vocab(ctx, ca, ctx->stream));
    ca.vocab_resident = true;
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    return GASR_OK;
}

void wave_destroy(gasr_asr *a) {
    WaveState *w = a->wave;
    if (!w) return;
    gasr_ctx *ctx = a->ctx;
    cudaDeviceSynchronize();
    for (void *p : w->h_planes) if (p) gasr_free_device(ctx, p);
    for (void *p : w->wih) if (p) gasr_free_device(ctx, p);
    for (void *p : w->whh) if (p) gasr_free_device(ctx, p);
    for (float *p : w->xp) if (p) gasr_free_device(ctx, p);
    for (void *p : {w->x_planes, w->fc_wbuf, (void *)w->fc_b_pad, (void *)w->bias_all, (void *)w->logp, (void *)w->logp_dense, (void *)w->flags})
        if (p) gasr_free_device(ctx, p);
    if (w->serial) {
        if (w->st_in) cudaStreamDestroy(w->st_in);
    } else {
        for (cudaStream_t s : {w->st_in, w->st_fc, w->st_dec}) if (s) cudaStreamDestroy(s);
        for (cudaStream_t s : w->st_g) if (s) cudaStreamDestroy(s);
        for (cudaStream_t s : w->st_r) if (s) cudaStreamDestroy(s);
    }
    for (cudaEvent_t e : w->ev_in) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : w->ev_fc) if (e) cudaEventDestroy(e);
    for (auto &v : w->ev_g) for (cudaEvent_t e : v) if (e) cudaEventDestroy(e);
    for (auto &v : w->ev_r) for (cudaEvent_t e : v) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {w->ev_start, w->ev_done, w->ev_t0, w->ev_t1}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : w->t0) cudaEventDestroy(e);
    for (cudaEvent_t e : w->t1) cudaEventDestroy(e);
    delete w;
    a->wave = nullptr;
}

// Called by gasr_asr_set_weights once the fp32 parameters are resident: bf16 planes of every weight matrix, the summed
// biases (b_hh + b_ih, RNN_Cell.cu:10), the padded output layer, and all TMA descriptors.
int wave_set_weights(gasr_asr *a, const float *fc_w, const float *fc_b) {
    gasr_ctx *ctx = a->ctx;
    WaveState *w = a->wave;
    const gasr_asr_config &c = a->cfg;
    const int L = c.L, H = c.H;
    cudaStream_t st = ctx->stream;
    w->rec.resize(L);
    w->maps.resize(L + 1);
    w->pmaps.resize(L);
    for (int l = 0; l < L; l++) {
        const int K = l == 0 ? c.in : H, Kp = l == 0 ? w->Kp0 : H;
        GASR_TRY(xproj_tc_prepare_weights(ctx, a->w_ih[l], K, H, w->wih[l], st));
        GASR_TRY(launch_matadd(ctx, a->b_ih[l], H, a->b_hh[l], H, w->bias_all + (size_t)l * H, H, 1, H, 1.0f, st));
        GASR_TRY(rnn_wide_plan(ctx, w->rec[l], a->w_hh[l], c.T, c.N, H, w->whh[l], w->h_planes[l], st, w->pair ? 256 : 128));
        unsigned char *ab = static_cast<unsigned char *>(l == 0 ? w->x_planes : w->h_planes[l - 1]);
        const size_t a_half = l == 0 ? planes_bytes(w->rows_p, Kp) / 2 : rnn_wide_plane_bytes(c.T, w->Npad, H);
        unsigned char *wb = static_cast<unsigned char *>(w->wih[l]);
        XsMaps &m = w->maps[l];
        memset(&m, 0, sizeof(m));
        GASR_TRY(tc_make_map(&m.m[0], ab, (int)w->rows_p, Kp, TC_BM));
        GASR_TRY(tc_make_map(&m.m[1], ab + a_half, (int)w->rows_p, Kp, TC_BM));
        const int bn = (ctx->opt.gemm_bn == 256 && H % 256 == 0) ? 256 : TC_BN;
        GASR_TRY(tc_make_map(&m.m[2], wb, H, Kp, bn));
        GASR_TRY(tc_make_map(&m.m[3], wb + xproj_tc_w_bytes(K, H) / 2, H, Kp, bn));
        XsMaps &pm = w->pmaps[l];
        memset(&pm, 0, sizeof(pm));
        pm.m[0] = m.m[0]; pm.m[1] = m.m[1];
        GASR_TRY(tc_make_map(&pm.m[2], wb, H, Kp, TC_BN));
        GASR_TRY(tc_make_map(&pm.m[3], wb + xproj_tc_w_bytes(K, H) / 2, H, Kp, TC_BN));
    }
    {
        // output layer as a 32-column target: W_fc padded to [H, 32] -> W^T hi/lo planes; bias padded with zeros
        std::vector<float> wpad((size_t)H * 32, 0.0f), bpad(32, 0.0f);
        for (int k = 0; k < H; k++) for (int v = 0; v < c.V; v++) wpad[(size_t)k * 32 + v] = fc_w[(size_t)k * c.V + v];
        for (int v = 0; v < c.V; v++) bpad[v] = fc_b[v];
        GASR_TRY(ws_reserve(ctx, ctx->ws_misc, sizeof(float) * (size_t)H * 32 + 256));
        GASR_TRY(gasr_memcpy_h2d(ctx, ctx->ws_misc.ptr, wpad.data(), sizeof(float) * (size_t)H * 32));
        GASR_TRY(xproj_tc_prepare_weights(ctx, static_cast<const float *>(ctx->ws_misc.ptr), H, 32, w->fc_wbuf, st));
        GASR_TRY(gasr_memcpy_h2d(ctx, w->fc_b_pad, bpad.data(), sizeof(float) * 32));
        unsigned char *ab = static_cast<unsigned char *>(w->h_planes[L - 1]);
        unsigned char *wb = static_cast<unsigned char *>(w->fc_wbuf);
        XsMaps &m = w->maps[L];
        memset(&m, 0, sizeof(m));
        GASR_TRY(tc_make_map(&m.m[0], ab, (int)w->rows_p, H, TC_BM));
        GASR_TRY(tc_make_map(&m.m[1], ab + rnn_wide_plane_bytes(c.T, w->Npad, H), (int)w->rows_p, H, TC_BM));
        GASR_TRY(tc_make_map(&m.m[2], wb, 32, H, 32));
        GASR_TRY(tc_make_map(&m.m[3], wb + xproj_tc_w_bytes(H, 32) / 2, 32, H, 32));
    }
    GASR_CUDA(cudaStreamSynchronize(st));
    return GASR_OK;
}

// one launch of the persistent tile engine over the rows [row0, row0 + rows) of a layer's A planes
static int wave_gemm(gasr_asr *a, int target, int ci, int row0, int rows, cudaStream_t st) {
    gasr_ctx *ctx = a->ctx;
    WaveState *w = a->wave;
    const gasr_asr_config &c = a->cfg;
    const int L = c.L, H = c.H;
    if (target < L && ctx->opt.gemm_pair && gemm_pair_supported(ctx, rows, H))
        return launch_gemm_pair(ctx, w->pmaps[target].m, row0, rows, target == 0 ? c.in : H, H,
                                w->xp[target] + (size_t)(ci % w->xp_slots) * w->Tc * w->Npad * H, H,
                                w->bias_all + (size_t)target * H, c.precision, st);
    XsParams p = {};
    p.M = rows; p.row0 = row0; p.n_blocks = ceil_div(rows, TC_BM); p.n_targets = 1;
    p.stages = ctx->opt.gemm_stages;
    p.abort = w->flags + 2 * (size_t)w->max_blocks; p.error = nullptr;
    XsTarget &t = p.target[0];
    t.src_done = w->flags; t.src_need = 1; t.dst_ready = w->flags + w->max_blocks;
    t.cta0 = 0;
    if (target < L) {
        const int K = target == 0 ? c.in : H;
        const int bn = (ctx->opt.gemm_bn == 256 && H % 256 == 0) ? 256 : TC_BN;   // 256-column tiles: the A block is read once per 256 outputs
        p.wide = bn == 256;
        t.kind = XS_KIND_XPROJ; t.n_tiles = H / bn; t.bn = bn; t.V = 0;
        t.kblocks = ceil_div(K, TC_BK); t.terms = c.precision == GASR_PREC_BF16 ? 1 : 3;
        t.C = w->xp[target] + (size_t)(ci % w->xp_slots) * w->Tc * w->Npad * H; t.ldc = H; t.bias = w->bias_all + (size_t)target * H;
    } else {
        t.kind = XS_KIND_LOGSOFTMAX; t.n_tiles = 1; t.bn = 32; t.V = c.V; t.kblocks = H / TC_BK; t.terms = 3;
        t.C = w->logp + (size_t)row0 * 32; t.ldc = 32; t.bias = w->fc_b_pad;
    }
    const int items = p.n_blocks * t.n_tiles;
    t.nctas = items < ctx->sm_count ? items : ctx->sm_count;
    return launch_xproj_stream(ctx, w->maps[target], p, t.nctas, st);
}

int wave_submit(gasr_asr *a, const float *x_dev, const float *x_host) {
    gasr_ctx *ctx = a->ctx;
    WaveState *w = a->wave;
    const gasr_asr_config &c = a->cfg;
    GASR_CHECK(!w->pending, "gasr_asr: a batch is already in flight on this pipeline (collect it first)");
    const int T = c.T, N = c.N, L = c.L, Npad = w->Npad;
    const bool prof = a->profile;
    w->n_timed = 0;
    auto timed_begin = [&](int tag, cudaStream_t st) -> int {
        if (!prof) return GASR_OK;
        if (w->n_timed == w->t0.size()) {
            cudaEvent_t e0, e1;
            if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return GASR_ERR_CUDA;
            w->t0.push_back(e0); w->t1.push_back(e1); w->tag.push_back(tag);
        }
        w->tag[w->n_timed] = tag;
        return cudaEventRecord(w->t0[w->n_timed], st) == cudaSuccess ? GASR_OK : GASR_ERR_CUDA;
    };
    auto timed_end = [&](cudaStream_t st) -> int {
        if (!prof) return GASR_OK;
        return cudaEventRecord(w->t1[w->n_timed++], st) == cudaSuccess ? GASR_OK : GASR_ERR_CUDA;
    };
    GASR_CUDA(cudaEventRecord(w->ev_t0, ctx->stream));
    GASR_CUDA(cudaEventRecord(w->ev_start, ctx->stream));
    for (cudaStream_t s : {w->st_in, w->st_fc, w->st_dec}) GASR_CUDA(cudaStreamWaitEvent(s, w->ev_start, 0));
    for (int l = 0; l < L; l++) {
        GASR_CUDA(cudaStreamWaitEvent(w->st_g[l], w->ev_start, 0));
        GASR_CUDA(cudaStreamWaitEvent(w->st_r[l], w->ev_start, 0));
    }
    const float *x_src = x_host ? a->x_dev : x_dev;
    __nv_bfloat16 *xh = static_cast<__nv_bfloat16 *>(w->x_planes);
    __nv_bfloat16 *xl = reinterpret_cast<__nv_bfloat16 *>(static_cast<unsigned char *>(w->x_planes) + planes_bytes(w->rows_p, w->Kp0) / 2);
    CtcArgs ca = w->ca;
    a->decode_extras(ca);
    for (int ci = 0; ci < w->C; ci++) {
        const int f0 = ci * w->Tc, f1 = (ci + 1) * w->Tc < T ? (ci + 1) * w->Tc : T;
        const int row0 = f0 * Npad, rows = (f1 - f0) * Npad;
        // ---- input: (copy +) split into bf16 planes -----------------------------------------------------------------
        if (x_host)
            GASR_CUDA(cudaMemcpyAsync(a->x_dev + (size_t)f0 * N * c.in, x_host + (size_t)f0 * N * c.in,
                                      sizeof(float) * (size_t)(f1 - f0) * N * c.in, cudaMemcpyHostToDevice, w->st_in));
        {
            const size_t total = (size_t)(f1 - f0) * N * (w->Kp0 / 2);
            int blocks = (int)((total + 255) / 256);
            if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
            GASR_TRY(timed_begin(0, w->st_in));
            wave_split_kernel<<<blocks, 256, 0, w->st_in>>>(x_src + (size_t)f0 * N * c.in, c.in, N, Npad, f1 - f0, c.in, w->Kp0,
                                                            xh + (size_t)row0 * w->Kp0, xl + (size_t)row0 * w->Kp0);
            GASR_CUDA(cudaGetLastError());
            ctx->launches += 1;
            GASR_TRY(timed_end(w->st_in));
        }
        GASR_CUDA(cudaEventRecord(w->ev_in[ci], w->st_in));
        // ---- layers: projection GEMM, then the recurrence ---------------------------------------------------------------
        for (int l = 0; l < L; l++) {
            GASR_CUDA(cudaStreamWaitEvent(w->st_g[l], l == 0 ? w->ev_in[ci] : w->ev_r[l - 1][ci], 0));
            if (ci >= w->xp_slots) GASR_CUDA(cudaStreamWaitEvent(w->st_g[l], w->ev_r[l][ci - w->xp_slots], 0));   // the xproj slot is free again
            GASR_TRY(timed_begin(0, w->st_g[l]));
            GASR_TRY(wave_gemm(a, l, ci, row0, rows, w->st_g[l]));
            GASR_TRY(timed_end(w->st_g[l]));
            GASR_CUDA(cudaEventRecord(w->ev_g[l][ci], w->st_g[l]));
            GASR_CUDA(cudaStreamWaitEvent(w->st_r[l], w->ev_g[l][ci], 0));
            RnnWideRun r = {};
            // the kernels address xproj as row t * Npad + n: bias the slot's base so that frame f0 lands on its first row
            r.s0 = f0; r.s1 = f1; r.ldxp = c.H; r.xp_rows_per_frame = Npad;
            r.xp = w->xp[l] + ((ptrdiff_t)(ci % w->xp_slots) * w->Tc - (ptrdiff_t)f0) * (ptrdiff_t)Npad * c.H;
            r.out = nullptr; r.groups_per_cluster = w->rec_groups; r.multicast = ctx->opt.rnn_mc;
            GASR_TRY(timed_begin(1, w->st_r[l]));
            GASR_TRY(w->pair ? launch_rnn_wide2(ctx, w->rec[l], r, w->st_r[l]) : launch_rnn_wide(ctx, w->rec[l], r, w->st_r[l]));
            GASR_TRY(timed_end(w->st_r[l]));
            GASR_CUDA(cudaEventRecord(w->ev_r[l][ci], w->st_r[l]));
        }
        // ---- output layer + log-softmax, decode ---------------------------------------------------------------------------
        GASR_CUDA(cudaStreamWaitEvent(w->st_fc, w->ev_r[L - 1][ci], 0));
        GASR_TRY(timed_begin(2, w->st_fc));
        GASR_TRY(wave_gemm(a, L, ci, row0, rows, w->st_fc));
        GASR_TRY(timed_end(w->st_fc));
        GASR_CUDA(cudaEventRecord(w->ev_fc[ci], w->st_fc));
        GASR_CUDA(cudaStreamWaitEvent(w->st_dec, w->ev_fc[ci], 0));
        ca.t0 = f0; ca.t1 = f1;
        GASR_TRY(timed_begin(3, w->st_dec));
        GASR_TRY(ctc_decode_launch(ctx, ca, w->st_dec));
        GASR_TRY(timed_end(w->st_dec));
    }
    GASR_CUDA(cudaEventRecord(w->ev_done, w->st_dec));
    GASR_CUDA(cudaStreamWaitEvent(ctx->stream, w->ev_done, 0));
    GASR_CUDA(cudaEventRecord(w->ev_t1, ctx->stream));
    w->pending = true;
    return GASR_OK;
}

int wave_collect(gasr_asr *a, char *out_paths, int *out_lens, float *out_scores) {
    gasr_ctx *ctx = a->ctx;
    WaveState *w = a->wave;
    GASR_CHECK(w->pending, "gasr_asr: no batch in flight");
    w->pending = false;
    {
        // Bounded wait: a pipeline that stops making progress becomes an error that names the stuck stages, not a hung process.
        const auto t_begin = std::chrono::steady_clock::now();
        const double limit_s = ctx->opt.wave_timeout_s;
        cudaError_t q;
        unsigned spins = 0;
        while ((q = cudaEventQuery(w->ev_t1)) == cudaErrorNotReady) {
            if ((++spins & 63u) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() > limit_s) {
                std::string busy;
                auto probe = [&](const char *name, int l, cudaStream_t s) {
                    if (cudaStreamQuery(s) == cudaErrorNotReady) { busy += name; if (l >= 0) busy += std::to_string(l); busy += ' '; }
                };
                probe("input", -1, w->st_in);
                for (int l = 0; l < a->cfg.L; l++) { probe("gemm", l, w->st_g[l]); probe("recurrence", l, w->st_r[l]); }
                probe("output-layer", -1, w->st_fc); probe("decoder", -1, w->st_dec);
                // first chunk whose event has not completed, per stage: the op that is stuck (or waiting for the stuck one)
                auto first_open = [&](const std::vector<cudaEvent_t> &ev) { int ci = 0; while (ci < (int)ev.size() && cudaEventQuery(ev[ci]) == cudaSuccess) ci++; return ci; };
                busy += "| first incomplete chunk: input " + std::to_string(first_open(w->ev_in));
                for (int l = 0; l < a->cfg.L; l++)
                    busy += " gemm" + std::to_string(l) + " " + std::to_string(first_open(w->ev_g[l])) + " recurrence" + std::to_string(l) + " " + std::to_string(first_open(w->ev_r[l]));
                busy += " output-layer " + std::to_string(first_open(w->ev_fc)) + " of " + std::to_string(w->C);
                gemm_pair_trace_dump();
                set_error("wave engine: no completion after %.0f s; streams still busy: %s", limit_s, busy.c_str());
                return GASR_ERR_CUDA;
            }
            if (spins > 256) std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        if (q != cudaSuccess) { set_error("wave engine: %s", cudaGetErrorString(q)); return GASR_ERR_CUDA; }
    }
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&w->last_ms, w->ev_t0, w->ev_t1);
    for (int i = 0; i < 4; i++) { a->stage_ms[i] = 0.0f; a->stage_launches[i] = 0; }
    for (size_t i = 0; i < w->n_timed; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, w->t0[i], w->t1[i]);
        a->stage_ms[w->tag[i]] += ms;
        a->stage_launches[w->tag[i]] += 1;
    }
    CtcArgs ca = w->ca;
    a->decode_extras(ca);
    ca.out_paths = out_paths; ca.out_lens = out_lens; ca.out_scores = out_scores; ca.out_counts = nullptr;
    return ctc_decode_finish(ctx, ca);
}

// dense [T*N, 32] copy of the padded log-prob matrix (parity checks)
int wave_logprobs(gasr_asr *a, const float **logp_dev, int *ldp) {
    gasr_ctx *ctx = a->ctx;
    WaveState *w = a->wave;
    const gasr_asr_config &c = a->cfg;
    *ldp = 32;
    if (w->Npad == c.N) { *logp_dev = w->logp; return GASR_OK; }
    if (!w->logp_dense) GASR_TRY(gasr_malloc_device(ctx, sizeof(float) * (size_t)c.T * c.N * 32, (void **)&w->logp_dense));
    GASR_CUDA(cudaMemcpy2DAsync(w->logp_dense, sizeof(float) * (size_t)c.N * 32, w->logp, sizeof(float) * (size_t)w->Npad * 32,
                                sizeof(float) * (size_t)c.N * 32, c.T, cudaMemcpyDeviceToDevice, ctx->stream));
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    *logp_dev = w->logp_dense;
    return GASR_OK;
}

// The decoder workspace grows when per-token timesteps are switched on (creation frame per trie node); a re-allocation
// drops the resident vocabulary, so it is uploaded again.  Not while a batch is in flight.
int wave_refresh_decoder(gasr_asr *a) {
    gasr_ctx *ctx = a->ctx;
    WaveState *w = a->wave;
    GASR_CHECK(!w->pending, "gasr_asr: a batch is in flight");
    GASR_CUDA(cudaDeviceSynchronize());
    CtcArgs ca = w->ca;
    a->decode_extras(ca);
    GASR_TRY(ctc_decode_reserve(ctx, ca));
    GASR_TRY(ctc_decode_upload_vocab(ctx, ca, ctx->stream));
    GASR_CUDA(cudaStreamSynchronize(ctx->stream));
    return GASR_OK;
}

bool wave_pending(const gasr_asr *a) { return a->wave && a->wave->pending; }

int wave_chunk_frames(const gasr_asr *a) { return a->wave ? a->wave->Tc : 0; }
float wave_last_ms(const gasr_asr *a) { return a->wave ? a->wave->last_ms : 0.0f; }

}  // namespace gasr
