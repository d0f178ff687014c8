// ctc_beam.cu -- host side of the CTC prefix beam search (CTCBeamSearch::decode, reference CTCBeamSearch.cu:262-312): workspace
// layout, argument checks, kernel dispatch, result unpacking.  The kernels: ctc_beam_general.cu / _warp.cu / _cta.cu / _cta2.cu;
// what they share: ctc_beam.cuh.
#include <stdlib.h>
#include <string.h>

#include "ctc_beam.cuh"

namespace gasr {

// shared memory of the general kernel (must mirror the carve-up at the top of ctc_beam_kernel)
static size_t ctc_smem_bytes(int B, int V, int Vp, int n_pad) {
    size_t s = sizeof(unsigned long long) * n_pad + sizeof(float) * Vp;
    s += (sizeof(float) + 2 * sizeof(int)) * 2 * B;   // score, node, pnode (x2 buffers)
    s += sizeof(int) * B;                             // newflag
    s += 3 * sizeof(short) * B;                       // twin, P0, P1
    s += sizeof(short) * 2 * B;                       // last (x2)
    s += 2 * sizeof(uint16_t) * (size_t)B * V;        // redir0/1
    s += (V + 3) / 4 * 4;                             // vocab chars
    s += 2 * B;                                       // eb (x2)
    s = (s + 7) / 8 * 8;
    s += sizeof(int) * 4 * B;                         // depth (x2), selected parent / label
    return s + 16;
}
static size_t ctc_rel_bytes(int B) { return sizeof(unsigned short) * 2 * (size_t)B * B; }

struct CtcLayout {
    int Vp, n_pad, cap, threads;
    size_t smem, off_vocab, off_parent, off_meta, off_anc, off_child, off_state, state_stride, off_paths, off_lens, off_scores,
        off_counts, off_stats, off_born, off_ts, total;
    size_t out_bytes;
};

static int ctc_layout(const CtcArgs &a, CtcLayout &L) {
    L.Vp = (a.V + 3) / 4 * 4;
    L.n_pad = 32;
    while (L.n_pad < a.beam * a.V) L.n_pad <<= 1;
    L.cap = 1 + a.beam * a.T;
    L.threads = L.n_pad / 2;
    if (L.threads < 128) L.threads = 128;
    if (L.threads > 1024) L.threads = 1024;
    L.smem = ctc_smem_bytes(a.beam, a.V, L.Vp, L.n_pad);
    size_t o = 0;
    L.off_vocab = o; o = align_up(o + a.V, 256);
    L.off_parent = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.off_meta = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.off_anc = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.off_child = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap * L.Vp, 256);
    L.state_stride = align_up((sizeof(CtaBeam<32>) > sizeof(WarpBeam<32>) ? sizeof(CtaBeam<32>) : sizeof(WarpBeam<32>)) + sizeof(int4), 256);
    L.off_state = o; o = align_up(o + L.state_stride * (size_t)a.N, 256);
    L.off_born = o; if (a.out_timesteps) o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.total = o;
    size_t q = 0;
    L.off_paths = q; q = align_up(q + (size_t)a.N * a.nbest * a.max_len, 256);
    L.off_lens = q; q = align_up(q + sizeof(int) * (size_t)a.N * a.nbest, 256);
    L.off_scores = q; q = align_up(q + sizeof(float) * (size_t)a.N * a.nbest, 256);
    L.off_counts = q; q = align_up(q + sizeof(int) * (size_t)a.N, 256);
    L.off_stats = q; q = align_up(q + 2 * sizeof(int) * (size_t)a.N, 256);
    L.off_ts = q; if (a.out_timesteps) q = align_up(q + sizeof(int) * (size_t)a.N * a.nbest * a.max_len, 256);
    L.out_bytes = q;
    return GASR_OK;
}

// Grow the decoder's workspaces for this problem size (allocation synchronises the device: the streaming pipeline
// calls this before its persistent kernels start).
int ctc_decode_reserve(gasr_ctx *ctx, const CtcArgs &a) {
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    GASR_TRY(ws_reserve(ctx, ctx->ws_ctc, L.total));
    GASR_TRY(ws_reserve(ctx, ctx->ws_out, L.out_bytes));
    GASR_TRY(pinned_reserve(ctx, L.out_bytes));
    return GASR_OK;
}

// Upload the vocabulary ahead of the launch (CtcArgs::vocab_resident then skips it): a small pageable copy issued at
// launch time would queue behind whatever the copy engine is busy with.
int ctc_decode_upload_vocab(gasr_ctx *ctx, const CtcArgs &a, cudaStream_t st) {
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    GASR_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(ctx->ws_ctc.ptr) + L.off_vocab, a.vocab_host, a.V, cudaMemcpyHostToDevice, st));
    return GASR_OK;
}

int ctc_decode_launch(gasr_ctx *ctx, const CtcArgs &a, cudaStream_t st) {
    GASR_CHECK(a.scores != nullptr && a.vocab_host != nullptr, "ctc_decode: null scores/vocab");
    GASR_CHECK(a.T >= 1 && a.N >= 0, "ctc_decode: T must be >= 1 and N >= 0 (T=%d N=%d)", a.T, a.N);
    GASR_CHECK(a.V >= 1 && a.V <= 255 && a.ld >= a.V, "ctc_decode: vocabulary size %d / ld %d unsupported", a.V, a.ld);
    GASR_CHECK(a.beam >= 1 && a.beam <= 1024, "ctc_decode: beam %d out of range", a.beam);
    GASR_CHECK(a.blank >= 0 && a.blank < a.V, "ctc_decode: blank id %d outside the vocabulary", a.blank);
    GASR_CHECK(a.nbest >= 1 && a.nbest <= a.beam && a.max_len >= 0, "ctc_decode: nbest/max_len out of range");
    GASR_CHECK(a.domain == GASR_DOMAIN_PROB || a.domain == GASR_DOMAIN_LOG, "ctc_decode: unknown score domain");
    for (int v = 0; v < a.V; v++) {
        GASR_CHECK(a.vocab_host[v] > 0, "ctc_decode: vocab chars must be in 1..127 (entry %d)", v);
        for (int u = 0; u < v; u++)
            GASR_CHECK(a.vocab_host[u] != a.vocab_host[v], "ctc_decode: duplicate vocab char at %d and %d", u, v);
    }
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    bool general_rel = false;
    if (!(a.beam <= 32 && a.V <= 32) && L.smem + ctc_rel_bytes(a.beam) <= (size_t)ctx->max_smem_optin) {
        general_rel = true;                           // the prefix-relation matrix fits: O(1) raw-string tie-breaks
        L.smem += ctc_rel_bytes(a.beam);
    }
    if (!(a.beam <= 32 && a.V <= 32) && L.smem > (size_t)ctx->max_smem_optin) {
        set_error("ctc_decode: beam %d x vocab %d needs %zu B of shared memory (> %d)", a.beam, a.V, L.smem,
                  ctx->max_smem_optin);
        return GASR_ERR_UNSUPPORTED;
    }
    GASR_TRY(ws_reserve(ctx, ctx->ws_ctc, L.total));
    GASR_TRY(ws_reserve(ctx, ctx->ws_out, L.out_bytes));
    GASR_TRY(pinned_reserve(ctx, L.out_bytes));
    unsigned char *ws = static_cast<unsigned char *>(ctx->ws_ctc.ptr);
    unsigned char *wo = static_cast<unsigned char *>(ctx->ws_out.ptr);
    const int t0 = a.t0, t1 = a.t1 > 0 ? a.t1 : a.T;
    const bool fast = a.beam <= 32 && a.V <= 32;
    GASR_CHECK(t0 >= 0 && t0 < t1 && t1 <= a.T, "ctc_decode: bad frame range [%d, %d)", t0, t1);
    GASR_CHECK(fast || (t0 == 0 && t1 == a.T), "ctc_decode: time-chunked decoding needs beam <= 32 and vocab <= 32");
    if (t0 == 0 && !a.vocab_resident) GASR_CUDA(cudaMemcpyAsync(ws + L.off_vocab, a.vocab_host, a.V, cudaMemcpyHostToDevice, st));

    CtcParams p;
    p.lens = a.lens_dev;
    p.born = nullptr; p.out_ts = nullptr;
    p.scores = a.scores; p.T = a.T; p.N = a.N; p.V = a.V; p.ld = a.ld; p.beam = a.beam; p.blank = a.blank;
    p.frame_rows = a.frame_rows > 0 ? a.frame_rows : a.N;
    GASR_CHECK(p.frame_rows >= a.N, "ctc_decode: frame_rows %d < N %d", p.frame_rows, a.N);
    p.Vp = L.Vp; p.n_pad = L.n_pad; p.cap = L.cap; p.max_len = a.max_len; p.nbest = a.nbest;
    p.vocab = reinterpret_cast<const char *>(ws + L.off_vocab);
    p.parent = reinterpret_cast<int *>(ws + L.off_parent);
    p.meta = reinterpret_cast<int *>(ws + L.off_meta);
    p.anc = reinterpret_cast<int *>(ws + L.off_anc);
    p.child = reinterpret_cast<int *>(ws + L.off_child);
    p.out_paths = reinterpret_cast<char *>(wo + L.off_paths);
    p.out_lens = reinterpret_cast<int *>(wo + L.off_lens);
    p.out_scores = reinterpret_cast<float *>(wo + L.off_scores);
    p.out_counts = reinterpret_cast<int *>(wo + L.off_counts);
    p.out_stats = reinterpret_cast<int *>(wo + L.off_stats);
    if (a.out_timesteps) { p.born = reinterpret_cast<int *>(ws + L.off_born); p.out_ts = reinterpret_cast<int *>(wo + L.off_ts); }
    if (!fast) GASR_CUDA(cudaMemsetAsync(p.out_stats, 0, 2 * sizeof(int) * (size_t)a.N, st));
    p.t0 = t0; p.t1 = t1; p.use_rel = general_rel ? 1 : 0;
    p.state = ws + L.off_state; p.state_stride = L.state_stride;
    p.lp_ready = a.lp_ready; p.lp_need = a.lp_need; p.lp_fpb = a.lp_fpb > 0 ? a.lp_fpb : 1; p.error = a.error; p.abort = a.abort;
    const char force_kc = ctx->opt.ctc_kernel;
    const char *force_k = force_kc ? &force_kc : nullptr;
    const bool use_cta = fast && (a.lp_ready != nullptr || (force_k ? (force_k[0] == 'c' || force_k[0] == 'd') : a.N <= 2 * ctx->sm_count));
    GASR_CHECK(a.lp_ready == nullptr || (fast && t0 == 0 && t1 == a.T), "ctc_decode: streaming needs beam <= 32, vocab <= 32, whole sequence");
    {
        // probe cells of the prune lower bound: the (parent rank, score rank) pairs with the smallest (i+1)(j+1)
        p.n_cells = (use_cta && a.beam > 16) ? 128 : 64;
        // warp kernel, beam <= 16: one probe cell per lane.  The bound is looser (21 instead of 16 survivors on peaky / random
        // log-probabilities, 26 instead of 19 on flat ones) but the all-pairs counting of 64 cells was 17 % of the kernel's
        // instructions: decoder alone 24.8 -> 23.6 ms per 4096 utterances, cfg5 step 118.4 -> 115.9 ms (same-box A/B, 3 runs each)
        if (!use_cta && fast && a.beam <= 16 && ctx->opt.ctc_cells != 64) p.n_cells = 32;
        if (!use_cta && fast && ctx->opt.ctc_cells == 32) p.n_cells = 32;      // GASR_CTC_CELLS = 32 | 64 forces either
        int taken = 0;
        for (int prod = 1; taken < p.n_cells && prod <= a.beam * a.V; prod++)
            for (int i = 0; i < a.beam && taken < p.n_cells; i++) {
                if (prod % (i + 1)) continue;
                const int j = prod / (i + 1) - 1;
                if (j >= a.V) continue;
                p.cell_i[taken] = (unsigned char)i; p.cell_j[taken] = (unsigned char)j; taken++;
            }
        for (; taken < 128; taken++) { p.cell_i[taken] = 255; p.cell_j[taken] = 0; }
        memset(p.cellmap, 255, sizeof(p.cellmap));
        for (int c = 0; c < p.n_cells; c++)
            if (p.cell_i[c] < 32 && p.cell_j[c] < 32) p.cellmap[p.cell_i[c] * 32 + p.cell_j[c]] = (unsigned char)c;
    }

    if (t0 == 0) GASR_CUDA(cudaMemsetAsync(wo + L.off_paths, 0, (size_t)a.N * a.nbest * a.max_len, st));
    if (t0 == 0 && a.out_timesteps) GASR_CUDA(cudaMemsetAsync(wo + L.off_ts, 0, sizeof(int) * (size_t)a.N * a.nbest * a.max_len, st));
    const bool whole = t0 == 0 && t1 == a.T;
    // ---- dispatch --------------------------------------------------------------------------------------------------------
    //   beam <= 32 and vocabulary <= 32 ("fast"):
    //     few utterances (<= 2 per SM) or the streaming pipeline:   whole sequence -> ctc_beam_cta2.cu (main + fetch + trie warps)
    //                                                                 time chunk     -> ctc_beam_cta.cu  (128-thread CTA, resumable)
    //     many utterances:                                           ctc_beam_warp.cu (one warp per utterance, resumable)
    //   anything else (beam <= 1024, vocabulary <= 255):             ctc_beam_general.cu (whole sequences only)
    //   GASR_CTC_KERNEL = w | c | d forces the warp / CTA / second-generation CTA kernel inside the fast envelope.
    if (use_cta && whole && !(force_k && force_k[0] == 'c')) {
        // no shared-memory padding: in the streaming pipeline these CTAs may share an SM with a GEMM CTA (whose few threads
        // leave the issue slots free); the recurrence CTAs fill their SMs' register file, so nothing lands there
        GASR_TRY(ctc_launch_cta2(p, a.domain, a.N, ctx->opt.ctc_mw, (size_t)ctx->opt.ctc_pad, st));
    } else if (use_cta) {
        // 20 KB of (unused) dynamic shared memory keeps these CTAs off the SMs whose shared memory is filled by a recurrence
        // CTA (208 KB) when the pipeline runs both at once
        GASR_TRY(ctc_launch_cta(p, a.domain, a.N, 20480, st));
    } else if (fast) {
        // few warps per CTA when utterances are scarce (latency), 8 when plentiful
        int W = ceil_div(a.N, ctx->sm_count);
        if (W > 8) W = 8;
        if (a.warps_per_cta >= 1 && a.warps_per_cta <= 8) W = a.warps_per_cta;
        GASR_TRY(ctc_launch_warp(p, a.domain, ceil_div(a.N, W), W, st));
    } else {
        GASR_TRY(ctc_launch_general(p, a.domain, a.N, L.threads, L.smem, st));
    }
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (t1 == a.T) GASR_CUDA(cudaMemcpyAsync(ctx->pinned_out, wo, L.out_bytes, cudaMemcpyDeviceToHost, st));
    return GASR_OK;
}

int ctc_decode_finish(gasr_ctx *ctx, const CtcArgs &a) {
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    const unsigned char *h = static_cast<const unsigned char *>(ctx->pinned_out);
    memcpy(a.out_paths, h + L.off_paths, (size_t)a.N * a.nbest * a.max_len);
    memcpy(a.out_lens, h + L.off_lens, sizeof(int) * (size_t)a.N * a.nbest);
    memcpy(a.out_scores, h + L.off_scores, sizeof(float) * (size_t)a.N * a.nbest);
    if (a.out_counts) memcpy(a.out_counts, h + L.off_counts, sizeof(int) * (size_t)a.N);
    if (a.out_timesteps) memcpy(a.out_timesteps, h + L.off_ts, sizeof(int) * (size_t)a.N * a.nbest * a.max_len);
    {
        const int *stt = reinterpret_cast<const int *>(h + L.off_stats);
        ctx->ctc_fallback_frames = 0; ctx->ctc_survivors = 0;
        for (int n = 0; n < a.N; n++) { ctx->ctc_fallback_frames += stt[2 * n]; ctx->ctc_survivors += stt[2 * n + 1]; }
    }
    int status = GASR_OK;
    for (size_t i = 0; i < (size_t)a.N * a.nbest; i++)
        if (a.out_lens[i] > a.max_len) status = GASR_ERR_TRUNCATED;
    if (status != GASR_OK) set_error("ctc_decode: at least one path is longer than max_len=%d", a.max_len);
    return status;
}

}  // namespace gasr
