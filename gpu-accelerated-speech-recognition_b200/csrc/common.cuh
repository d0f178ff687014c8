// common.cuh -- context, error plumbing and launch helpers shared by the kernels of libgasr.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "gasr.h"

namespace gasr {

void set_error(const char *fmt, ...);

#define GASR_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            gasr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return GASR_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

#define GASR_CHECK(cond, ...)                                                                    \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            gasr::set_error(__VA_ARGS__);                                                        \
            return GASR_ERR_INVALID;                                                             \
        }                                                                                        \
    } while (0)

#define GASR_TRY(expr)                                                                           \
    do {                                                                                         \
        int s__ = (expr);                                                                        \
        if (s__ != GASR_OK) return s__;                                                          \
    } while (0)

// A grow-only device scratch block (reused across calls; freed with the ctx).
struct Workspace {
    void *ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace gasr

// Environment switches, read ONCE when the context is created (never on a per-call path).
struct gasr_options {
    char rnn = 0;            // GASR_RNN: w = wide tcgen05 recurrence, f / m / ... = the round-1 kernels (first letter)
    int gru_pp = 0;          // GASR_GRU_PP: row blocks per CTA of the persistent GRU recurrence (1 or 2; 0 = by batch size)
    int gru_units = 0;       // GASR_GRU_UNITS: hidden units per CTA of the persistent GRU recurrence (16 or 32; 0 = automatic)
    int rnn_mc = 1;          // GASR_RNN_MC: TMA multicast of the h boxes in the wide recurrence
    int rnn_groups = 0;      // GASR_RNN_G: groups of utterances per cluster (1 or 2; 0 = by batch size)
    int rnn_pair = 1;        // GASR_RNN_PAIR: CTA-pair recurrence (tcgen05.mma.cta_group::2, groups of 256 utterances)
    char ctc_kernel = 0;     // GASR_CTC_KERNEL
    int ctc_mw = 8;          // GASR_CTC_MW
    int ctc_cells = 0;       // GASR_CTC_CELLS: probe cells of the warp decoder's prune bound (32 or 64; 0 = 32 for beam <= 16, else 64)
    int ctc_pad = 0;         // GASR_CTC_PAD
    char gru = 0;            // GASR_GRU: t = per-timestep tcgen05 kernel even in bf16 mode, g = three launches per step, s = SIMT step kernel
    bool gru_no_pdl = false, no_graph = false, bidir_serial = false, linear_simt = false;
    char xproj = 0;          // GASR_XPROJ
    int chunk = -1;          // GASR_CHUNK (-1: default)
    int stream = -1;         // GASR_STREAM (-1: default)
    int wave = -1;           // GASR_WAVE (-1: default): throughput (wave) engine on/off
    int stream_gemm_ctas = 24;
    int rnn_nsub = -1;
    int gemm_stages = 3;     // GASR_GEMM_STAGES: ring depth of the wave engine's GEMM launches (2 or 3)
    int gemm_pair = 1;       // GASR_GEMM_PAIR: projection GEMMs on CTA pairs (256 x 256 tiles); 0 = one-CTA tile engine
    int gemm_bn = 256;       // GASR_GEMM_BN: tile width of the wave engine's projection GEMMs (128 or 256)
    int wave_timeout_s = 60;  // GASR_WAVE_TIMEOUT_S: a batch that has not completed after this many seconds is reported as an error
    int wave_prio = -1;       // GASR_WAVE_PRIO: stream priority scheme of the wave engine (-1: by batch size, see asr_wave.cu)
    bool wave_serial = false; // GASR_WAVE_SERIAL: diagnostic, all stages of the wave engine on one stream
    int ctc_warps = 8;       // GASR_CTC_WARPS: utterances (warps) per decoder CTA in the wave engine (0: balanced automatically)
};

struct gasr_ctx {
    gasr_options opt;
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    int cluster_ok = 0;
    cudaStream_t stream = nullptr;       // main stream: every kernel of the C ABI is launched here
    cudaStream_t side[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // pipeline side streams
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    long long launches = 0;
    size_t device_bytes = 0, host_bytes = 0;
    std::map<void *, size_t> dev_blocks, host_blocks;
    gasr::Workspace ws_ctc, ws_rnn, ws_misc, ws_out, ws_gru, ws_lin, ws_lens;
    gasr::Workspace ws_wide;             // planes + W_hh^T planes of the wide recurrence (C-ABI path)
    gasr::Workspace ws_rnn_b, ws_misc_b, ws_gru_b;   // second set: the backward direction of a bidirectional layer runs concurrently
    int ws_sel = 0;                      // which set the recurrent-layer helpers use (0 / 1)
    // instantiated CUDA graphs of launch-bound per-timestep loops (GRU recurrence), keyed by their operands
    struct StepGraph { const void *k[5]; int dims[8]; cudaGraphExec_t exec; };
    std::vector<StepGraph> step_graphs;
    cudaEvent_t ev_bi[2] = {nullptr, nullptr};
    void *pinned_out = nullptr;          // pinned staging for decode results
    size_t pinned_out_bytes = 0;
    long long ctc_fallback_frames = 0, ctc_survivors = 0;   // diagnostics of the last decode
    unsigned attr_mask = 0;              // kernels whose function attributes are set (setting them is not stream-safe)
};

namespace gasr {

int ws_reserve(gasr_ctx *ctx, Workspace &ws, size_t bytes);
int pinned_reserve(gasr_ctx *ctx, size_t bytes);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// kernels (one translation unit each) -----------------------------------------------------------------
int launch_matmul(gasr_ctx *ctx, const float *x, int ldx, int tx, const float *y, int ldy, int ty, float *z, int ldz,
                  int m, int k, int n, const float *bias, cudaStream_t st);
int launch_matadd(gasr_ctx *ctx, const float *x, int ldx, const float *y, int ldy, float *z, int ldz, int rows,
                  int cols, float lambda, cudaStream_t st);
int launch_linear(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *b, float *y, int ldy, int rows,
                  int in, int out, int act, cudaStream_t st);
int launch_log_softmax(gasr_ctx *ctx, const float *x, int ldx, float *y, int ldy, int rows, int cols, cudaStream_t st);
int launch_rnn_cell(gasr_ctx *ctx, const float *x, const float *h_prev, const float *w_ih, const float *w_hh,
                    const float *b_ih, const float *b_hh, float *out, int batch, int in, int hidden, cudaStream_t st);

// tcgen05 / TMEM / TMA projection GEMM (xproj_gemm_tc.cu)
bool xproj_tc_supported(int M, int K, int N);
size_t xproj_tc_a_bytes(int M, int K);      // scratch for the bf16 hi/lo planes of A
size_t xproj_tc_w_bytes(int K, int N);      // prepared (transposed, split) weights
int xproj_tc_prepare_weights(gasr_ctx *ctx, const float *W, int K, int N, void *wbuf, cudaStream_t st);
int xproj_tc_split_rows(gasr_ctx *ctx, const float *A, int lda, int M, int K, void *abuf, cudaStream_t st);
struct XprojTcPlan { CUtensorMap maps[4]; int M, K, N; void *abuf; };
int xproj_tc_plan(XprojTcPlan &pl, int M, int K, int N, const void *wbuf, void *abuf);
int xproj_tc_run(gasr_ctx *ctx, const XprojTcPlan &pl, const float *A, int lda, const float *bias, float *C, int ldc,
                 int precision, cudaStream_t st, int relu = 0);
int xproj_tc_split_rows_range(gasr_ctx *ctx, const float *A, int lda, int M_total, int row0, int nrows, int K, void *abuf,
                              cudaStream_t st);
int launch_xproj_tc(gasr_ctx *ctx, const float *A, int lda, int M, int K, int N, const void *wbuf, void *abuf,
                    const float *bias, float *C, int ldc, int precision, cudaStream_t st, int relu = 0);

// fused GRU timestep (gru_tc.cu): permuted W_hh^T planes + ping-pong bf16 planes of h, TMA descriptors built once
struct GruTcPlan { CUtensorMap maps[2][2]; CUtensorMap wmaps[2]; void *plane[2][2]; int N, H, Kp; };
bool gru_tc_supported(int N, int H, int ldxp, int ldo, int col0);
size_t gru_tc_w_bytes(int H);
size_t gru_tc_plane_bytes(int N, int H);
int gru_tc_prepare(gasr_ctx *ctx, GruTcPlan &pl, const float *w_hh, int N, int H, void *wbuf, void *planes, cudaStream_t st);
int gru_tc_step(gasr_ctx *ctx, const GruTcPlan &pl, int src, const float *xp, int ldxp, const float *b_hh, const float *hprev,
                int ldh, float *out, int ldo, bool overlap, cudaStream_t st);

// Linear (<= 32 outputs) + log-softmax over many rows on the tensor cores: operand split + one launch of the streaming
// tile engine with a single output-layer target (xproj_stream.cu); y must have ldy >= 32, ldy % 4 == 0
bool linear_tc_supported(int rows, int in, int out, int ldy, const float *y, int act);
int launch_linear_logsoftmax_tc(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *b, float *y, int ldy,
                                int rows, int in, int out, cudaStream_t st);

struct RnnLayerArgs {
    int cell, T, N, H, reverse;
    const float *xproj;   // [T*N, ldxp]: x*W_ih + (b_ih [+ b_hh for tanh]) for this direction
    int ldxp;
    const float *w_hh;    // [H, G*H]
    const float *b_hh;    // [G*H] (GRU only: b_hn stays inside the reset product)
    float *out;           // [T*N, ldo] written at column offset col0
    int ldo, col0;
    int s0 = 0, s1 = 0;   // steps [s0, s1) only (s1 = 0: all T); h before step s0 is read back from `out`
    int precision = GASR_PREC_FP32;   // GASR_PREC_BF16: the GRU recurrence may use single-plane fp16 operands (gru_seq.cu)
    bool concurrent = false;          // another recurrence (the other direction of the layer) runs at the same time
};
int launch_rnn_recurrence(gasr_ctx *ctx, const RnnLayerArgs &a, cudaStream_t st);

// wide hidden layer, a handful of utterances: persistent fp32 kernel with W_hh resident in shared memory (rnn_resident.cu)
bool rnn_resident_supported(const gasr_ctx *ctx, const RnnLayerArgs &a);
int launch_rnn_resident(gasr_ctx *ctx, const RnnLayerArgs &a, void *ws, cudaStream_t st);

// persistent GRU recurrence with W_hh resident in shared memory (gru_seq.cu; GASR_PREC_BF16 mode)
bool gru_seq_supported(const gasr_ctx *ctx, int T, int N, int H, int ldxp, int ldo, int col0);
size_t gru_seq_ws_bytes(int N, int H);
int launch_gru_seq(gasr_ctx *ctx, const RnnLayerArgs &a, void *ws, cudaStream_t st);

struct CtcArgs {
    const float *scores; int domain, T, N, V, ld, beam, blank; const char *vocab_host; int max_len, nbest;
    char *out_paths; int *out_lens; float *out_scores; int *out_counts;   // host
    int t0 = 0, t1 = 0;   // frames [t0, t1) only (t1 = 0: all T); the beam is parked in the ctx workspace between chunks
    // streaming pipeline: scores of frame t may be read once lp_ready[t / lp_fpb] >= lp_need (device counters)
    const unsigned *lp_ready = nullptr; int lp_need = 0, lp_fpb = 1; int *error = nullptr; volatile unsigned *abort = nullptr;
    bool vocab_resident = false;   // the vocabulary was uploaded by ctc_decode_upload_vocab
    int frame_rows = 0;            // rows of `scores` per frame (0: N)
    int warps_per_cta = 0;         // warp kernel: utterances per CTA (0: automatic)
    const int *lens_dev = nullptr; // frames per utterance (device, N ints; null: all T) -- baseline/main.py:45-46 out_lens
    int *out_timesteps = nullptr;  // host [N, nbest, max_len]: frame at which each output token's prefix first entered the beam (null: not wanted)
};
int ctc_decode_reserve(gasr_ctx *ctx, const CtcArgs &a);                  // allocations only (device-synchronising)
int ctc_decode_upload_vocab(gasr_ctx *ctx, const CtcArgs &a, cudaStream_t st);
int ctc_decode_launch(gasr_ctx *ctx, const CtcArgs &a, cudaStream_t st);   // enqueue kernel + D2H into pinned staging
int ctc_decode_finish(gasr_ctx *ctx, const CtcArgs &a);                    // after stream sync: unpack to caller buffers

}  // namespace gasr
