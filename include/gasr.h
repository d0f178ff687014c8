/*
 * gasr.h -- C ABI of the B200-native speech-recognition hot path
 *           (RNN acoustic-model forward -> Linear -> log-softmax -> CTC prefix beam search).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference
 * (jrxk/GPU-Accelerated-Speech-Recognition) has no C ABI -- its callers `new` C++ module classes
 * (main.cpp:31-45,64-72; nn_test.cpp:13-17,62-68).  Each entry point below names the reference
 * interface it stands behind; include/{cuMatrix,Linear,RNN_Cell,RNN,CTCBeamSearch,MemoryMonitor}.h
 * re-create those classes (same names, constructors, methods) as inline wrappers over this ABI.
 *
 * Conventions: every function returns a gasr_status (0 = ok) and never exits the process (the
 * reference printf+exit(0)s, cuMatrix.cpp:37-42); gasr_last_error() gives the message.  A gasr_ctx
 * owns one device, one main CUDA stream and the scratch workspaces; it is re-entrant per ctx (one
 * ctx per GPU / host thread -- no process-global singletons like getHandle(), cuMatrix.cpp:18-30).
 * Matrices are row-major with an explicit leading dimension `ld` in elements.  There is no CPU
 * fallback: with no usable CUDA device gasr_ctx_create fails with GASR_ERR_CUDA.
 */
#ifndef GASR_H
#define GASR_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gasr_ctx gasr_ctx;

typedef enum {
    GASR_OK = 0,
    GASR_ERR_INVALID = 1,      /* bad argument / shape mismatch (reference: printf + exit(0))   */
    GASR_ERR_CUDA = 2,         /* CUDA runtime error or no device                                */
    GASR_ERR_NOMEM = 3,
    GASR_ERR_UNSUPPORTED = 4,
    GASR_ERR_TRUNCATED = 5     /* an output string did not fit max_len (length is still reported) */
} gasr_status;

enum { GASR_ACT_NONE = 0, GASR_ACT_RELU = 1, GASR_ACT_LOGSOFTMAX = 2 };
enum { GASR_DOMAIN_PROB = 0, GASR_DOMAIN_LOG = 1 };
enum { GASR_CELL_TANH = 0, GASR_CELL_GRU = 1 };
enum { GASR_PREC_FP32 = 0, GASR_PREC_BF16 = 1 };

int gasr_version(void);
const char *gasr_last_error(void);              /* thread-local message of the last failing call */
int gasr_device_count(int *count);

/* ---- context ------------------------------------------------------------------------------------ */
int gasr_ctx_create(int device, gasr_ctx **ctx);
int gasr_ctx_destroy(gasr_ctx *ctx);
int gasr_ctx_sync(gasr_ctx *ctx);               /* wait for everything queued on the ctx           */
int gasr_ctx_sm_count(gasr_ctx *ctx, int *sms);
/* CUDA-event timer on the ctx's main stream (the stream every kernel below is launched on).       */
int gasr_timer_start(gasr_ctx *ctx);
int gasr_timer_stop(gasr_ctx *ctx, float *elapsed_ms);   /* records, synchronises, returns ms      */
/* Number of kernels this library has launched on the ctx since creation (bench.py gpu_launches).  */
int gasr_ctx_launch_count(gasr_ctx *ctx, long long *launches);

/* ---- memory: cuMatrix.h:201-228 (mallocHost/mallocDev), MemoryMonitor.cpp:9-51 ------------------- */
/* Device blocks are 256-byte aligned and zero-filled (cuMatrix::mallocDev memsets, cuMatrix.h:221). */
int gasr_malloc_device(gasr_ctx *ctx, size_t bytes, void **ptr);
int gasr_free_device(gasr_ctx *ctx, void *ptr);
int gasr_malloc_host(gasr_ctx *ctx, size_t bytes, void **ptr);     /* pinned, zero-filled           */
int gasr_free_host(gasr_ctx *ctx, void *ptr);
int gasr_memcpy_h2d(gasr_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);  /* toGpu  */
int gasr_memcpy_d2h(gasr_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);  /* toCpu  */
int gasr_memcpy_h2d_async(gasr_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int gasr_memcpy_d2h_async(gasr_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
/* cuMatrix::toGpu(cudaStream_t) (cuMatrix.h:108-115): upload on the caller's stream; cuda_stream is a cudaStream_t passed
 * as an opaque pointer (NULL = the ctx's own stream).                                                              */
int gasr_memcpy_h2d_on_stream(gasr_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, void *cuda_stream);
int gasr_memset_device(gasr_ctx *ctx, void *dst_dev, int value, size_t bytes);           /* gpuClear */
int gasr_memory_stats(gasr_ctx *ctx, size_t *device_bytes, size_t *host_bytes);  /* print*Memory   */

/* ---- dense math: cuMatrix.cpp:33-168 ------------------------------------------------------------- */
/* z[m,n] = op(x) * op(y), fp32.  trans_x/trans_y select matrixMul / matrixMulTA / matrixMulTB.     */
int gasr_matmul(gasr_ctx *ctx, const float *x, int ldx, int trans_x, const float *y, int ldy, int trans_y,
                float *z, int ldz, int m, int k, int n);
/* z = x + lambda * y (matrixAdd, cublasSgeam).                                                     */
int gasr_matadd(gasr_ctx *ctx, const float *x, int ldx, const float *y, int ldy, float *z, int ldz, int rows,
                int cols, float lambda);

/* ---- batched input projection (the x*W_ih of RNN_Cell.cu:66, hoisted over all frames) --------------- */
/* y[rows,out] = x[rows,in] * W[in,out] + bias on the tcgen05 tensor cores when out % 128 == 0 (operands split
 * into bf16 hi/lo planes: GASR_PREC_FP32 accumulates hi*hi + hi*lo + lo*hi, GASR_PREC_BF16 hi*hi only);
 * other shapes use the fp32 FFMA GEMM.  bias may be NULL.                                                 */
int gasr_xproj_gemm(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *bias, float *y, int ldy,
                    int rows, int in, int out, int precision);

/* ---- Linear: Linear.cu:3-10,42-49 (+ log-softmax of baseline/model.py:49) ------------------------ */
/* y[rows,out] = act(x[rows,in] * W[in,out] + b), one fused kernel.  act = GASR_ACT_*.               */
int gasr_linear_forward(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *b, float *y,
                        int ldy, int rows, int in, int out, int act);
int gasr_log_softmax(gasr_ctx *ctx, const float *x, int ldx, float *y, int ldy, int rows, int cols);

/* ---- RNN_Cell / RNN: RNN_Cell.cu:5-13,65-74; RNN.cu:9-30 ---------------------------------------- */
/* One Elman step: out = tanh(x*W_ih + h_prev*W_hh + (b_hh + b_ih)).  All pointers are device memory. */
int gasr_rnn_cell_forward(gasr_ctx *ctx, const float *x, const float *h_prev, const float *w_ih,
                          const float *w_hh, const float *b_ih, const float *b_hh, float *out, int batch,
                          int in, int hidden);
/*
 * Whole stacked recurrent forward.  x: [T*N, in] time-major (row t*N+n, RNN.cu:17), dense (ld = in).
 * Parameter arrays hold device pointers indexed [l * D + d] (D = 2 when bidirectional):
 *   w_ih [in_l, G*H], w_hh [H, G*H], b_ih/b_hh [G*H]   (G = 1 tanh, 3 GRU with gate order r,z,n)
 * hiddens[l]: device [T*N, D*H], every layer's full hidden sequence (RNN.h:18); the last one is
 * what RNN::forward returns.  h_0 = 0 (RNN.h:16-17).  precision selects the tensor-core arithmetic:
 * GASR_PREC_FP32: fp32-grade everywhere (three-term bf16 split, 2^-16 per product);
 * GASR_PREC_BF16: input projections with bf16 operands and, for GRU stacks, the recurrence with fp16
 * operands (one persistent launch per layer and direction); fp32 accumulation, gates and state.  Stated
 * tolerance of that mode: 2e-2 on the log-probabilities.
 */
int gasr_rnn_forward(gasr_ctx *ctx, int cell, int bidirectional, int T, int N, int in, int H, int L,
                     const float *const *w_ih, const float *const *w_hh, const float *const *b_ih,
                     const float *const *b_hh, const float *x, float *const *hiddens, int precision);

/* ---- CTCBeamSearch: CTCBeamSearch.h:107-131, CTCBeamSearch.cu:262-312 ---------------------------- */
/*
 * scores: device [T*N, ld] time-major, first V columns used; probabilities (GASR_DOMAIN_PROB, the
 * reference's arithmetic: fp32 multiply / add) or log-probabilities (GASR_DOMAIN_LOG: fp32 add /
 * deterministic log-add-exp).  vocab: V distinct chars in 1..127, vocab[blank] is the blank.
 * Host outputs, best beam first: out_paths [N, nbest, max_len] bytes (not NUL-terminated),
 * out_lens / out_scores [N, nbest], out_counts [N] (kept beams, may be NULL).  nbest = 1 is the
 * reference's result (top-1 string and its merged score, CTCBeamSearch.cu:290-298).
 */
int gasr_ctc_decode(gasr_ctx *ctx, const float *scores, int domain, int T, int N, int V, int ld, int beam,
                    int blank, const char *vocab, int max_len, int nbest, char *out_paths, int *out_lens,
                    float *out_scores, int *out_counts);
/*
 * The same decoder with the two optional features baseline/main.py:45-46 takes from its decoder
 * (decoder.decode(output, out_lens) -> output, scores, timesteps, out_seq_len):
 *   lens_host [N] (may be NULL): frames of each utterance, 1..T.  Utterance n is decoded as if T were lens_host[n]:
 *       frames beyond it are never read and the last-frame rule (CTCBeamSearch.cu:452-456) applies at frame lens[n]-1.
 *   out_timesteps [N, nbest, max_len] (host, may be NULL): for output character j, the frame at which the prefix
 *       path[0..j] first entered the kept beam (0-based; entries beyond the path's length are 0).
 */
int gasr_ctc_decode_ex(gasr_ctx *ctx, const float *scores, int domain, int T, int N, int V, int ld, int beam,
                       int blank, const char *vocab, int max_len, int nbest, const int *lens_host, char *out_paths,
                       int *out_lens, float *out_scores, int *out_counts, int *out_timesteps);
/* Diagnostics of the last decode on this ctx: utterance-frames whose prune exceeded the 64-survivor fast path,
 * and the sum of prune survivors over all utterance-frames (mean survivors = that / (N*T)).             */
int gasr_ctc_last_stats(gasr_ctx *ctx, long long *fallback_frames, long long *survivors);
/* Same with the scores in host memory (copied to the device inside the call).                       */
int gasr_ctc_decode_host(gasr_ctx *ctx, const float *scores_host, int domain, int T, int N, int V, int beam,
                         int blank, const char *vocab, int max_len, int nbest, char *out_paths, int *out_lens,
                         float *out_scores, int *out_counts);

/* ---- fused pipeline: RNN stack -> Linear -> log-softmax -> CTC beam search ----------------------- */
typedef struct gasr_asr gasr_asr;
typedef struct {
    int cell;            /* GASR_CELL_*                       */
    int bidirectional;
    int T, N;            /* frames, utterances per batch      */
    int in, H, L;        /* feature dim, hidden size, layers  */
    int V;               /* vocabulary incl. blank            */
    int beam, blank;
    int precision;       /* GASR_PREC_*                       */
    int nbest;
    int max_len;
} gasr_asr_config;
int gasr_asr_create(gasr_ctx *ctx, const gasr_asr_config *cfg, const char *vocab, gasr_asr **asr);
int gasr_asr_destroy(gasr_asr *asr);
/* Host parameter arrays in reference layout, indexed like gasr_rnn_forward; fc_w [D*H, V], fc_b [V]. */
int gasr_asr_set_weights(gasr_asr *asr, const float *const *w_ih, const float *const *w_hh,
                         const float *const *b_ih, const float *const *b_hh, const float *fc_w,
                         const float *fc_b);
/* x_host: [T*N, in] in (pinned or pageable) host memory; copies in, runs, copies results out, syncs. */
int gasr_asr_run_host(gasr_asr *asr, const float *x_host, char *out_paths, int *out_lens, float *out_scores);
/* Inputs already resident in HBM (x_dev [T*N, in]); results still land in host memory.              */
int gasr_asr_run_device(gasr_asr *asr, const float *x_dev, char *out_paths, int *out_lens, float *out_scores);
/* Device pointer of the log-prob matrix [T*N, ldp] the last run produced (for parity checks).       */
int gasr_asr_logprobs(gasr_asr *asr, const float **logp_dev, int *ldp);
/* Per-stage device times (ms) of the last run: projection, recurrence, linear+log-softmax, decode.  */
int gasr_asr_stage_times(gasr_asr *asr, float *ms4);
/* Kernel launches per stage in the last run (same order) and the execution mode in *chunk_frames:
 *   0   the stages ran back to back on one stream;
 *   >0  time chunks of that many frames flowed through the stages on concurrent streams (a stage time is the sum of
 *       its launches' durations);
 *   -2  wave engine (default wherever it applies: unidirectional tanh stacks, H in {128, 256, 512}, beam / vocabulary <= 32):
 *       groups of 128 utterances, time chunks on per-stage streams ordered by events; a stage time is the sum of its launches'
 *       durations (only with gasr_asr_profile on);
 *   -1  streaming (opt-in, GASR_STREAM=1): one persistent kernel per stage (recurrence of all layers / projection + output-layer GEMM / decoder)
 *       ran concurrently for the whole sequence, coupled by progress counters in HBM; a stage time is the duration of
 *       its kernel and the Linear + log-softmax stage is part of the GEMM kernel (reported as 0).                  */
int gasr_asr_stage_launches(gasr_asr *asr, int *n4, int *chunk_frames);
/*
 * Asynchronous form (wave engine; GASR_ERR_UNSUPPORTED for configurations outside it): submit enqueues one batch on the
 * pipeline's streams and returns; collect waits for it and unpacks transcripts / lengths / scores.  One batch in flight
 * per gasr_asr; several gasr_asr objects (each on its own gasr_ctx) overlap on one GPU -- gasr_job_* below does that.
 */
int gasr_asr_submit_host(gasr_asr *asr, const float *x_host);
int gasr_asr_submit_device(gasr_asr *asr, const float *x_dev);
int gasr_asr_collect(gasr_asr *asr, char *out_paths, int *out_lens, float *out_scores);
/* Variable-length batches (baseline/main.py:45 out_lens): lens_host [N], 1..T, applies to every later run; NULL switches it off.
 * The acoustic model still runs all T frames of the padded batch (a unidirectional stack never looks ahead, so frames
 * < lens[n] are unaffected; bidirectional stacks are GASR_ERR_INVALID); the decoder stops at lens[n].                  */
int gasr_asr_set_lengths(gasr_asr *asr, const int *lens_host);
/* Per-token timesteps (see gasr_ctc_decode_ex) of later runs on / off; gasr_asr_timesteps copies those of the last run,
 * out_timesteps [N, nbest, max_len].                                                                               */
int gasr_asr_enable_timesteps(gasr_asr *asr, int on);
int gasr_asr_timesteps(gasr_asr *asr, int *out_timesteps);
/* Per-launch stage timing on/off (two event records per kernel launch; off by default) and the device time of the last
 * batch, first kernel to last result copy, measured with CUDA events.                                                 */
int gasr_asr_profile(gasr_asr *asr, int on);
int gasr_asr_last_ms(gasr_asr *asr, float *ms);

/* ---- jobs: many batches of utterances on one GPU, several in flight (BASELINE.json cfg5) ------------------------------- */
/*
 * A job owns `lanes` complete pipelines of cfg->N utterances each; batch b runs on lane b % lanes, so consecutive batches
 * overlap on the GPU.  x[b] is batch b in the reference's layout (time-major [T*N, in], RNN.cu:17), in host memory (pinned
 * for the copies to overlap) or device memory; results are written in batch order: out_paths [n_batches*N, nbest, max_len],
 * out_lens / out_scores [n_batches*N, nbest].  Configurations outside the wave engine are GASR_ERR_UNSUPPORTED.  The
 * multi-GPU layer gives every process / GPU a contiguous range of batches and gathers the results on the host
 * (utterances are independent: RNN.cu:15-27, CTCBeamSearch.cu:416 -- no collective on the data path).
 */
typedef struct gasr_job gasr_job;
int gasr_job_create(int device, const gasr_asr_config *cfg, const char *vocab, int lanes, gasr_job **job);
int gasr_job_destroy(gasr_job *job);
int gasr_job_set_weights(gasr_job *job, const float *const *w_ih, const float *const *w_hh, const float *const *b_ih,
                         const float *const *b_hh, const float *fc_w, const float *fc_b);
int gasr_job_run_host(gasr_job *job, const float *const *x_host, int n_batches, char *out_paths, int *out_lens,
                      float *out_scores);
int gasr_job_run_device(gasr_job *job, const float *const *x_dev, int n_batches, char *out_paths, int *out_lens,
                        float *out_scores);
int gasr_job_last_ms(gasr_job *job, float *ms);                 /* device time of the last run (CUDA events)            */
int gasr_job_launch_count(gasr_job *job, long long *launches);  /* kernels launched by all lanes since creation         */
int gasr_job_lane(gasr_job *job, int lane, gasr_ctx **ctx, gasr_asr **asr);   /* borrowed handles (allocation, log-probs) */
/* Per-launch stage timing for every lane; stage_times = sums over all batches of the last run (projection, recurrence,
 * linear + log-softmax, decode) and the launch counts.                                                              */
int gasr_job_profile(gasr_job *job, int on);
int gasr_job_stage_times(gasr_job *job, float *ms4, int *n4);

/* ---- synthetic inputs (counter-based RNG reproducible on host and device) --------------------------------- */
/* x_dev[T*N, D] time-major: element (t, d) of utterance first_utt + n = top 24 bits of
 * splitmix64((u*T*D + t*D + d) ^ splitmix64(seed)) / 2^24, like the torch.rand input of baseline/main.py:39.            */
int gasr_synth_spectrogram(gasr_ctx *ctx, float *x_dev, unsigned long long seed, int T, int N, int D, long long first_utt);

#ifdef __cplusplus
}
#endif
#endif /* GASR_H */
