// asr.cuh -- the fused pipeline object behind gasr_asr_* (capi.cu: sequential / time-chunked / streaming modes; asr_wave.cu: the
// throughput engine).
#pragma once
#include <vector>

#include "common.cuh"
#include "stream.cuh"

namespace gasr {

// Optional per-kernel timing: events recorded between launches on the same stream (no host sync).
struct StageEvents {
    std::vector<cudaEvent_t> pool;
    std::vector<int> tag;        // tag[i] = stage that ends at event i (0 proj, 1 recurrence, 2 linear, 3 decode, -1 start)
    size_t used = 0;
    int mark(int stage, cudaStream_t st) {
        if (used == pool.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return GASR_ERR_CUDA;
            pool.push_back(e); tag.push_back(-1);
        }
        tag[used] = stage;
        return cudaEventRecord(pool[used++], st) == cudaSuccess ? GASR_OK : GASR_ERR_CUDA;
    }
};

struct WaveState;                               // asr_wave.cu

// the calling thread's current device is switched to the context's for the duration of an entry point
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// stacked recurrent forward on a given stream (capi.cu; gasr_rnn_forward and the fused pipeline's sequential mode)
int rnn_forward_impl(gasr_ctx *ctx, int cell, int bidir, int T, int N, int in, int H, int L, const float *const *w_ih,
                     const float *const *w_hh, const float *const *b_ih, const float *const *b_hh, const float *x,
                     float *const *hiddens, int precision, cudaStream_t st, StageEvents *prof = nullptr);

}  // namespace gasr

struct gasr_asr {
    gasr_ctx *ctx = nullptr;
    gasr_asr_config cfg;
    std::vector<char> vocab;
    int D = 1, G = 1, ldp = 0;
    std::vector<float *> w_ih, w_hh, b_ih, b_hh, hiddens;
    float *fc_w = nullptr, *fc_b = nullptr, *x_dev = nullptr, *logp = nullptr;
    bool have_weights = false;
    gasr::StageEvents prof;
    float stage_ms[4] = {0, 0, 0, 0};
    int stage_launches[4] = {0, 0, 0, 0};
    // pipelined execution: time chunks flow through (layer 0 .. L-1, linear + decode) on separate streams
    int chunk = 0;                              // frames per chunk (0 = sequential path)
    float *xproj_all = nullptr, *bias_all = nullptr;   // [L][T*N*H], [L][H]
    std::vector<void *> tc_abuf, tc_wbuf;       // per layer: bf16 hi/lo planes of the layer input / of W_ih^T
    bool use_tc = false;
    // streaming execution (stream_*): persistent kernels coupled by progress counters, no kernel boundaries in time
    bool stream_ok = false;
    int stream_fpb = 0, stream_blocks = 0, stream_gemm_ctas = 0;
    void *x_planes = nullptr;                   // bf16 hi/lo planes of the input batch [rows, Kp]
    std::vector<void *> h_planes;               // per layer: bf16 hi/lo planes of the hidden sequence [rows, H] x 2
    void *fc_wbuf = nullptr;                    // W_fc^T hi/lo planes, padded to 32 output rows
    float *fc_b_pad = nullptr;
    unsigned *flags = nullptr;                  // [h_done L][xp_ready L][lp_ready][x_ready][misc 16] x blocks
    size_t flags_bytes = 0;
    int *host_words = nullptr, *host_words_dev = nullptr;   // mapped host memory: [0] go, [1] error
    int epoch = 0;
    gasr::XsMaps xs_maps;
    cudaEvent_t ev_cp = nullptr, ev_go = nullptr, ev_r0 = nullptr, ev_r1 = nullptr, ev_g0 = nullptr, ev_g1 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
    std::vector<cudaEvent_t> sync_ev;           // cross-stream dependencies (no timing)
    std::vector<cudaEvent_t> t0_ev, t1_ev;      // per-launch timing pairs
    std::vector<int> t_tag;
    size_t n_timed = 0;
    // wave engine (asr_wave.cu): throughput mode, stream-ordered time chunks over groups of 128 utterances
    gasr::WaveState *wave = nullptr;
    bool profile = false;                       // per-launch stage timing (adds two event records per launch)
    // decoder options of baseline/main.py:45-46 (decoder.decode(output, out_lens) -> ..., timesteps, out_seq_len)
    int *lens_dev = nullptr;                    // frames per utterance [N] (null: every utterance has T frames)
    bool want_ts = false;                       // per-token timesteps
    std::vector<int> ts_host;                   // [N, nbest, max_len] of the last run
    void decode_extras(gasr::CtcArgs &ca) { ca.lens_dev = lens_dev; ca.out_timesteps = want_ts ? ts_host.data() : nullptr; }
};

namespace gasr {

// throughput engine (asr_wave.cu)
bool wave_supported(const gasr_ctx *ctx, const gasr_asr_config &c);
int wave_create(gasr_asr *a);                   // buffers, streams, events, TMA descriptors (device-synchronising)
void wave_destroy(gasr_asr *a);
int wave_set_weights(gasr_asr *a, const float *fc_w_host, const float *fc_b_host);   // after the fp32 weights are resident
int wave_submit(gasr_asr *a, const float *x_dev, const float *x_host);   // enqueue one batch; returns without waiting
int wave_collect(gasr_asr *a, char *out_paths, int *out_lens, float *out_scores);   // wait + unpack the results
int wave_logprobs(gasr_asr *a, const float **logp_dev, int *ldp);
int wave_refresh_decoder(gasr_asr *a);           // decoder workspaces after lengths / timesteps were switched on
bool wave_pending(const gasr_asr *a);            // a submitted batch has not been collected yet
int wave_chunk_frames(const gasr_asr *a);
float wave_last_ms(const gasr_asr *a);

}  // namespace gasr

#define GASR_ENTER(ctx)                                              \
    if ((ctx) == nullptr) { gasr::set_error("null gasr_ctx"); return GASR_ERR_INVALID; } \
    gasr::DeviceGuard guard__((ctx)->device)
