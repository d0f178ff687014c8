// job.cu -- gasr_job_*: many batches ("waves") of utterances through the wave engine on ONE GPU, several waves in flight.
//
// BASELINE.json cfg5 decodes 8192 utterances; a GPU holds the activations of a few thousand at a time, so a job is a list
// of batches of the pipeline's N utterances each (reference layout per batch: time-major [T*N, in], RNN.cu:17).  A job owns
// `lanes` complete pipelines (gasr_asr + a private gasr_ctx, i.e. private streams and workspaces); batch b runs on lane
// b % lanes, so while one wave drains its decoder tail the next wave's GEMMs and recurrences already fill the SMs.  Batches
// are independent (no utterance ever interacts with another: RNN.cu:15-27, CTCBeamSearch.cu:416), results are written in
// batch order.  The multi-GPU layer above this (bench.py, shard.py) gives every rank a contiguous range of batches.
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "asr.cuh"
#include "common.cuh"

struct gasr_job {
    int device = 0, lanes = 0;
    gasr_asr_config cfg;
    std::vector<gasr_ctx *> ctx;
    std::vector<gasr_asr *> asr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.0f;
    float stage_ms[4] = {0, 0, 0, 0};          // sums over the batches of the last run (profiling on)
    int stage_launches[4] = {0, 0, 0, 0};
};

using namespace gasr;

extern "C" {

int gasr_job_destroy(gasr_job *job) {
    if (!job) return GASR_OK;
    for (gasr_asr *a : job->asr) if (a) gasr_asr_destroy(a);
    if (job->ev0) cudaEventDestroy(job->ev0);
    if (job->ev1) cudaEventDestroy(job->ev1);
    for (gasr_ctx *c : job->ctx) if (c) gasr_ctx_destroy(c);
    delete job;
    return GASR_OK;
}

int gasr_job_create(int device, const gasr_asr_config *cfg, const char *vocab, int lanes, gasr_job **out) {
    GASR_CHECK(cfg && vocab && out, "gasr_job_create: null argument");
    GASR_CHECK(lanes >= 1 && lanes <= 8, "gasr_job_create: lanes must be 1..8");
    *out = nullptr;
    gasr_job *job = new gasr_job();
    job->device = device; job->lanes = lanes; job->cfg = *cfg;
    job->ctx.assign(lanes, nullptr); job->asr.assign(lanes, nullptr);
    int st = GASR_OK;
    for (int i = 0; i < lanes && st == GASR_OK; i++) {
        st = gasr_ctx_create(device, &job->ctx[i]);
        if (st == GASR_OK) st = gasr_asr_create(job->ctx[i], cfg, vocab, &job->asr[i]);
        if (st == GASR_OK && job->asr[i]->wave == nullptr) {
            set_error("gasr_job_create: this configuration runs outside the wave engine (unidirectional tanh, H in {128,256,512}, beam and vocabulary <= 32)");
            st = GASR_ERR_UNSUPPORTED;
        }
    }
    if (st == GASR_OK) {
        cudaSetDevice(device);
        if (cudaEventCreate(&job->ev0) != cudaSuccess || cudaEventCreate(&job->ev1) != cudaSuccess) { set_error("gasr_job_create: event creation failed"); st = GASR_ERR_CUDA; }
    }
    if (st != GASR_OK) { gasr_job_destroy(job); return st; }
    *out = job;
    return GASR_OK;
}

int gasr_job_set_weights(gasr_job *job, const float *const *w_ih, const float *const *w_hh, const float *const *b_ih,
                         const float *const *b_hh, const float *fc_w, const float *fc_b) {
    GASR_CHECK(job != nullptr, "null gasr_job");
    for (gasr_asr *a : job->asr) GASR_TRY(gasr_asr_set_weights(a, w_ih, w_hh, b_ih, b_hh, fc_w, fc_b));
    return GASR_OK;
}

static int job_run(gasr_job *job, const float *const *x, bool host, int n_batches, char *out_paths, int *out_lens, float *out_scores) {
    GASR_CHECK(job != nullptr && x != nullptr && out_paths && out_lens && out_scores, "gasr_job_run: null argument");
    GASR_CHECK(n_batches >= 0, "gasr_job_run: negative batch count");
    if (n_batches == 0) { job->last_ms = 0.0f; return GASR_OK; }
    const gasr_asr_config &c = job->cfg;
    const size_t per = (size_t)c.N * c.nbest;
    std::vector<int> inflight(job->lanes, -1);
    for (int i = 0; i < 4; i++) { job->stage_ms[i] = 0.0f; job->stage_launches[i] = 0; }
    auto collect = [&](int lane) -> int {
        const int b = inflight[lane];
        inflight[lane] = -1;
        const int r = gasr_asr_collect(job->asr[lane], out_paths + (size_t)b * per * c.max_len, out_lens + (size_t)b * per, out_scores + (size_t)b * per);
        for (int i = 0; i < 4; i++) { job->stage_ms[i] += job->asr[lane]->stage_ms[i]; job->stage_launches[i] += job->asr[lane]->stage_launches[i]; }
        return r;
    };
    cudaSetDevice(job->device);
    GASR_CUDA(cudaEventRecord(job->ev0, job->ctx[0]->stream));
    int rc = GASR_OK;
    for (int b = 0; b < n_batches && rc == GASR_OK; b++) {
        const int lane = b % job->lanes;
        if (x[b] == nullptr) { set_error("gasr_job_run: null batch %d", b); rc = GASR_ERR_INVALID; break; }
        if (inflight[lane] >= 0) rc = collect(lane);
        if (rc == GASR_OK) rc = host ? gasr_asr_submit_host(job->asr[lane], x[b]) : gasr_asr_submit_device(job->asr[lane], x[b]);
        if (rc == GASR_OK) inflight[lane] = b;
    }
    // drain in submission order (also after an error: nothing may stay in flight)
    for (int k = 0; k < job->lanes; k++) {
        int oldest = -1;
        for (int lane = 0; lane < job->lanes; lane++)
            if (inflight[lane] >= 0 && (oldest < 0 || inflight[lane] < inflight[oldest])) oldest = lane;
        if (oldest < 0) break;
        const int r2 = collect(oldest);
        if (rc == GASR_OK) rc = r2;
    }
    cudaSetDevice(job->device);
    if (rc == GASR_OK) {
        GASR_CUDA(cudaEventRecord(job->ev1, job->ctx[0]->stream));
        GASR_CUDA(cudaEventSynchronize(job->ev1));
        GASR_CUDA(cudaEventElapsedTime(&job->last_ms, job->ev0, job->ev1));
    }
    return rc;
}

int gasr_job_run_host(gasr_job *job, const float *const *x_host, int n_batches, char *out_paths, int *out_lens, float *out_scores) {
    return job_run(job, x_host, true, n_batches, out_paths, out_lens, out_scores);
}

int gasr_job_run_device(gasr_job *job, const float *const *x_dev, int n_batches, char *out_paths, int *out_lens, float *out_scores) {
    return job_run(job, x_dev, false, n_batches, out_paths, out_lens, out_scores);
}

int gasr_job_last_ms(gasr_job *job, float *ms) {
    GASR_CHECK(job && ms, "gasr_job_last_ms: null argument");
    *ms = job->last_ms;
    return GASR_OK;
}

int gasr_job_launch_count(gasr_job *job, long long *launches) {
    GASR_CHECK(job && launches, "gasr_job_launch_count: null argument");
    *launches = 0;
    for (gasr_ctx *c : job->ctx) *launches += c->launches;
    return GASR_OK;
}

int gasr_job_profile(gasr_job *job, int on) {
    GASR_CHECK(job != nullptr, "null gasr_job");
    for (gasr_asr *a : job->asr) a->profile = on != 0;
    return GASR_OK;
}

int gasr_job_stage_times(gasr_job *job, float *ms4, int *n4) {
    GASR_CHECK(job && ms4 && n4, "gasr_job_stage_times: null argument");
    for (int i = 0; i < 4; i++) { ms4[i] = job->stage_ms[i]; n4[i] = job->stage_launches[i]; }
    return GASR_OK;
}

int gasr_job_lane(gasr_job *job, int lane, gasr_ctx **ctx, gasr_asr **asr) {
    GASR_CHECK(job && lane >= 0 && lane < job->lanes, "gasr_job_lane: bad lane");
    if (ctx) *ctx = job->ctx[lane];
    if (asr) *asr = job->asr[lane];
    return GASR_OK;
}

/* ---- test-pattern generator ------------------------------------------------------------------------------------------ */
// synthetic spectrograms come from a counter-based RNG reproducible on host and device -- value of element
// (t, d) of utterance u = top 24 bits of splitmix64((u*T*D + t*D + d) ^ splitmix64(seed)) / 2^24, in [0, 1)
// (synth.py:uniform01 / spectrogram_batch is the host twin; tests compare the two bit for bit).
__device__ __forceinline__ unsigned long long job_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void synth_spectrogram_kernel(float *__restrict__ x, unsigned long long seed_mix, int T, int N, int D, long long first_utt) {
    const size_t total = (size_t)T * N * D;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / D;
        const int d = (int)(i - row * D);
        const int t = (int)(row / N), n = (int)(row - (size_t)t * N);
        const unsigned long long idx = (unsigned long long)(first_utt + n) * (unsigned long long)T * D + (unsigned long long)t * D + d;
        const unsigned long long r = job_splitmix64(idx ^ seed_mix);
        x[i] = (float)(unsigned)(r >> 40) / 16777216.0f;
    }
}

int gasr_synth_spectrogram(gasr_ctx *ctx, float *x_dev, unsigned long long seed, int T, int N, int D, long long first_utt) {
    GASR_CHECK(ctx && x_dev && T >= 0 && N >= 0 && D >= 0 && first_utt >= 0, "gasr_synth_spectrogram: bad arguments");
    cudaSetDevice(ctx->device);
    const size_t total = (size_t)T * N * D;
    if (total == 0) return GASR_OK;
    // host twin of the seed mixing: splitmix64(seed)
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull, z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    synth_spectrogram_kernel<<<blocks, 256, 0, ctx->stream>>>(x_dev, z, T, N, D, first_utt);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

}  // extern "C"
