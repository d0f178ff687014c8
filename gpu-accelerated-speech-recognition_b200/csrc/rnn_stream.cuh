// rnn_stream.cuh -- parameters of the persistent multi-layer recurrence kernel (rnn_stream.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace gasr {

constexpr int RS_MAX_LAYERS = 4;

struct RnnStreamLayer {
    const float *xproj; int ldxp;     // [T*N, ldxp] x*W_ih + biases of this layer
    const float *w_hh;                // [H, H] reference layout [in, out]
    float *out; int ldo;              // fp32 hidden sequence [T*N, ldo] (may be null when only the planes are needed)
    __nv_bfloat16 *out_hi, *out_lo;   // bf16 hi/lo planes [T*N, ldp]: the next layer's projection GEMM operand (may be null)
    int ldp;
    const unsigned *xp_ready;         // [blocks] projection progress: block b is complete when xp_ready[b] >= xp_need (null: all ready)
    unsigned *h_done;                 // [blocks] += 1 per CTA once the block's outputs are visible (null: no consumer)
};

struct RnnStreamParams {
    int T, N, L, groups;              // groups = ceil(N / (8 * sub-batches per cluster))
    int nsub;                         // 0: first-generation kernel (2 sub-batches); 2 or 4: software-pipelined kernel
    int frames_per_block;             // frames covered by one progress counter
    int xp_need;                      // arrivals that complete an xproj block
    int *error;                       // set to 1 if a watchdog fired
    volatile unsigned *abort;         // device word shared by the pipeline's kernels: stop waiting (null: none)
    unsigned *started;                // device counter of CTAs that are running (null: no handshake)
    volatile int *host_go; int epoch; // mapped host word: set to epoch once every CTA of the grid is resident
    RnnStreamLayer layer[RS_MAX_LAYERS];
};

bool rnn_stream_supported(const gasr_ctx *ctx, int H, int N, int L);
int rnn_stream_default_nsub(int N);
int rnn_stream_max_clusters(gasr_ctx *ctx, int H, int *clusters, int *ctas_per_cluster);   // device-wide residency of the recurrence kernel
int launch_rnn_stream(gasr_ctx *ctx, const RnnStreamParams &p, int H, cudaStream_t st);

}  // namespace gasr
