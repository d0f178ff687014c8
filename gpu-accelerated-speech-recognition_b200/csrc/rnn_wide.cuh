// rnn_wide.cuh -- throughput-mode tanh recurrence (rnn_wide.cu): groups of 128 utterances on tcgen05, W_hh resident in shared memory.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace gasr {

// One layer's fixed operands: the bf16 hi/lo planes of its hidden sequence ([T * Npad, H], row t * Npad + n; Npad = N rounded
// up to whole groups of 128) and the prepared W_hh^T planes, with their TMA descriptors.
struct RnnWidePlan {
    CUtensorMap maps[4];           // h hi, h lo, W^T hi, W^T lo
    __nv_bfloat16 *hi, *lo;
    int T, N, Npad, H;
};

struct RnnWideRun {
    int s0, s1;                    // steps [s0, s1): h_{s0-1} is read back from the planes (h_{-1} = 0)
    const float *xp; int ldxp;     // x*W_ih + (b_ih + b_hh), row t * xp_rows_per_frame + n
    int xp_rows_per_frame;
    float *out; int ldo, col0;     // optional fp32 h_t, row t * out_rows_per_frame + n (null: planes only)
    int out_rows_per_frame;
    int groups_per_cluster;        // 1 or 2
    int multicast;                 // TMA multicast of the h boxes across the cluster
};

bool rnn_wide_supported(const gasr_ctx *ctx, int H);
size_t rnn_wide_smem_bytes(int H);
size_t rnn_wide_plane_bytes(int T, int Npad, int H);       // bytes of ONE plane; the planes buffer holds hi then lo
int rnn_wide_prepare(gasr_ctx *ctx, int H);                // function attributes (call before concurrent kernels run)
int rnn_wide_plan(gasr_ctx *ctx, RnnWidePlan &pl, const float *w_hh, int T, int N, int H, void *wbuf, void *planes,
                  cudaStream_t st, int pad_to = 128);
int launch_rnn_wide(gasr_ctx *ctx, const RnnWidePlan &pl, const RnnWideRun &r, cudaStream_t st);
// CTA-pair variant (rnn_wide2.cu, tcgen05.mma.cta_group::2): groups of 256 utterances, plan padded to 256
bool rnn_wide2_supported(const gasr_ctx *ctx, int H);
int rnn_wide2_prepare(gasr_ctx *ctx);
int launch_rnn_wide2(gasr_ctx *ctx, const RnnWidePlan &pl, const RnnWideRun &r, cudaStream_t st);

}  // namespace gasr
