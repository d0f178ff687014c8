// linear.cu -- Linear (+bias) fused with its activation: ReLU (the reference's Linear, Linear.cu:3-10,42-49),
// none, or log-softmax over the vocabulary (baseline/model.py:49; absent from the reference's C++).
//
// Fast path (out <= 32, W fits shared memory): persistent CTAs keep W[in, 32] resident in shared memory; one
// warp owns 8 rows at a time, lane j owns output column j, x is read with broadcast 128-bit loads, the
// row-wise max / sum-exp of the log-softmax are warp shuffles.  One pass over x, one write of y.
// General path: FFMA GEMM with fused bias (gemm_simt.cu) + a row-wise activation kernel.
#include <math.h>

#include "common.cuh"

namespace gasr {

constexpr int LIN_ROWS = 8;   // rows per warp per iteration

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int ACT>
__global__ void __launch_bounds__(256) linear_small_out_kernel(const float *__restrict__ x, int ldx,
                                                               const float *__restrict__ W, const float *__restrict__ b,
                                                               float *__restrict__ y, int ldy, int rows, int in, int out) {
    extern __shared__ __align__(16) float Ws[];   // [in][32], columns >= out are zero
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < in * 32; i += blockDim.x) {
        const int k = i >> 5, c = i & 31;
        Ws[i] = c < out ? W[(size_t)k * out + c] : 0.0f;
    }
    __syncthreads();
    const float bias = (b != nullptr && lane < out) ? b[lane] : 0.0f;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
    const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int r0 = gw * LIN_ROWS; r0 < rows; r0 += warps_total * LIN_ROWS) {
        float acc[LIN_ROWS];
#pragma unroll
        for (int r = 0; r < LIN_ROWS; r++) acc[r] = 0.0f;
        const int nr = min(LIN_ROWS, rows - r0);
        int k = 0;
        if (vec) {
            for (; k + 4 <= in; k += 4) {
                const float w0 = Ws[(k + 0) * 32 + lane], w1 = Ws[(k + 1) * 32 + lane];
                const float w2 = Ws[(k + 2) * 32 + lane], w3 = Ws[(k + 3) * 32 + lane];
#pragma unroll
                for (int r = 0; r < LIN_ROWS; r++) {
                    if (r < nr) {
                        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + (size_t)(r0 + r) * ldx + k));
                        acc[r] = fmaf(xv.x, w0, acc[r]);
                        acc[r] = fmaf(xv.y, w1, acc[r]);
                        acc[r] = fmaf(xv.z, w2, acc[r]);
                        acc[r] = fmaf(xv.w, w3, acc[r]);
                    }
                }
            }
        }
        for (; k < in; k++) {
            const float w0 = Ws[k * 32 + lane];
#pragma unroll
            for (int r = 0; r < LIN_ROWS; r++)
                if (r < nr) acc[r] = fmaf(__ldg(x + (size_t)(r0 + r) * ldx + k), w0, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < LIN_ROWS; r++) {
            if (r >= nr) break;
            float v = acc[r] + bias;
            if (ACT == GASR_ACT_RELU) v = v < 0.0f ? 0.0f : v;
            if (ACT == GASR_ACT_LOGSOFTMAX) {
                const float mx = warp_max(lane < out ? v : -INFINITY);
                const float e = lane < out ? expf(v - mx) : 0.0f;
                const float lse = logf(warp_sum(e));
                v = (v - mx) - lse;
            }
            if (lane < out) y[(size_t)(r0 + r) * ldy + lane] = v;
        }
    }
}

// row-wise activation for the general path: one warp per row, any number of columns
template <int ACT>
__global__ void __launch_bounds__(256) row_act_kernel(const float *x, int ldx, float *y,
                                                      int ldy, int rows, int cols) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int r = gw; r < rows; r += warps_total) {
        const float *xr = x + (size_t)r * ldx;
        float *yr = y + (size_t)r * ldy;
        if (ACT == GASR_ACT_RELU) {
            for (int c = lane; c < cols; c += 32) { const float v = xr[c]; yr[c] = v < 0.0f ? 0.0f : v; }
        } else {
            float mx = -INFINITY;
            for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c]);
            mx = warp_max(mx);
            float s = 0.0f;
            for (int c = lane; c < cols; c += 32) s += expf(xr[c] - mx);
            const float lse = logf(warp_sum(s));
            for (int c = lane; c < cols; c += 32) yr[c] = (xr[c] - mx) - lse;
        }
    }
}

int launch_log_softmax(gasr_ctx *ctx, const float *x, int ldx, float *y, int ldy, int rows, int cols, cudaStream_t st) {
    GASR_CHECK(x && y && rows >= 0 && cols >= 1 && ldx >= cols && ldy >= cols, "log_softmax: bad arguments");
    if (rows == 0) return GASR_OK;
    int blocks = ceil_div(rows, 8);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    row_act_kernel<GASR_ACT_LOGSOFTMAX><<<blocks, 256, 0, st>>>(x, ldx, y, ldy, rows, cols);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

int launch_linear(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *b, float *y, int ldy, int rows,
                  int in, int out, int act, cudaStream_t st) {
    GASR_CHECK(x && W && y, "linear: null operand");
    GASR_CHECK(rows >= 0 && in >= 1 && out >= 1 && ldx >= in && ldy >= out, "linear: bad shape");
    GASR_CHECK(act == GASR_ACT_NONE || act == GASR_ACT_RELU || act == GASR_ACT_LOGSOFTMAX, "linear: unknown activation");
    if (rows == 0) return GASR_OK;
    if (!ctx->opt.linear_simt && linear_tc_supported(rows, in, out, ldy, y, act))
        return launch_linear_logsoftmax_tc(ctx, x, ldx, W, b, y, ldy, rows, in, out, st);
    // Wide layers (the DeepSpeech FC front / back layers, main.cpp:31-45, baseline/model.py:22-35): the tcgen05 tile engine
    // with bias + ReLU in its epilogue -- fp32-grade (3-term bf16 split), one GEMM launch instead of FFMA GEMM + activation.
    if (!ctx->opt.linear_simt && act != GASR_ACT_LOGSOFTMAX && rows >= 32 && xproj_tc_supported(rows, in, out) && ldy % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        const size_t wb = xproj_tc_w_bytes(in, out), ab = xproj_tc_a_bytes(rows, in);
        GASR_TRY(ws_reserve(ctx, ctx->ws_lin, wb + ab + 2048));
        unsigned char *base = static_cast<unsigned char *>(ctx->ws_lin.ptr);
        GASR_TRY(xproj_tc_prepare_weights(ctx, W, in, out, base, st));
        return launch_xproj_tc(ctx, x, ldx, rows, in, out, base, base + align_up(wb, 1024), b, y, ldy, GASR_PREC_FP32, st,
                               act == GASR_ACT_RELU ? 1 : 0);
    }
    const size_t smem = (size_t)in * 32 * sizeof(float);
    if (out <= 32 && smem <= (size_t)ctx->max_smem_optin - 1024) {
        const int tiles = ceil_div(rows, LIN_ROWS * 8);
        int blocks = tiles < ctx->sm_count ? tiles : ctx->sm_count;
        if (smem <= 100 * 1024 && tiles > ctx->sm_count) blocks = tiles < 2 * ctx->sm_count ? tiles : 2 * ctx->sm_count;
#define GASR_LIN_LAUNCH(A)                                                                                         \
    GASR_CUDA(cudaFuncSetAttribute(linear_small_out_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    linear_small_out_kernel<A><<<blocks, 256, smem, st>>>(x, ldx, W, b, y, ldy, rows, in, out)
        if (act == GASR_ACT_NONE) { GASR_LIN_LAUNCH(GASR_ACT_NONE); }
        else if (act == GASR_ACT_RELU) { GASR_LIN_LAUNCH(GASR_ACT_RELU); }
        else { GASR_LIN_LAUNCH(GASR_ACT_LOGSOFTMAX); }
#undef GASR_LIN_LAUNCH
        GASR_CUDA(cudaGetLastError());
        ctx->launches += 1;
        return GASR_OK;
    }
    GASR_TRY(launch_matmul(ctx, x, ldx, 0, W, out, 0, y, ldy, rows, in, out, b, st));
    if (act != GASR_ACT_NONE) {
        int blocks = ceil_div(rows, 8);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        if (act == GASR_ACT_RELU) row_act_kernel<GASR_ACT_RELU><<<blocks, 256, 0, st>>>(y, ldy, y, ldy, rows, out);
        else row_act_kernel<GASR_ACT_LOGSOFTMAX><<<blocks, 256, 0, st>>>(y, ldy, y, ldy, rows, out);
        GASR_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    return GASR_OK;
}

}  // namespace gasr
