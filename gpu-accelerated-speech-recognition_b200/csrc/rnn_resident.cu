// rnn_resident.cu -- tanh recurrence for WIDE hidden layers and a HANDFUL of utterances (BASELINE.json cfg1:
// baseline/config.json, one utterance, H = 2048) as one persistent fp32 kernel with W_hh resident in shared memory.
//
// Stands behind the time loop of RNN::forward (reference RNN.cu:9-30) for shapes outside the tensor-core kernels (H not in
// {64 .. 512}).  The per-timestep fallback read all of W_hh (16.8 MB at H = 2048) through L2 every step: 95 us per step for
// ONE utterance.  Here the whole GPU holds W_hh once: CTA c keeps its slice of columns ([H x cols] fp32, 112 KB at H = 2048
// over 147 CTAs) in shared memory for all T steps; per step it reads h_{t-1} (N x H floats, L2), accumulates its columns
// (thread = slice of K, conflict-free shared-memory reads, warp + CTA reduction), applies tanh and writes its columns of h_t;
// the CTAs then meet at a grid barrier (release: __threadfence + atomicAdd, acquire: ld.acquire.gpu; every CTA is resident,
// checked by the launcher).  Exact fp32 arithmetic (no operand split), tanhf.
// Every wait is bounded (4 s -> __trap): a CTA that never became resident is a launch failure, not a hang.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace gasr {

constexpr int RR_THREADS = 256;
constexpr int RR_MAXN = 4;                       // utterances (accumulator rows per thread)
constexpr int RR_MAXC = 16;                      // columns per CTA
constexpr unsigned long long RR_TIMEOUT_NS = 4000000000ull;

struct RnnResidentParams {
    int T, N, H, cols, reverse;
    const float *xp; int ldxp;                   // x * W_ih + bias, [T * N, >= H]
    const float *w_hh;                           // [H, H] row-major (reference layout [in, out])
    float *out; int ldo;                         // h, [T * N, ldo], already offset to this direction's columns
    unsigned *cnt;                               // grid-barrier counter, zeroed before the launch
};

__device__ __forceinline__ unsigned rr_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long rr_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(RR_THREADS, 1) rnn_resident_kernel(const RnnResidentParams p) {
    extern __shared__ __align__(16) float rr_smem[];
    const int H = p.H, N = p.N, cols = p.cols;
    float *w_s = rr_smem;                        // [cols][H]: k fastest, so consecutive threads read consecutive words
    float *h_s = w_s + (size_t)cols * H;         // [N][H]
    float *red = h_s + (size_t)N * H;            // [8 warps][RR_MAXN][RR_MAXC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * cols;
    const int nc = min(cols, H - c0);            // columns of this CTA (the last CTA may have fewer)

    for (int i = tid; i < cols * H; i += RR_THREADS) {
        const int c = i / H, k = i - c * H;
        w_s[i] = c < nc ? p.w_hh[(size_t)k * H + c0 + c] : 0.0f;
    }
    __syncthreads();

    for (int s = 0; s < p.T; s++) {
        const int t = p.reverse ? p.T - 1 - s : s;
        float acc[RR_MAXN][RR_MAXC];
#pragma unroll
        for (int n = 0; n < RR_MAXN; n++)
#pragma unroll
            for (int c = 0; c < RR_MAXC; c++) acc[n][c] = 0.0f;
        if (s > 0) {
            // every CTA has stored its columns of h_{s-1}
            if (tid == 0) {
                const unsigned need = (unsigned)s * gridDim.x;
                if (rr_ld_acquire(p.cnt) < need) {
                    const unsigned long long t0 = rr_now_ns();
                    unsigned spins = 0;
                    while (rr_ld_acquire(p.cnt) < need)
                        if ((++spins & 1023u) == 0 && rr_now_ns() - t0 > RR_TIMEOUT_NS) __trap();
                }
            }
            __syncthreads();
            const int tp = p.reverse ? t + 1 : t - 1;
            for (int i = tid; i < N * H; i += RR_THREADS) {
                const int n = i / H, k = i - n * H;
                h_s[i] = __ldcg(p.out + ((size_t)tp * N + n) * p.ldo + k);
            }
            __syncthreads();
            for (int k = tid; k < H; k += RR_THREADS) {
                float hk[RR_MAXN];
#pragma unroll
                for (int n = 0; n < RR_MAXN; n++) hk[n] = n < N ? h_s[n * H + k] : 0.0f;
#pragma unroll
                for (int c = 0; c < RR_MAXC; c++) {
                    if (c >= cols) break;
                    const float wv = w_s[c * H + k];
#pragma unroll
                    for (int n = 0; n < RR_MAXN; n++) acc[n][c] = fmaf(hk[n], wv, acc[n][c]);
                }
            }
            // warp reduction, then across the eight warps through shared memory
#pragma unroll
            for (int n = 0; n < RR_MAXN; n++)
#pragma unroll
                for (int c = 0; c < RR_MAXC; c++) {
                    if (n >= N || c >= cols) continue;
                    float v = acc[n][c];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if (lane == 0) red[(warp * RR_MAXN + n) * RR_MAXC + c] = v;
                }
            __syncthreads();
        }
        if (tid < N * nc) {
            const int n = tid / nc, c = tid - n * nc;
            float sum = 0.0f;
            if (s > 0)
                for (int w = 0; w < RR_THREADS / 32; w++) sum += red[(w * RR_MAXN + n) * RR_MAXC + c];
            const size_t row = (size_t)t * N + n;
            p.out[row * p.ldo + c0 + c] = tanhf(p.xp[row * p.ldxp + c0 + c] + sum);
        }
        if (s + 1 < p.T) {
            __syncthreads();                     // this CTA's columns of h_s are stored (and `red` may be reused)
            if (tid == 0) { __threadfence(); atomicAdd(p.cnt, 1u); }
        }
    }
}

// Shapes: a few utterances, a hidden layer too wide for the tensor-core kernels; W_hh must fit in the shared memory of the
// CTAs that can be resident at once (`share` = 2 when another recurrence runs concurrently: the other direction of a
// bidirectional layer).
static bool rr_shape(const gasr_ctx *ctx, int N, int H, int share, int &cols, int &ctas, size_t &smem) {
    if (N < 1 || N > RR_MAXN || H < 1024) return false;
    const int max_ctas = ctx->sm_count / (share > 1 ? share : 1);
    cols = ceil_div(H, max_ctas);
    if (cols > RR_MAXC) return false;
    ctas = ceil_div(H, cols);
    smem = sizeof(float) * ((size_t)cols * H + (size_t)N * H + (RR_THREADS / 32) * RR_MAXN * RR_MAXC);
    return smem <= (size_t)ctx->max_smem_optin && N * cols <= RR_THREADS;
}

bool rnn_resident_supported(const gasr_ctx *ctx, const RnnLayerArgs &a) {
    int cols, ctas; size_t smem;
    return a.cell == GASR_CELL_TANH && a.s0 == 0 && (a.s1 == 0 || a.s1 == a.T) && a.T >= 2 &&
           rr_shape(ctx, a.N, a.H, a.concurrent ? 2 : 1, cols, ctas, smem);
}

// ws: at least 256 bytes of device memory (the barrier counter)
int launch_rnn_resident(gasr_ctx *ctx, const RnnLayerArgs &a, void *ws, cudaStream_t st) {
    int cols = 0, ctas = 0; size_t smem = 0;
    GASR_CHECK(rr_shape(ctx, a.N, a.H, a.concurrent ? 2 : 1, cols, ctas, smem), "rnn_resident: unsupported shape N=%d H=%d", a.N, a.H);
    RnnResidentParams p;
    p.T = a.T; p.N = a.N; p.H = a.H; p.cols = cols; p.reverse = a.reverse;
    p.xp = a.xproj; p.ldxp = a.ldxp; p.w_hh = a.w_hh; p.out = a.out + a.col0; p.ldo = a.ldo;
    p.cnt = static_cast<unsigned *>(ws);
    GASR_CUDA(cudaMemsetAsync(ws, 0, 256, st));
    if (!(ctx->attr_mask & 65536u)) {
        GASR_CUDA(cudaFuncSetAttribute(rnn_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
        ctx->attr_mask |= 65536u;
    }
    rnn_resident_kernel<<<ctas, RR_THREADS, smem, st>>>(p);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

}  // namespace gasr
