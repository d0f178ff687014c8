// rnn.cu -- recurrent part of RNN_Cell / RNN (reference RNN_Cell.cu:5-13,65-74, RNN.cu:9-30) and the GRU of cfg3.
//
// The reference runs, per (timestep, layer), 2 cublasSgemm + 1 cublasSgeam + 1 bias/tanh kernel with a host sync
// after each -- 12 000 device round trips at cfg2.  Here a layer is split into
//   (1) x*W_ih (+ both biases) for ALL timesteps at once -- a dense GEMM (gemm_simt.cu / xproj_gemm_tc.cu), and
//   (2) the recurrence h_t = tanh(xproj_t + h_{t-1}*W_hh): ONE persistent kernel for all T steps.
//
// rnn_tanh_cluster_kernel: one thread-block cluster (CS = H/64 CTAs, <= 8) per group of 16 utterances.  CTA c
// keeps the 64-column slice W_hh[:, 64c:64c+64] resident in shared memory for the whole sequence (128 KB at
// H = 512) plus a double-buffered copy of the group's full h (16 x H fp32).  Per step each CTA computes its
// [16 x 64] slice (8-way split-K over 16 warps, 4x4 register tiles, operands read as 128-bit shared loads),
// reduces the partials, adds xproj (prefetched one step ahead), applies tanh, stores h_t to HBM and broadcasts
// its slice into every CTA's shared memory through DSMEM; one cluster barrier per step.  Per step a CTA touches
// HBM only for its xproj slice and its output slice.
//
// Shapes outside the cluster kernel's envelope (H not 64 * 2^j <= 512, GRU) use one small kernel per timestep
// with W_hh streamed from L2.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "rnn_wide.cuh"
#include "rnn_stream.cuh"

namespace cg = cooperative_groups;

namespace gasr {

// ---- single cell step (RNN_Cell::forward) ------------------------------------------------------------------
__global__ void rnn_cell_kernel(const float *__restrict__ x, const float *__restrict__ h, const float *__restrict__ w_ih,
                                const float *__restrict__ w_hh, const float *__restrict__ b_ih,
                                const float *__restrict__ b_hh, float *__restrict__ out, int batch, int in, int hidden) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (j >= hidden || n >= batch) return;
    float ih = 0.0f, hh = 0.0f;
    for (int k = 0; k < in; k++) ih = fmaf(x[(size_t)n * in + k], w_ih[(size_t)k * hidden + j], ih);
    for (int k = 0; k < hidden; k++) hh = fmaf(h[(size_t)n * hidden + k], w_hh[(size_t)k * hidden + j], hh);
    // matrixAdd(ih, hh) then data += (bias1 + bias2), tanh  (RNN_Cell.cu:68, :10-12)
    out[(size_t)n * hidden + j] = tanhf((ih + hh) + (b_hh[j] + b_ih[j]));
}

int launch_rnn_cell(gasr_ctx *ctx, const float *x, const float *h_prev, const float *w_ih, const float *w_hh,
                    const float *b_ih, const float *b_hh, float *out, int batch, int in, int hidden, cudaStream_t st) {
    GASR_CHECK(x && h_prev && w_ih && w_hh && b_ih && b_hh && out, "rnn_cell: null operand");
    GASR_CHECK(batch >= 0 && in >= 1 && hidden >= 1 && batch <= 65535, "rnn_cell: bad shape");
    if (batch == 0) return GASR_OK;
    dim3 grid(ceil_div(hidden, 128), batch);
    rnn_cell_kernel<<<grid, 128, 0, st>>>(x, h_prev, w_ih, w_hh, b_ih, b_hh, out, batch, in, hidden);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

// ---- per-timestep fallback kernels ---------------------------------------------------------------------------
// out[n, j] = tanh(xp[n, j] + sum_k hprev[n, k] * W[k, j])
__global__ void rnn_tanh_step_kernel(const float *__restrict__ xp, int ldxp, const float *__restrict__ hprev, int ldh,
                                     const float *__restrict__ W, float *__restrict__ out, int ldo, int N, int H) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (j >= H) return;
    float acc = 0.0f;
    if (hprev != nullptr)
        for (int k = 0; k < H; k++) acc = fmaf(hprev[(size_t)n * ldh + k], W[(size_t)k * H + j], acc);
    out[(size_t)n * ldo + j] = tanhf(xp[(size_t)n * ldxp + j] + acc);
}

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// torch.nn.GRU step, gate order r,z,n; xp already holds x*W_ih + b_ih; b_hh stays on the hidden side.
__global__ void gru_step_kernel(const float *__restrict__ xp, int ldxp, const float *__restrict__ hprev, int ldh,
                                const float *__restrict__ W, const float *__restrict__ b_hh, float *__restrict__ out,
                                int ldo, int N, int H) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (j >= H) return;
    float gr = 0.0f, gz = 0.0f, gn = 0.0f, hj = 0.0f;
    if (hprev != nullptr) {
        const float *hp = hprev + (size_t)n * ldh;
        for (int k = 0; k < H; k++) {
            const float hv = hp[k];
            const float *w = W + (size_t)k * 3 * H;
            gr = fmaf(hv, w[j], gr);
            gz = fmaf(hv, w[H + j], gz);
            gn = fmaf(hv, w[2 * H + j], gn);
        }
        hj = hp[j];
    }
    const float *xr = xp + (size_t)n * ldxp;
    const float r = sigmoid_f(xr[j] + (gr + b_hh[j]));
    const float z = sigmoid_f(xr[H + j] + (gz + b_hh[H + j]));
    const float nn = tanhf(xr[2 * H + j] + r * (gn + b_hh[2 * H + j]));
    out[(size_t)n * ldo + j] = (1.0f - z) * nn + z * hj;
}

// GRU gates once the hidden-side products hh = h_{t-1} * W_hh + b_hh are known (one GEMM per timestep): torch.nn.GRU,
// gate order r, z, n; xp already holds x * W_ih + b_ih.  hh == nullptr: first step, h_0 = 0, so hh is just b_hh.
__global__ void gru_gate_kernel(const float *__restrict__ xp, int ldxp, const float *__restrict__ hh, const float *__restrict__ b_hh,
                                const float *__restrict__ hprev, int ldh, float *__restrict__ out, int ldo, int N, int H) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (j >= H) return;
    const float *xr = xp + (size_t)n * ldxp;
    float gr, gz, gn, hj = 0.0f;
    if (hh != nullptr) {
        const float *hr = hh + (size_t)n * 3 * H;
        gr = hr[j]; gz = hr[H + j]; gn = hr[2 * H + j];
        hj = hprev[(size_t)n * ldh + j];
    } else { gr = b_hh[j]; gz = b_hh[H + j]; gn = b_hh[2 * H + j]; }
    const float r = sigmoid_f(xr[j] + gr);
    const float z = sigmoid_f(xr[H + j] + gz);
    const float nn = tanhf(xr[2 * H + j] + r * gn);
    out[(size_t)n * ldo + j] = (1.0f - z) * nn + z * hj;
}

// ---- persistent cluster kernel (tanh) ------------------------------------------------------------------------
constexpr int RC_NB = 16;     // utterances per cluster
constexpr int RC_HC = 64;     // hidden columns per CTA
constexpr int RC_THREADS = 512;   // 16 warps: 8-way split-K x 2 column halves (4 warps per scheduler hide the LDS latency)
constexpr int RC_KSPLIT = 8;

struct RnnClusterParams {
    const float *xproj; int ldxp;
    const float *w_hh;
    float *out; int ldo, col0;
    int T, N, H, CS, reverse;
    int s0, s1;          // steps [s0, s1) of the T-step sequence are computed by this launch (time chunking)
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(RC_THREADS, 1) rnn_tanh_cluster_kernel(const RnnClusterParams p) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int H = p.H, CS = p.CS;
    const int HS = H + 4;                          // padded h row (keeps 128-bit loads of 4 rows conflict-free)
    float *Ws = smem;                              // [H][64]
    float *hbuf = Ws + (size_t)H * RC_HC;          // [2][16][HS]
    float *red = hbuf + 2 * RC_NB * HS;            // [RC_KSPLIT][16][64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int group = blockIdx.x / CS;
    const int n0 = group * RC_NB;
    const int colbase = rank * RC_HC;

    // resident weights: Ws[k][c] = W_hh[k][colbase + c]
    for (int i = tid; i < H * (RC_HC / 4); i += RC_THREADS) {
        const int k = i / (RC_HC / 4), c4 = i % (RC_HC / 4);
        reinterpret_cast<float4 *>(Ws)[i] = __ldg(reinterpret_cast<const float4 *>(p.w_hh + (size_t)k * H + colbase) + c4);
    }
    // h before the first step of this launch: zeros for step 0 (h_0 = 0, RNN.h:16-17), otherwise the previous
    // chunk's last output row, read back from HBM
    for (int i = tid; i < 2 * RC_NB * HS; i += RC_THREADS) hbuf[i] = 0.0f;
    if (p.s0 > 0) {
        __syncthreads();
        const int tp = p.reverse ? p.T - p.s0 : p.s0 - 1;
        float *h0buf = hbuf + (size_t)(p.s0 & 1) * RC_NB * HS;
        for (int i = tid; i < RC_NB * (H / 4); i += RC_THREADS) {
            const int u = i / (H / 4), c4 = i % (H / 4);
            if (n0 + u < p.N)
                *reinterpret_cast<float4 *>(h0buf + (size_t)u * HS + c4 * 4) =
                    *reinterpret_cast<const float4 *>(p.out + ((size_t)tp * p.N + n0 + u) * p.ldo + p.col0 + c4 * 4);
        }
    }

    // compute-phase mapping: warp -> (K quarter, column half); lane -> (utterance set, 4-column group)
    const int kq = warp & (RC_KSPLIT - 1), ch = warp / RC_KSPLIT;
    const int cgp = lane & 7, ug = lane >> 3;      // utterances ug, ug+4, ug+8, ug+12
    const int kbeg = kq * (H / RC_KSPLIT), kend = kbeg + H / RC_KSPLIT;
    const float *wcol = Ws + ch * 32 + cgp * 4;
    // epilogue mapping: thread -> (utterance, 4 columns)
    const bool epi = tid < 256;                    // the first 8 warps also run the epilogue (one float4 each)
    const int eu = (tid & 255) >> 4, ec = (tid & 15) * 4;
    const int en = n0 + eu;
    const bool evalid = epi && en < p.N;

    float4 xnext = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const int t0 = p.reverse ? p.T - 1 - p.s0 : p.s0;
        if (evalid) xnext = __ldg(reinterpret_cast<const float4 *>(p.xproj + ((size_t)t0 * p.N + en) * p.ldxp + colbase + ec));
    }
    cluster.sync();   // weights + zeroed h visible, all CTAs of the cluster are running

    for (int s = p.s0; s < p.s1; s++) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const float *hc = hbuf + (size_t)(s & 1) * RC_NB * HS;
        float *hn = hbuf + (size_t)((s & 1) ^ 1) * RC_NB * HS;
        const float4 xcur = xnext;
        if (s + 1 < p.s1 && evalid) {
            const int tn = p.reverse ? t - 1 : t + 1;
            xnext = __ldg(reinterpret_cast<const float4 *>(p.xproj + ((size_t)tn * p.N + en) * p.ldxp + colbase + ec));
        }
        // ---- partial products over this warp's K quarter ------------------------------------------------
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
        const float *h0 = hc + (size_t)ug * HS;
#pragma unroll 2
        for (int k = kbeg; k < kend; k += 4) {
            float4 hv[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) hv[i] = *reinterpret_cast<const float4 *>(h0 + (size_t)(4 * i) * HS + k);
#pragma unroll
            for (int q = 0; q < 4; q++) wv[q] = *reinterpret_cast<const float4 *>(wcol + (size_t)(k + q) * RC_HC);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                acc[i][0] = fmaf(hv[i].x, wv[0].x, acc[i][0]); acc[i][1] = fmaf(hv[i].x, wv[0].y, acc[i][1]);
                acc[i][2] = fmaf(hv[i].x, wv[0].z, acc[i][2]); acc[i][3] = fmaf(hv[i].x, wv[0].w, acc[i][3]);
                acc[i][0] = fmaf(hv[i].y, wv[1].x, acc[i][0]); acc[i][1] = fmaf(hv[i].y, wv[1].y, acc[i][1]);
                acc[i][2] = fmaf(hv[i].y, wv[1].z, acc[i][2]); acc[i][3] = fmaf(hv[i].y, wv[1].w, acc[i][3]);
                acc[i][0] = fmaf(hv[i].z, wv[2].x, acc[i][0]); acc[i][1] = fmaf(hv[i].z, wv[2].y, acc[i][1]);
                acc[i][2] = fmaf(hv[i].z, wv[2].z, acc[i][2]); acc[i][3] = fmaf(hv[i].z, wv[2].w, acc[i][3]);
                acc[i][0] = fmaf(hv[i].w, wv[3].x, acc[i][0]); acc[i][1] = fmaf(hv[i].w, wv[3].y, acc[i][1]);
                acc[i][2] = fmaf(hv[i].w, wv[3].z, acc[i][2]); acc[i][3] = fmaf(hv[i].w, wv[3].w, acc[i][3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            *reinterpret_cast<float4 *>(red + ((size_t)kq * RC_NB + ug + 4 * i) * RC_HC + ch * 32 + cgp * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        __syncthreads();
        // ---- reduce the 4 K quarters, add xproj, tanh, publish --------------------------------------------
        float4 v = xcur;
        if (epi) {
#pragma unroll
            for (int q = 0; q < RC_KSPLIT; q++) {
                const float4 r = *reinterpret_cast<const float4 *>(red + ((size_t)q * RC_NB + eu) * RC_HC + ec);
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w);
        }
        if (epi && s + 1 < p.s1) {
            // broadcast this CTA's slice of h_t into every CTA's next-step buffer (DSMEM)
            float *dst_local = hn + (size_t)eu * HS + colbase + ec;
            for (int r = 0; r < CS; r++) {
                float *dst = cluster.map_shared_rank(dst_local, r);
                *reinterpret_cast<float4 *>(dst) = v;
            }
        }
        cluster_arrive();
        if (evalid) *reinterpret_cast<float4 *>(p.out + ((size_t)t * p.N + en) * p.ldo + p.col0 + colbase + ec) = v;
        cluster_wait();
    }
}

// ---- persistent cluster kernel, tensor-core variant (default) ---------------------------------------------------
// Same decomposition (cluster of H/64 CTAs per 16 utterances, one cluster barrier per step), but the weights never
// touch shared memory: every warp keeps its 8-column slice of W_hh for ALL k as bf16 hi/lo MMA B-fragments in
// registers (128 registers at H = 512), h lives in shared memory as bf16 hi/lo planes and is read with ldmatrix, and
// the step is 3 x H/16 mma.sync.m16n8k16 (hi*hi + hi*lo + lo*hi, fp32 accumulate: fp32-grade, like the projection).
// Shared memory traffic per step drops from W + h fragments (broadcast loads, ~6k wavefronts) to h only (2k).
#include <cuda_bf16.h>

constexpr int RM_THREADS = 256;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pair(float x, float y, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat16 xh = __float2bfloat16_rn(x), yh = __float2bfloat16_rn(y);
    const __nv_bfloat162 h2 = __halves2bfloat162(xh, yh);
    const __nv_bfloat162 l2 = __halves2bfloat162(__float2bfloat16_rn(x - __bfloat162float(xh)),
                                                 __float2bfloat16_rn(y - __bfloat162float(yh)));
    hi = *reinterpret_cast<const uint32_t *>(&h2);
    lo = *reinterpret_cast<const uint32_t *>(&l2);
}

template <int H>
__global__ void __launch_bounds__(RM_THREADS, 1) rnn_tanh_mma_kernel(const RnnClusterParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int KS = H / 16;                       // k-steps of the m16n8k16 MMA
    constexpr int HSB = H * 2 + 16;                  // bytes of one bf16 h row (+16: ldmatrix rows hit distinct banks)
    constexpr int PLANE = RC_NB * HSB;               // one plane (hi or lo) of one buffer
    // layout: [buf 0/1][plane hi/lo][16 rows][HSB] | staging [plane][16 rows][128 B]
    unsigned char *hbuf = smraw;
    uint32_t *stage = reinterpret_cast<uint32_t *>(smraw + 4 * PLANE);
    const int CS = p.CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int rank = (int)cluster.block_rank();
    const int group = blockIdx.x / CS;
    const int n0 = group * RC_NB;
    const int colbase = rank * RC_HC;

    // ---- resident weights: B fragments of W_hh[:, colbase + 8*warp + g] for every k-step, bf16 hi and lo ----------
    uint32_t bhi[KS][2], blo[KS][2];
    {
        const float *wc = p.w_hh + colbase + 8 * warp + g;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const int k = 16 * ks + 2 * tg;
            split_pair(__ldg(wc + (size_t)k * H), __ldg(wc + (size_t)(k + 1) * H), bhi[ks][0], blo[ks][0]);
            split_pair(__ldg(wc + (size_t)(k + 8) * H), __ldg(wc + (size_t)(k + 9) * H), bhi[ks][1], blo[ks][1]);
        }
    }
    // ---- h before the first step: zeros (h_0 = 0, RNN.h:16-17) or the previous chunk's last output row ------------
    for (int i = tid; i < 4 * PLANE / 4; i += RM_THREADS) reinterpret_cast<uint32_t *>(hbuf)[i] = 0u;
    if (p.s0 > 0) {
        __syncthreads();
        const int tp = p.reverse ? p.T - p.s0 : p.s0 - 1;
        unsigned char *b0 = hbuf + (size_t)(p.s0 & 1) * 2 * PLANE;
        for (int i = tid; i < RC_NB * (H / 2); i += RM_THREADS) {
            const int u = i / (H / 2), c = (i % (H / 2)) * 2;
            if (n0 + u < p.N) {
                const float2 v = *reinterpret_cast<const float2 *>(p.out + ((size_t)tp * p.N + n0 + u) * p.ldo + p.col0 + c);
                uint32_t hi, lo;
                split_pair(v.x, v.y, hi, lo);
                *reinterpret_cast<uint32_t *>(b0 + (size_t)u * HSB + c * 2) = hi;
                *reinterpret_cast<uint32_t *>(b0 + PLANE + (size_t)u * HSB + c * 2) = lo;
            }
        }
    }
    // ---- per-lane roles --------------------------------------------------------------------------------------------
    // ldmatrix: lanes 0-7 rows 0-7 k+0 | 8-15 rows 8-15 k+0 | 16-23 rows 0-7 k+8 | 24-31 rows 8-15 k+8  (A fragment order)
    const uint32_t lm_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * HSB + ((lane >> 4) & 1) * 16);
    const uint32_t hbase = (uint32_t)__cvta_generic_to_shared(hbuf);
    // accumulator fragment: rows g and g+8 (utterances), columns 8*warp + 2*tg + {0,1}
    const int ecol = 8 * warp + 2 * tg;
    const int u0 = n0 + g, u1 = n0 + g + 8;
    const bool v0 = u0 < p.N, v1 = u1 < p.N;
    float2 xn0 = make_float2(0.f, 0.f), xn1 = make_float2(0.f, 0.f);
    {
        const int t0 = p.reverse ? p.T - 1 - p.s0 : p.s0;
        if (v0) xn0 = __ldg(reinterpret_cast<const float2 *>(p.xproj + ((size_t)t0 * p.N + u0) * p.ldxp + colbase + ecol));
        if (v1) xn1 = __ldg(reinterpret_cast<const float2 *>(p.xproj + ((size_t)t0 * p.N + u1) * p.ldxp + colbase + ecol));
    }
    cluster.sync();

    for (int s = p.s0; s < p.s1; s++) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const uint32_t cur = hbase + (uint32_t)((s & 1) * 2 * PLANE) + lm_off;
        unsigned char *nxt = hbuf + (size_t)((s & 1) ^ 1) * 2 * PLANE;
        const float2 xc0 = xn0, xc1 = xn1;
        if (s + 1 < p.s1) {
            const int tn = p.reverse ? t - 1 : t + 1;
            if (v0) xn0 = __ldg(reinterpret_cast<const float2 *>(p.xproj + ((size_t)tn * p.N + u0) * p.ldxp + colbase + ecol));
            if (v1) xn1 = __ldg(reinterpret_cast<const float2 *>(p.xproj + ((size_t)tn * p.N + u1) * p.ldxp + colbase + ecol));
        }
        // ---- h_{t-1} * W_hh slice: three independent accumulator chains (one per split term) ----------------------
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            uint32_t ah[4], al[4];
            ldmatrix_x4(ah, cur + ks * 32);
            ldmatrix_x4(al, cur + PLANE + ks * 32);
            mma_bf16_16816(c0, ah, bhi[ks][0], bhi[ks][1]);
            mma_bf16_16816(c1, ah, blo[ks][0], blo[ks][1]);
            mma_bf16_16816(c2, al, bhi[ks][0], bhi[ks][1]);
        }
        // ---- + xproj, tanh, publish --------------------------------------------------------------------------------
        float2 r0, r1;
        r0.x = tanhf(xc0.x + (c0[0] + (c1[0] + c2[0]))); r0.y = tanhf(xc0.y + (c0[1] + (c1[1] + c2[1])));
        r1.x = tanhf(xc1.x + (c0[2] + (c1[2] + c2[2]))); r1.y = tanhf(xc1.y + (c0[3] + (c1[3] + c2[3])));
        if (s + 1 < p.s1) {
            uint32_t hi, lo;
            split_pair(r0.x, r0.y, hi, lo);
            stage[g * 32 + 4 * warp + tg] = hi; stage[512 + g * 32 + 4 * warp + tg] = lo;
            split_pair(r1.x, r1.y, hi, lo);
            stage[(g + 8) * 32 + 4 * warp + tg] = hi; stage[512 + (g + 8) * 32 + 4 * warp + tg] = lo;
        }
        __syncthreads();
        if (s + 1 < p.s1) {
            // 4 KB slice (2 planes x 16 rows x 128 B) -> every CTA's next-step buffer: one 16-byte chunk per thread
            const int pl = tid >> 7, row = (tid >> 3) & 15, c16 = tid & 7;
            const int4 chunk = reinterpret_cast<const int4 *>(stage)[tid];
            unsigned char *dst_local = nxt + (size_t)pl * PLANE + (size_t)row * HSB + colbase * 2 + c16 * 16;
            for (int r = 0; r < CS; r++) *reinterpret_cast<int4 *>(cluster.map_shared_rank(dst_local, r)) = chunk;
        }
        cluster_arrive();
        if (v0) *reinterpret_cast<float2 *>(p.out + ((size_t)t * p.N + u0) * p.ldo + p.col0 + colbase + ecol) = r0;
        if (v1) *reinterpret_cast<float2 *>(p.out + ((size_t)t * p.N + u1) * p.ldo + p.col0 + colbase + ecol) = r1;
        cluster_wait();
    }
}

template <int H>
static int launch_rnn_mma(const RnnClusterParams &p, int groups, cudaStream_t st) {
    // the kernel needs 4 h planes + 4 KB of staging; asking for 200 KB keeps every other CTA (GEMM, Linear, decoder) off
    // the SMs of the cluster, whose per-step barrier makes it latency critical
    size_t smem = 4 * (size_t)RC_NB * (H * 2 + 16) + 4096;
    if (smem < 200 * 1024) smem = 200 * 1024;
    GASR_CUDA(cudaFuncSetAttribute(rnn_tanh_mma_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * p.CS);
    cfg.blockDim = dim3(RM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GASR_CUDA(cudaLaunchKernelEx(&cfg, rnn_tanh_mma_kernel<H>, p));
    return GASR_OK;
}

static bool cluster_kernel_supported(const gasr_ctx *ctx, const RnnLayerArgs &a) {
    if (a.cell != GASR_CELL_TANH || !ctx->cluster_ok) return false;
    const int H = a.H;
    if (H != 64 && H != 128 && H != 256 && H != 512) return false;
    if (a.ldxp % 4 != 0 || a.ldo % 4 != 0 || a.col0 % 4 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.xproj) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15) ||
        (reinterpret_cast<uintptr_t>(a.w_hh) & 15))
        return false;
    return true;
}

static size_t cluster_smem_bytes(int H) {
    return sizeof(float) * ((size_t)H * RC_HC + 2 * RC_NB * (H + 4) + RC_KSPLIT * RC_NB * RC_HC);
}

int launch_rnn_recurrence(gasr_ctx *ctx, const RnnLayerArgs &a, cudaStream_t st) {
    GASR_CHECK(a.xproj && a.w_hh && a.out, "rnn_recurrence: null operand");
    GASR_CHECK(a.T >= 0 && a.N >= 0 && a.H >= 1, "rnn_recurrence: bad shape");
    GASR_CHECK(a.cell == GASR_CELL_TANH || (a.cell == GASR_CELL_GRU && a.b_hh), "rnn_recurrence: bad cell");
    if (a.T == 0 || a.N == 0) return GASR_OK;
    {
        // Many utterances: groups of 128 on the tcgen05 tensor cores (rnn_wide.cu).  The planes it exchanges h through live
        // in the ctx workspace, so only whole-sequence calls take this path here; the wave engine (asr_wave.cu) owns its
        // planes and resumes chunk by chunk.
        const char force = ctx->opt.rnn;
        const bool whole = a.s0 == 0 && (a.s1 == 0 || a.s1 == a.T);
        const bool want = force ? force == 'w' : a.N >= 96;
        if (want && whole && a.cell == GASR_CELL_TANH && !a.reverse && rnn_wide_supported(ctx, a.H) && cluster_kernel_supported(ctx, a)) {
            const bool pair = ctx->opt.rnn_pair && rnn_wide2_supported(ctx, a.H) && a.N > 128;   // CTA pairs: groups of 256
            const int unit = pair ? 256 : 128;
            const int Npad = ceil_div(a.N, unit) * unit;
            const size_t pb = rnn_wide_plane_bytes(a.T, Npad, a.H), wb = xproj_tc_w_bytes(a.H, a.H);
            GASR_TRY(ws_reserve(ctx, ctx->ws_wide, 2 * pb + wb + 2048));
            unsigned char *base = static_cast<unsigned char *>(ctx->ws_wide.ptr);
            RnnWidePlan pl;
            GASR_TRY(rnn_wide_plan(ctx, pl, a.w_hh, a.T, a.N, a.H, base + 2 * pb, base, st, unit));
            RnnWideRun r = {};
            r.s0 = 0; r.s1 = a.T; r.xp = a.xproj; r.ldxp = a.ldxp; r.xp_rows_per_frame = a.N;
            r.out = a.out; r.ldo = a.ldo; r.col0 = a.col0; r.out_rows_per_frame = a.N;
            r.groups_per_cluster = ctx->opt.rnn_groups > 0 ? ctx->opt.rnn_groups : 2; r.multicast = ctx->opt.rnn_mc;
            return pair ? launch_rnn_wide2(ctx, pl, r, st) : launch_rnn_wide(ctx, pl, r, st);
        }
    }
    if (cluster_kernel_supported(ctx, a)) {
        RnnClusterParams p;
        p.xproj = a.xproj; p.ldxp = a.ldxp; p.w_hh = a.w_hh; p.out = a.out; p.ldo = a.ldo; p.col0 = a.col0;
        p.T = a.T; p.N = a.N; p.H = a.H; p.CS = a.H / RC_HC; p.reverse = a.reverse;
        p.s0 = a.s0; p.s1 = a.s1 > 0 ? a.s1 : a.T;
        const int groups = ceil_div(a.N, RC_NB);
        const char force_c = ctx->opt.rnn;
        const char *force = force_c ? &force_c : nullptr;
        // whole-sequence calls: the latency-optimised persistent kernel (rnn_stream.cu); chunk-resumed calls and the
        // envelope outside it (more than 16 co-resident clusters) stay on the kernels below
        if (!(force && (force[0] == 'f' || force[0] == 'm')) && !a.reverse && p.s0 == 0 && p.s1 == a.T &&
            rnn_stream_supported(ctx, a.H, a.N, 1) && a.col0 == 0) {
            RnnStreamParams sp = {};
            sp.T = a.T; sp.N = a.N; sp.L = 1; sp.frames_per_block = 1; sp.xp_need = 0; sp.error = nullptr;
            sp.nsub = rnn_stream_default_nsub(a.N);
            sp.groups = sp.nsub == 4 ? ceil_div(a.N, 32) : groups;
            sp.layer[0].xproj = a.xproj; sp.layer[0].ldxp = a.ldxp; sp.layer[0].w_hh = a.w_hh;
            sp.layer[0].out = a.out; sp.layer[0].ldo = a.ldo;
            return launch_rnn_stream(ctx, sp, a.H, st);
        }
        if (!(force && force[0] == 'f') && a.ldxp % 2 == 0 && a.ldo % 2 == 0) {
            int rc = GASR_OK;
            if (a.H == 512) rc = launch_rnn_mma<512>(p, groups, st);
            else if (a.H == 256) rc = launch_rnn_mma<256>(p, groups, st);
            else if (a.H == 128) rc = launch_rnn_mma<128>(p, groups, st);
            else rc = launch_rnn_mma<64>(p, groups, st);
            if (rc == GASR_OK) ctx->launches += 1;
            return rc;
        }
        const size_t smem = cluster_smem_bytes(a.H);
        GASR_CUDA(cudaFuncSetAttribute(rnn_tanh_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(groups * p.CS);
        cfg.blockDim = dim3(RC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = p.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        GASR_CUDA(cudaLaunchKernelEx(&cfg, rnn_tanh_cluster_kernel, p));
        ctx->launches += 1;
        return GASR_OK;
    }
    GASR_CHECK(a.N <= 65535, "rnn_recurrence: batch too large for the per-step path");
    const char force_gc = ctx->opt.gru;
    const char *force_g = force_gc ? &force_gc : nullptr;
    const bool gru_batched = a.cell == GASR_CELL_GRU && !(force_g && force_g[0] == 's') && a.N >= 32 && a.s0 == 0 &&
                             (a.s1 == 0 || a.s1 == a.T);
    const bool gru_fused = gru_batched && !(force_g && force_g[0] == 'g') && gru_tc_supported(a.N, a.H, a.ldxp, a.ldo, a.col0) &&
                           ((reinterpret_cast<uintptr_t>(a.xproj) | reinterpret_cast<uintptr_t>(a.out) |
                             reinterpret_cast<uintptr_t>(a.b_hh)) & 15) == 0;
    if (gru_fused && a.precision == GASR_PREC_BF16 && !(force_g && force_g[0] == 't') &&
        gru_seq_supported(ctx, a.T, a.N, a.H, a.ldxp, a.ldo, a.col0)) {
        // bf16 mode: ONE persistent launch for the whole sequence, W_hh (fp16) resident in shared memory (gru_seq.cu)
        Workspace &wsg = ctx->ws_sel ? ctx->ws_gru_b : ctx->ws_gru;
        GASR_TRY(ws_reserve(ctx, wsg, gru_seq_ws_bytes(a.N, a.H)));
        return launch_gru_seq(ctx, a, wsg.ptr, st);
    }
    if (gru_fused || (gru_batched && xproj_tc_supported(a.N, a.H, 3 * a.H))) {
        // GRU with a real batch, one launch per timestep (gru_tc.cu): tcgen05 GEMM h_{t-1} * W_hh (fp32-grade 3 x bf16
        // split, permuted W_hh^T prepared once, TMA descriptors reused) with the gate math in its epilogue, which also
        // writes the bf16 planes the next step reads.  (GASR_GRU=g: the earlier three-launch form -- split, GEMM
        // hh = h W_hh + b_hh, gate kernel; GASR_GRU=s: one SIMT kernel per step.)
        const int H = a.H, G3 = 3 * a.H;
        Workspace &wsg = ctx->ws_sel ? ctx->ws_gru_b : ctx->ws_gru;
        XprojTcPlan pl;
        GruTcPlan gp;
        float *hh = nullptr;
        unsigned char *base = nullptr;
        if (gru_fused) {
            const size_t wb = align_up(gru_tc_w_bytes(H), 1024);
            GASR_TRY(ws_reserve(ctx, wsg, wb + 2 * gru_tc_plane_bytes(a.N, H) + 1024));
            base = static_cast<unsigned char *>(wsg.ptr);
            GASR_TRY(gru_tc_prepare(ctx, gp, a.w_hh, a.N, H, base, base + wb, st));
        } else {
            const size_t wb = align_up(xproj_tc_w_bytes(H, G3), 1024), ab = align_up(xproj_tc_a_bytes(a.N, H), 1024);
            const size_t hb = align_up(sizeof(float) * (size_t)a.N * G3, 1024);
            GASR_TRY(ws_reserve(ctx, wsg, wb + ab + hb + 1024));
            base = static_cast<unsigned char *>(wsg.ptr);
            hh = reinterpret_cast<float *>(base + wb + ab);
            GASR_TRY(xproj_tc_prepare_weights(ctx, a.w_hh, H, G3, base, st));
            GASR_TRY(xproj_tc_plan(pl, a.N, H, G3, base, base + wb));
        }
        dim3 ggrid(ceil_div(H, 128), a.N);
        auto issue_steps = [&]() -> int {
            for (int s = 0; s < a.T; s++) {
                const int t = a.reverse ? a.T - 1 - s : s;
                const int tp = a.reverse ? t + 1 : t - 1;
                const float *xp = a.xproj + (size_t)t * a.N * a.ldxp;
                float *o = a.out + (size_t)t * a.N * a.ldo + a.col0;
                const float *hp = s == 0 ? nullptr : a.out + (size_t)tp * a.N * a.ldo + a.col0;
                if (gru_fused) {
                    GASR_TRY(gru_tc_step(ctx, gp, s & 1, xp, a.ldxp, a.b_hh, hp, a.ldo, o, a.ldo, s > 0 && !ctx->opt.gru_no_pdl, st));
                    continue;
                }
                if (s > 0) GASR_TRY(xproj_tc_run(ctx, pl, hp, a.ldo, a.b_hh, hh, G3, GASR_PREC_FP32, st));
                gru_gate_kernel<<<ggrid, 128, 0, st>>>(xp, a.ldxp, s > 0 ? hh : nullptr, a.b_hh, hp, a.ldo, o, a.ldo, a.N, H);
                ctx->launches += 1;
            }
            return GASR_OK;
        };
        // The step loop is launch-bound from the host.  The whole T-step sequence is captured once into a CUDA graph
        // (operands are stable across calls: workspaces, weights, layer buffers) and replayed; the weight preparation
        // and the zeroing of the h planes above stay outside the graph (they run on every call).
        if (a.T >= 64 && !ctx->opt.no_graph) {
            gasr_ctx::StepGraph key = {};
            key.k[0] = a.xproj; key.k[1] = a.out; key.k[2] = a.w_hh; key.k[3] = base; key.k[4] = a.b_hh;
            const int dims[8] = {a.T, a.N, H, a.reverse, a.col0, a.ldo, a.ldxp, gru_fused ? 1 : 0};
            memcpy(key.dims, dims, sizeof(dims));
            for (auto &g : ctx->step_graphs)
                if (memcmp(g.k, key.k, sizeof(key.k)) == 0 && memcmp(g.dims, key.dims, sizeof(key.dims)) == 0) {
                    GASR_CUDA(cudaGraphLaunch(g.exec, st));
                    ctx->launches += gru_fused ? (long long)a.T : 3LL * a.T - 2;
                    return GASR_OK;
                }
            if (!gru_fused && !(ctx->attr_mask & 1024u)) {   // function attributes cannot be set while capturing: warm the GEMM up
                GASR_TRY(xproj_tc_run(ctx, pl, a.out + a.col0, a.ldo, a.b_hh, hh, G3, GASR_PREC_FP32, st));
            }
            cudaGraph_t graph = nullptr;
            const long long launches0 = ctx->launches;
            GASR_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int rc = issue_steps();
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != GASR_OK || ce != cudaSuccess || graph == nullptr) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                ctx->launches = launches0;
                GASR_TRY(issue_steps());              // capture unavailable: plain launches
                GASR_CUDA(cudaGetLastError());
                return GASR_OK;
            }
            GASR_CUDA(cudaGraphInstantiate(&key.exec, graph, 0));
            cudaGraphDestroy(graph);
            if (ctx->step_graphs.size() >= 64) {      // bounded cache
                cudaGraphExecDestroy(ctx->step_graphs.front().exec);
                ctx->step_graphs.erase(ctx->step_graphs.begin());
            }
            ctx->step_graphs.push_back(key);
            GASR_CUDA(cudaGraphLaunch(key.exec, st));
            return GASR_OK;
        }
        GASR_TRY(issue_steps());
        GASR_CUDA(cudaGetLastError());
        return GASR_OK;
    }
    if (a.cell == GASR_CELL_TANH && !(ctx->opt.rnn == 'f') && rnn_resident_supported(ctx, a)) {
        // wide hidden layer, a few utterances (cfg1: H = 2048, one utterance): W_hh resident in the shared memory of the whole
        // GPU for the sequence, fp32, one launch (rnn_resident.cu)
        Workspace &wsg = ctx->ws_sel ? ctx->ws_gru_b : ctx->ws_gru;
        GASR_TRY(ws_reserve(ctx, wsg, 256));
        return launch_rnn_resident(ctx, a, wsg.ptr, st);
    }
    // fallback: one kernel per timestep
    dim3 grid(ceil_div(a.H, 128), a.N);
    for (int s = a.s0; s < (a.s1 > 0 ? a.s1 : a.T); s++) {
        const int t = a.reverse ? a.T - 1 - s : s;
        const int tp = a.reverse ? t + 1 : t - 1;
        const float *xp = a.xproj + (size_t)t * a.N * a.ldxp;
        float *o = a.out + (size_t)t * a.N * a.ldo + a.col0;
        const float *hp = s == 0 ? nullptr : a.out + (size_t)tp * a.N * a.ldo + a.col0;
        if (a.cell == GASR_CELL_TANH)
            rnn_tanh_step_kernel<<<grid, 128, 0, st>>>(xp, a.ldxp, hp, a.ldo, a.w_hh, o, a.ldo, a.N, a.H);
        else
            gru_step_kernel<<<grid, 128, 0, st>>>(xp, a.ldxp, hp, a.ldo, a.w_hh, a.b_hh, o, a.ldo, a.N, a.H);
        ctx->launches += 1;
    }
    GASR_CUDA(cudaGetLastError());
    return GASR_OK;
}

}  // namespace gasr
