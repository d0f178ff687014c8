// gru_seq.cu -- the GRU recurrence of one (layer, direction) as ONE persistent kernel with W_hh resident in shared memory
// (GASR_PREC_BF16 mode; BASELINE.json cfg3: 5-layer bidirectional GRU, H = 800, 256 utterances).
//
// Stands behind the time loop of RNN::forward (reference RNN.cu:9-30, one cell call per timestep); the GRU equations are
// torch.nn.GRU's, as in oracle/am_ref.c (gate order r, z, n).  The per-timestep kernel (gru_tc.cu) costs 14 us per step
// at cfg3 widths: every launch re-reads its 96 x 832 slice of W_hh^T (hi + lo planes, 320 KB per CTA) through L2 and pays
// the launch / dependency gap.  Here
//   * CTA (tile, rb) owns 32 hidden units (the r, z and n columns of W_hh for them: a [96 x Kp] K-major slice of the
//     permuted W_hh^T, SINGLE-PLANE FP16, 156 KB at H = 800) for a block rb of 128 utterances and keeps the slice in
//     shared memory for all T steps;
//   * per step it pulls h_{t-1} of its 128 utterances ([128 x Kp] fp16, 208 KB) through a TMA ring, runs
//     tcgen05.mma (M = 128, N = 96, fp32 accumulator in TMEM), and its eight epilogue warps (thread = utterance row, 16
//     units) add the input projection and biases, apply the gates, keep h_{t-1} of their units in REGISTERS, and write
//     h_t as fp32 (the layer output) and as fp16 into the other plane buffer;
//   * the CTAs of a row block exchange h_t through those planes in L2: every epilogue warp releases its rows with
//     __threadfence + atomicAdd on a per-row-block counter, the TMA producers acquire the count of step t before they load.
// Grid = ceil(H / 32) x ceil(N / 128) CTAs (25 x 2 at cfg3), all co-resident (one per SM, checked by the launcher);
// both directions of a layer run concurrently on two streams (100 of 148 SMs).
//
// Arithmetic of this mode: recurrent operands (h_{t-1}, W_hh) rounded to fp16 (11-bit significand, 8x finer than the bf16
// of the input projection this mode already uses), fp32 accumulation, fp32 gate math and fp32 state carried between steps
// in registers.  Stated tolerance of the mode: 2e-2 on the log-probabilities (tests/test_gpu_parity.py, test_gpu_sizes.py).
// Every wait is bounded (4 s -> __trap): a protocol bug or a CTA that never became resident is a launch failure, not a hang.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "rnn_wide_dev.cuh"
#include "tc_common.cuh"

namespace gasr {

constexpr int GS_UNITS = 32;                       // hidden units per CTA
constexpr int GS_BN = 3 * GS_UNITS;                // accumulator columns: r | z | n
constexpr int GS_W_TILE = GS_BN * TC_BK * 2;       // 12 KB: [96 x 64] fp16
constexpr int GS_A_TILE = TC_BM * TC_BK * 2;       // 16 KB: [128 x 64] fp16
constexpr int GS_THREADS = 320;                    // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int GS_EPI_WARPS = 8;
constexpr int GS_MAX_SMEM = 227 * 1024;

struct GruSeqParams {
    int T, N, H, KB, Kp, reverse, nst;             // KB = Kp / 64 k-blocks, nst = ring stages
    const float *xp; int ldxp;                     // x * W_ih + b_ih, [T * N, >= 3H]
    const float *b_hh;                             // [3H]
    float *out; int ldo;                           // h fp32, [T * N, ldo], already offset to this direction's columns
    __half *plane[2];                              // fp16 planes of h, [N, Kp] each (ping-pong by step parity)
    unsigned *cnt;                                 // [row blocks] CTAs that released their columns of h, zeroed before the launch
    unsigned long long *trace;                     // instrumented build only (make TRACE=1): cycle sums of CTA (0, 0)
};

#ifdef GASR_RW_TRACE
// cycle sums per role of CTA (0, 0): 0 producer waits for the peers' h, 1 producer waits for a free ring stage, 2 MMA issuer waits
// for a full stage, 3 epilogue waits for the accumulator, 4 epilogue math + stores, 5 epilogue release (barrier + fence + atomic)
#define GS_T0(v) const long long v = clock64()
#define GS_T1(v, k) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) <= 2) p.trace[k] += (unsigned long long)(clock64() - v); } while (0)
#else
#define GS_T0(v) do { } while (0)
#define GS_T1(v, k) do { } while (0)
#endif

__device__ __forceinline__ float gs_sigmoid(float v) {       // SFU: ex2.approx + rcp.approx (absolute error ~1e-7)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return r;
}
__device__ __forceinline__ void gs_tmem_ld16(uint32_t (&v)[16], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ unsigned gs_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(GS_THREADS, 1)
gru_seq_kernel(const __grid_constant__ CUtensorMap map_h0, const __grid_constant__ CUtensorMap map_h1,
               const __grid_constant__ CUtensorMap map_w, const GruSeqParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t w_smem = (raw + 1023u) & ~1023u;                      // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t ring = w_smem + (uint32_t)p.KB * GS_W_TILE;
    const uint32_t bars = ring + (uint32_t)p.nst * GS_A_TILE;            // full[4], empty[4], tfull, wfull, tmem slot
    const uint32_t full0 = bars, empty0 = bars + 32, tfull = bars + 64, wfull = bars + 72;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem_raw + (bars + 80 - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, rb = blockIdx.y, m0 = rb * TC_BM;
    const int n_tiles = gridDim.x;
    const int NST = p.nst;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1); mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: the resident W_hh^T slice once, then h_{t-1} of this row block for every step =====
        if (lane == 0) {
            mbar_expect_tx(wfull, (uint32_t)p.KB * GS_W_TILE);
            for (int kb = 0; kb < p.KB; kb++) tma_load_2d(w_smem + (uint32_t)kb * GS_W_TILE, &map_w, wfull, kb * TC_BK, tile * GS_BN);
            const unsigned *cnt = p.cnt + rb;
            int it = 0;
            for (int s = 1; s < p.T; s++) {                              // step 0: h_{-1} = 0, nothing to multiply
                // every CTA of this row block has stored its columns of h_{s-1} (and so has finished reading h_{s-2})
                const unsigned need = (unsigned)s * (unsigned)n_tiles;
                GS_T0(tp);
                if (gs_ld_acquire(cnt) < need) {
                    const unsigned long long t0 = rw_now_ns();
                    unsigned spins = 0;
                    while (gs_ld_acquire(cnt) < need)
                        if ((++spins & 1023u) == 0 && rw_now_ns() - t0 > RW_TIMEOUT_NS) __trap();
                }
                GS_T1(tp, 0);
                asm volatile("fence.proxy.async;" ::: "memory");         // generic-proxy stores of the peers -> TMA reads
                const CUtensorMap *mh = ((s - 1) & 1) ? &map_h1 : &map_h0;
                for (int kb = 0; kb < p.KB; kb++, it++) {
                    const int st = it % NST;
                    GS_T0(te);
                    rw_wait(empty0 + 8 * st, ((uint32_t)(it / NST) & 1u) ^ 1u);
                    GS_T1(te, 1);
                    mbar_expect_tx(full0 + 8 * st, GS_A_TILE);
                    tma_load_2d(ring + (uint32_t)st * GS_A_TILE, mh, full0 + 8 * st, kb * TC_BK, m0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[128 x 96] = h_{t-1}[128 x Kp] * W_slice^T, fp16 operands, fp32 accumulator in TMEM =====
        if (lane == 0) {
            // instruction descriptor: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N = 96, M = 128
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(GS_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            rw_wait(wfull, 0);
            int it = 0;
            for (int s = 1; s < p.T; s++) {
                for (int kb = 0; kb < p.KB; kb++, it++) {
                    const int st = it % NST;
                    GS_T0(tf);
                    rw_wait(full0 + 8 * st, (uint32_t)(it / NST) & 1u);
                    GS_T1(tf, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = umma_desc_sw128(ring + (uint32_t)st * GS_A_TILE);
                    const uint64_t db = umma_desc_sw128(w_smem + (uint32_t)kb * GS_W_TILE);
#pragma unroll
                    for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                        const uint64_t adv = (uint64_t)(k4 * 32 >> 4);
                        umma_bf16(tmem_base, da + adv, db + adv, idesc, (kb | k4) != 0);   // kind::f16; the descriptor selects fp16
                    }
                    umma_commit(empty0 + 8 * st);
                }
                umma_commit(tfull);
            }
        }
    } else {
        // ===== epilogue: thread = utterance (accumulator row), 16 hidden units; warps 2-5: units 0-15, warps 6-9: units 16-31
        //       of the tile (a warp may only read the TMEM lane quadrant warp % 4) =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = m0 + q * 32 + lane;
        const bool live = row < p.N;
        const int jb = tile * GS_UNITS + half * 16;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16);
        float hprev[16];
#pragma unroll
        for (int e = 0; e < 16; e++) hprev[e] = 0.0f;
        unsigned *cnt = p.cnt + rb;
        for (int s = 0; s < p.T; s++) {
            const int t = p.reverse ? p.T - 1 - s : s;
            const size_t grow = (size_t)t * p.N + (live ? row : 0);
            const float *xr = p.xp + grow * p.ldxp;
            // this step's input projections do not depend on the MMAs: requested now, they arrive while the h exchange,
            // the TMA ring and the MMAs of the step run
            float4 xr4[4], xz4[4], xn4[4];
#pragma unroll
            for (int g4 = 0; g4 < 4; g4++) {
                const int j = jb + 4 * g4;
                const int jc = (live && j < p.H) ? j : 0;                 // H % 4 == 0: a group of four is all in or all out
                xr4[g4] = __ldcs(reinterpret_cast<const float4 *>(xr + jc));
                xz4[g4] = __ldcs(reinterpret_cast<const float4 *>(xr + p.H + jc));
                xn4[g4] = __ldcs(reinterpret_cast<const float4 *>(xr + 2 * p.H + jc));
            }
            uint32_t ar[16], az[16], an[16];
            GS_T0(ta);
            if (s > 0) {
                rw_wait(tfull, (uint32_t)(s - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                gs_tmem_ld16(ar, taddr);
                gs_tmem_ld16(az, taddr + GS_UNITS);
                gs_tmem_ld16(an, taddr + 2 * GS_UNITS);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // ordered before the release below
            } else {
#pragma unroll
                for (int e = 0; e < 16; e++) { ar[e] = 0u; az[e] = 0u; an[e] = 0u; }   // h_0 = 0: hh = b_hh exactly
            }
            GS_T1(ta, 3);
            GS_T0(tc);
            if (live) {
                float *orow = p.out + grow * p.ldo;
                __half *prow = ((s & 1) ? p.plane[1] : p.plane[0]) + (size_t)row * p.Kp;
#pragma unroll
                for (int g4 = 0; g4 < 4; g4++) {
                    const int j = jb + 4 * g4;
                    if (j >= p.H) break;
                    const float4 br4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + j));
                    const float4 bz4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + p.H + j));
                    const float4 bn4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + 2 * p.H + j));
                    const float xrv[4] = {xr4[g4].x, xr4[g4].y, xr4[g4].z, xr4[g4].w}, xzv[4] = {xz4[g4].x, xz4[g4].y, xz4[g4].z, xz4[g4].w};
                    const float xnv[4] = {xn4[g4].x, xn4[g4].y, xn4[g4].z, xn4[g4].w};
                    const float brv[4] = {br4.x, br4.y, br4.z, br4.w}, bzv[4] = {bz4.x, bz4.y, bz4.z, bz4.w};
                    const float bnv[4] = {bn4.x, bn4.y, bn4.z, bn4.w};
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float gr = __uint_as_float(ar[4 * g4 + e]) + brv[e];
                        const float gz = __uint_as_float(az[4 * g4 + e]) + bzv[e];
                        const float gn = __uint_as_float(an[4 * g4 + e]) + bnv[e];
                        const float r = gs_sigmoid(xrv[e] + gr);
                        const float z = gs_sigmoid(xzv[e] + gz);
                        const float nn = rw_tanh(xnv[e] + r * gn);
                        o[e] = (1.0f - z) * nn + z * hprev[4 * g4 + e];
                        hprev[4 * g4 + e] = o[e];
                    }
                    __stcs(reinterpret_cast<float4 *>(orow + j), make_float4(o[0], o[1], o[2], o[3]));
                    const __half2 p0 = __floats2half2_rn(o[0], o[1]), p1 = __floats2half2_rn(o[2], o[3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t *>(&p0); pk.y = *reinterpret_cast<const uint32_t *>(&p1);
                    *reinterpret_cast<uint2 *>(prow + j) = pk;
                }
            }
            GS_T1(tc, 4);
            if (s + 1 < p.T) {
                // release this CTA's 128 rows x 32 units of h_s (and its TMEM reads) to the row block's TMA producers: the
                // eight epilogue warps meet at a named barrier, then ONE thread fences and counts (25 atomics per step and
                // counter instead of 200)
                GS_T0(tr);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && lane == 0) { __threadfence(); atomicAdd(cnt, 1u); }
                GS_T1(tr, 5);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
    }
}

// W_hh[H, 3H] (reference layout [in, out], gates r | z | n) -> permuted W^T as one fp16 plane [tiles * 96, Kp]:
// row tile * 96 + g * 32 + u  =  column g * H + tile * 32 + u of W_hh (zero beyond H)
__global__ void gru_seq_perm_kernel(const float *__restrict__ w, int H, int Kp, int rows, __half *__restrict__ wt) {
    const size_t total = (size_t)rows * Kp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int prow = (int)(i / Kp), k = (int)(i % Kp);
        const int tile = prow / GS_BN, rem = prow - tile * GS_BN, g = rem / GS_UNITS, u = rem - g * GS_UNITS;
        const int j = tile * GS_UNITS + u;
        wt[i] = __float2half_rn((k < H && j < H) ? w[(size_t)k * 3 * H + (size_t)g * H + j] : 0.0f);
    }
}

static int gs_kp(int H) { return ceil_div(H, TC_BK) * TC_BK; }
static int gs_rows(int H) { return ceil_div(H, GS_UNITS) * GS_BN; }
static size_t gs_w_bytes(int H) { return align_up((size_t)gs_rows(H) * gs_kp(H) * 2, 1024); }
static size_t gs_plane_bytes(int N, int H) { return align_up((size_t)N * gs_kp(H) * 2, 1024); }
static int gs_stages(int H) {
    const int left = GS_MAX_SMEM - 1024 - 256 - (gs_kp(H) / TC_BK) * GS_W_TILE;
    const int n = left / GS_A_TILE;
    return n > 4 ? 4 : n;
}

// The W_hh slice of a CTA must fit in shared memory next to a ring of at least two stages, and every CTA of both directions
// of a layer must be resident at once (they wait for each other's h columns).
bool gru_seq_supported(const gasr_ctx *ctx, int T, int N, int H, int ldxp, int ldo, int col0) {
    if (!(T >= 2 && N >= 1 && H >= 32 && H % 4 == 0 && ldxp % 4 == 0 && ldo % 4 == 0 && col0 % 4 == 0)) return false;
    if (gs_stages(H) < 2) return false;
    return 2 * ceil_div(H, GS_UNITS) * ceil_div(N, TC_BM) <= ctx->sm_count;
}

size_t gru_seq_ws_bytes(int N, int H) { return gs_w_bytes(H) + 2 * gs_plane_bytes(N, H) + 1024; }

// One launch for all T steps.  ws: gru_seq_ws_bytes(N, H) of device memory (weights plane, two h planes, counters).
int launch_gru_seq(gasr_ctx *ctx, const RnnLayerArgs &a, void *ws, cudaStream_t st) {
    const int H = a.H, N = a.N, Kp = gs_kp(H), rows = gs_rows(H);
    unsigned char *base = static_cast<unsigned char *>(ws);
    __half *wt = reinterpret_cast<__half *>(base);
    __half *pl0 = reinterpret_cast<__half *>(base + gs_w_bytes(H));
    __half *pl1 = reinterpret_cast<__half *>(base + gs_w_bytes(H) + gs_plane_bytes(N, H));
    unsigned *cnt = reinterpret_cast<unsigned *>(base + gs_w_bytes(H) + 2 * gs_plane_bytes(N, H));
    {
        const size_t total = (size_t)rows * Kp;
        int blocks = (int)((total + 255) / 256);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        gru_seq_perm_kernel<<<blocks, 256, 0, st>>>(a.w_hh, H, Kp, rows, wt);
        GASR_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    // both planes zero (the K padding stays zero for the whole run), counters zero
    GASR_CUDA(cudaMemsetAsync(pl0, 0, 2 * gs_plane_bytes(N, H) + 1024, st));
    CUtensorMap mh0, mh1, mw;
    GASR_TRY(tc_make_map(&mh0, pl0, N, Kp, TC_BM));
    GASR_TRY(tc_make_map(&mh1, pl1, N, Kp, TC_BM));
    GASR_TRY(tc_make_map(&mw, wt, rows, Kp, GS_BN));
    GruSeqParams p;
    p.T = a.T; p.N = N; p.H = H; p.Kp = Kp; p.KB = Kp / TC_BK; p.reverse = a.reverse; p.nst = gs_stages(H);
    p.xp = a.xproj; p.ldxp = a.ldxp; p.b_hh = a.b_hh; p.out = a.out + a.col0; p.ldo = a.ldo;
    p.plane[0] = pl0; p.plane[1] = pl1; p.cnt = cnt; p.trace = nullptr;
#ifdef GASR_RW_TRACE
    if (getenv("GASR_GS_TRACE")) {
        static unsigned long long *th[2] = {nullptr, nullptr}, *td[2] = {nullptr, nullptr};
        const int sl = a.reverse ? 1 : 0;
        if (!th[sl]) {
            GASR_CUDA(cudaHostAlloc((void **)&th[sl], 64, cudaHostAllocMapped));
            GASR_CUDA(cudaHostGetDevicePointer((void **)&td[sl], th[sl], 0));
        } else {
            GASR_CUDA(cudaStreamSynchronize(st));                         // the previous launch of this direction
            fprintf(stderr, "[gru_seq trace] dir %d, cycles per step of CTA (0,0): wait-peers %.0f ring-free %.0f full-stage %.0f accumulator %.0f math+stores %.0f release %.0f\n",
                    sl, th[sl][0] / (double)a.T, th[sl][1] / (double)a.T, th[sl][2] / (double)a.T, th[sl][3] / (double)a.T, th[sl][4] / (double)a.T, th[sl][5] / (double)a.T);
        }
        for (int i = 0; i < 8; i++) th[sl][i] = 0;
        p.trace = td[sl];
    }
#endif
    const size_t smem = 1024 + (size_t)p.KB * GS_W_TILE + (size_t)p.nst * GS_A_TILE + 256;
    if (!(ctx->attr_mask & 32768u)) {
        GASR_CUDA(cudaFuncSetAttribute(gru_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_MAX_SMEM));
        ctx->attr_mask |= 32768u;
    }
    dim3 grid(ceil_div(H, GS_UNITS), ceil_div(N, TC_BM));
    gru_seq_kernel<<<grid, GS_THREADS, smem, st>>>(mh0, mh1, mw, p);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

}  // namespace gasr
