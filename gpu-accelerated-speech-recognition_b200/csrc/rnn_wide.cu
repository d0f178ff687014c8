// rnn_wide.cu -- the tanh recurrence for MANY utterances per GPU (throughput mode), on the 5th-generation tensor cores.
//
//   h_t = tanh( xp_t + h_{t-1} * W_hh )      xp_t = x_t * W_ih + (b_ih + b_hh), computed for all frames by the projection GEMM
//
// Stands behind RNN::forward / RNN_Cell::forward (reference RNN.cu:15-27, RNN_Cell.cu:65-74: per (t, layer) two cublasSgemm,
// one Sgeam and the Tanh kernel, with a host sync after each).  Rows of the batch never interact (RNN.cu:15-27), so the
// batch is cut into GROUPS OF 128 UTTERANCES = the M dimension of one tcgen05.mma, and a thread-block cluster of H/64 CTAs
// owns one or two groups for a whole time chunk:
//
//   * CTA r of the cluster keeps columns [64r, 64r+64) of W_hh resident in shared memory for the whole launch, as the
//     K-major B operand (W_hh^T slice, bf16 hi + lo planes, 128-byte swizzle): H x 64 x 2 x 2 B = 128 KB at H = 512;
//   * h_{t-1} of the group -- [128 x H] as bf16 hi/lo planes, exactly the array this kernel writes for the next layer's
//     projection GEMM -- is the A operand.  It is pulled through a 2-stage ring by TMA ([128 x 64] boxes); with multicast the
//     H/64 CTAs of the cluster each issue 1/(H/64) of the boxes and every box lands in all of them;
//   * D accumulates in TMEM (fp32); the three terms hi*hi + hi*lo + lo*hi of the fp32-grade split (xproj_gemm_tc.cu) are two
//     instructions per K step: h_hi * [W_hi | W_lo] (N = 128) and h_lo * W_hi (N = 64) -- operand reads from shared memory,
//     not the tensor pipe, pace MMAs this narrow, so the A tile is read twice instead of three times;
//   * eight epilogue warps (thread = utterance row of the accumulator, 32 columns each) add xp, apply tanh, split h_t into
//     its bf16 hi/lo planes and store them (plus, optionally, the fp32 row the C ABI returns), then release the step to the
//     TMA producers of ALL CTAs of the cluster with one remote mbarrier arrive per warp -- no cluster-wide barrier;
//   * with two groups per cluster the chain  TMA -> MMA -> epilogue -> exchange  of one group overlaps the other's.
//
// HBM traffic per frame and utterance: read xp (4H B) + write the planes (4H B) -- the algorithmic bytes of SURVEY.md 8d;
// h_{t-1} comes back from L2, where the cluster has just put it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "common.cuh"
#include "rnn_wide.cuh"
#include "rnn_wide_dev.cuh"
#include "tc_common.cuh"

namespace gasr {

struct RnnWideParams {
    int T, N, Npad, KB;                           // frames, real utterances, plane rows per frame, H / 64 (= cluster size)
    int s0, s1;                                   // steps [s0, s1)
    int n_groups, G;                              // groups of 128 utterances; groups per cluster
    const float *xp; int ldxp, xp_rpf;            // x*W_ih + biases: row t * xp_rpf + n
    float *out; int ldo, col0, out_rpf;           // optional fp32 h_t: row t * out_rpf + n, columns col0 ..
    __nv_bfloat16 *hi, *lo;                       // planes [T * Npad, H]: row t * Npad + n
#ifdef GASR_RW_TRACE
    long long *trace;                             // [steps][16] cycle stamps of cluster 0 / CTA 0 (instrumented build only)
#endif
};

#ifdef GASR_RW_TRACE
#define RW_STAMP(slot) do { if (p.trace && blockIdx.x == 0) p.trace[(size_t)(t - p.s0) * 16 + (slot)] = clock64(); } while (0)
#else
#define RW_STAMP(slot) do { } while (0)
#endif

template <bool MC>
__global__ void __launch_bounds__(RW_THREADS, 1)
rnn_wide_kernel(const __grid_constant__ CUtensorMap map_h_hi, const __grid_constant__ CUtensorMap map_h_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const RnnWideParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t w_smem = (raw + 1023u) & ~1023u;                         // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t ring = w_smem + (uint32_t)p.KB * 2u * RW_W_TILE;
    const uint32_t epi_stage = ring + RW_STAGES * RW_STAGE_BYTES;
    const uint32_t bars = epi_stage + RW_EPI_WARPS * RW_EPI_STAGE;
    const uint32_t full0 = bars, empty0 = bars + 8 * RW_STAGES, accf0 = bars + 16 * RW_STAGES, hrdy0 = accf0 + 16, wfull = hrdy0 + 16;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem_raw + (wfull + 8 - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CL = p.KB;                                                    // cluster size = H / 64
    const uint32_t cr = rw_cluster_rank();
    const int task = blockIdx.x / CL;
    const int g0 = task * p.G;
    const int ng = (p.n_groups - g0) < p.G ? (p.n_groups - g0) : p.G;
    const uint16_t mc_mask = (uint16_t)((1u << CL) - 1u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < RW_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, MC ? CL : 1); }
        for (int g = 0; g < 2; g++) { mbar_init(accf0 + 8 * g, 1); mbar_init(hrdy0 + 8 * g, CL * RW_EPI_WARPS); }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    rw_cluster_sync();                       // every CTA's barriers exist before a peer arrives on them / multicasts into them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(wfull, (uint32_t)p.KB * 2u * RW_W_TILE);            // the resident W_hh^T slice of this CTA
            for (int kb = 0; kb < p.KB; kb++) {
                tma_load_2d(w_smem + (uint32_t)(2 * kb) * RW_W_TILE, &map_w_hi, wfull, kb * TC_BK, (int)cr * RW_COLS);
                tma_load_2d(w_smem + (uint32_t)(2 * kb + 1) * RW_W_TILE, &map_w_lo, wfull, kb * TC_BK, (int)cr * RW_COLS);
            }
            int it = 0;
            for (int t = p.s0; t < p.s1; t++) {
                if (t == 0) continue;                                        // h_{-1} = 0: nothing to multiply
                for (int g = 0; g < ng; g++) {
                    const int i = t - p.s0;
                    if (g == 0) RW_STAMP(0);
                    if (i >= 1) rw_wait_cluster(hrdy0 + 8 * g, (uint32_t)(i - 1) & 1u);   // h_{t-1} stored by every CTA of the cluster
                    if (g == 0) RW_STAMP(1);
                    asm volatile("fence.proxy.async;" ::: "memory");         // generic-proxy stores -> async-proxy (TMA) reads
                    const int row = (t - 1) * p.Npad + (g0 + g) * RW_U;
                    for (int kb = 0; kb < p.KB; kb++, it++) {
                        const int s = it % RW_STAGES;
                        rw_wait(empty0 + 8 * s, ((uint32_t)(it / RW_STAGES) & 1u) ^ 1u);
                        const uint32_t st = ring + (uint32_t)s * RW_STAGE_BYTES;
                        mbar_expect_tx(full0 + 8 * s, RW_STAGE_BYTES);
                        if (MC) {
                            if ((uint32_t)(kb % CL) == cr) {
                                rw_tma_load_mc(st, &map_h_hi, full0 + 8 * s, kb * TC_BK, row, mc_mask);
                                rw_tma_load_mc(st + RW_A_TILE, &map_h_lo, full0 + 8 * s, kb * TC_BK, row, mc_mask);
                            }
                        } else {
                            tma_load_2d(st, &map_h_hi, full0 + 8 * s, kb * TC_BK, row);
                            tma_load_2d(st + RW_A_TILE, &map_h_lo, full0 + 8 * s, kb * TC_BK, row);
                        }
                    }
                    if (g == 0) RW_STAMP(2);
                }
            }
            if (MC) {
                // peers' commits still arrive on this CTA's empty barriers: drain them before the CTA may exit
                for (int j = 0; j < RW_STAGES; j++, it++) rw_wait(empty0 + 8 * (it % RW_STAGES), ((uint32_t)(it / RW_STAGES) & 1u) ^ 1u);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptors: D = f32, A = B = bf16, both K-major, M = 128; N = 128 ([W_hi ; W_lo] rows) and N = 64
            constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * RW_COLS >> 3) << 17) | ((uint32_t)(RW_U >> 4) << 24);
            constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RW_COLS >> 3) << 17) | ((uint32_t)(RW_U >> 4) << 24);
            rw_wait(wfull, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int it = 0;
            for (int t = p.s0; t < p.s1; t++) {
                if (t == 0) continue;
                for (int g = 0; g < ng; g++) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(g * 2 * RW_COLS);
                    for (int kb = 0; kb < p.KB; kb++, it++) {
                        const int s = it % RW_STAGES;
                        rw_wait(full0 + 8 * s, (uint32_t)(it / RW_STAGES) & 1u);
                        if (g == 0 && kb == 0) RW_STAMP(3);
                        if (g == 0 && kb == p.KB - 1) RW_STAMP(4);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = ring + (uint32_t)s * RW_STAGE_BYTES;
                        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + RW_A_TILE);
                        // The hi and lo tiles of a k-block are adjacent: together they are ONE [128 x 64] B tile [W_hi ; W_lo].
                        // h_hi * [W_hi | W_lo] is a single N = 128 instruction (columns 0-63: hi*hi, 64-127: hi*lo, summed in
                        // the epilogue) and h_lo * W_hi accumulates onto columns 0-63: the A tile is read twice instead of
                        // three times -- with N = 64 the operand reads from shared memory, not the tensor pipe, set the pace.
                        const uint64_t b_hl = umma_desc_sw128(w_smem + (uint32_t)(2 * kb) * RW_W_TILE);
#pragma unroll
                        for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                            const uint64_t adv = (uint64_t)(k4 * 32 >> 4);   // 16 bf16 = 32 bytes along K inside the swizzle atom
                            umma_bf16(d_tmem, a_hi + adv, b_hl + adv, idesc2, (kb | k4) != 0);
                            umma_bf16(d_tmem, a_lo + adv, b_hl + adv, idesc1, 1);
                        }
                        if (MC) rw_commit_mc(empty0 + 8 * s, mc_mask);       // the stage is free once EVERY CTA's MMAs have read it
                        else umma_commit(empty0 + 8 * s);
                    }
                    umma_commit(accf0 + 8 * g);
                    if (g == 0) RW_STAMP(5);
                }
            }
        }
    } else {
        // ===== epilogue: thread = utterance row of the accumulator; warps 2-5 columns 0-31, warps 6-9 columns 32-63 =====
        const int q = warp & 3;                                              // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int c0 = (int)cr * RW_COLS + half * 32;                        // first hidden unit of this thread
        const int H = p.KB * 64;
        const int t_first = p.s0 == 0 ? 1 : p.s0;                            // first step with an accumulator
        for (int t = p.s0; t < p.s1; t++) {
            for (int g = 0; g < ng; g++) {
                const int u = (g0 + g) * RW_U + row;
                const bool live = u < p.N;
                // The projection values do not depend on the MMAs: request them before waiting for the accumulator.  Coalesced:
                // an instruction fetches 4 rows x 128 B of the warp's [32 rows x 32 columns] piece (a row-per-thread load
                // touches 32 cache lines per instruction); the piece is transposed through shared memory below.
                const int ubase = (g0 + g) * RW_U + q * 32;
                const int lr = lane >> 3, lc = lane & 7;
                float4 xg[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int uu = ubase + 4 * i + lr;
                    xg[i] = uu < p.N ? __ldcs(reinterpret_cast<const float4 *>(p.xp + ((size_t)t * p.xp_rpf + uu) * p.ldxp + c0) + lc)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                uint32_t v[32];
                const bool stamp = g == 0 && warp == 2 && lane == 0;
                if (stamp) RW_STAMP(6);
                if (t > 0) {
                    rw_wait(accf0 + 8 * g, (uint32_t)(t - t_first) & 1u);
                    if (stamp) RW_STAMP(7);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t v2[32];
                    rw_tmem_ld32(v, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 2 * RW_COLS + half * 32));
                    rw_tmem_ld32(v2, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 2 * RW_COLS + RW_COLS + half * 32));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the next step's MMAs overwrite the accumulator
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = 0u;
                }
                if (stamp) RW_STAMP(8);
                float4 x4[8];
                {
                    float4 *sx = reinterpret_cast<float4 *>(smem_raw + (epi_stage - raw) + (uint32_t)(warp - 2) * RW_EPI_STAGE);
#pragma unroll
                    for (int i = 0; i < 8; i++) { const int r = 4 * i + lr; sx[r * 8 + (lc ^ (r & 7))] = xg[i]; }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; j++) x4[j] = sx[lane * 8 + (j ^ (lane & 7))];
                    __syncwarp();
                }
                float h[32];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    h[4 * j + 0] = rw_tanh(__uint_as_float(v[4 * j + 0]) + x4[j].x);
                    h[4 * j + 1] = rw_tanh(__uint_as_float(v[4 * j + 1]) + x4[j].y);
                    h[4 * j + 2] = rw_tanh(__uint_as_float(v[4 * j + 2]) + x4[j].z);
                    h[4 * j + 3] = rw_tanh(__uint_as_float(v[4 * j + 3]) + x4[j].w);
                }
                // bf16 hi / lo planes of h_t: the next step's A operand and the next layer's GEMM operand
                uint32_t ph[16], pl[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {                               // packed conversions: one cvt.rn.bf16x2.f32 per pair
                    const __nv_bfloat162 th = __floats2bfloat162_rn(h[2 * j], h[2 * j + 1]);
                    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&th);
                    const float f0 = __uint_as_float(hb << 16), f1 = __uint_as_float(hb & 0xffff0000u);
                    const __nv_bfloat162 tl = __floats2bfloat162_rn(h[2 * j] - f0, h[2 * j + 1] - f1);
                    ph[j] = hb;
                    pl[j] = *reinterpret_cast<const uint32_t *>(&tl);
                }
                // Stores: a thread-per-row store touches 32 cache lines per instruction (the epilogue was bound by exactly that).
                // The warp's [32 rows x 64 B] pieces of both planes go through a swizzled shared-memory tile and leave as
                // 8 rows x 64 contiguous bytes per instruction.
                {
                    uint4 *sw = reinterpret_cast<uint4 *>(smem_raw + (epi_stage - raw) + (uint32_t)(warp - 2) * RW_EPI_STAGE);
                    const int xr = (lane >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        sw[lane * 4 + (j ^ xr)] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
                        sw[128 + lane * 4 + (j ^ xr)] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i2 = 0; i2 < 4; i2++) {
                        const int r = 8 * i2 + (lane >> 2), c = lane & 3;
                        const int sidx = r * 4 + (c ^ ((r >> 1) & 3));
                        const size_t prow = ((size_t)t * p.Npad + ubase + r) * H + c0;
                        reinterpret_cast<uint4 *>(p.hi + prow)[c] = sw[sidx];
                        reinterpret_cast<uint4 *>(p.lo + prow)[c] = sw[128 + sidx];
                    }
                    __syncwarp();
                }
                if (p.out != nullptr && live) {
                    float4 *o = reinterpret_cast<float4 *>(p.out + ((size_t)t * p.out_rpf + u) * p.ldo + p.col0 + c0);
#pragma unroll
                    for (int j = 0; j < 8; j++) __stcs(o + j, make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
                }
                // publish: this warp's rows of h_t are stored -> ONE cluster-scope release fence, then a relaxed arrive on every
                // CTA's "h ready" barrier (a release per arrive would drain the stores once per destination)
                if (stamp) RW_STAMP(9);
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.acq_rel.cluster;" ::: "memory");
                    if (stamp) RW_STAMP(10);
                    for (int r = 0; r < CL; r++) rw_arrive_remote(hrdy0 + 8 * g, (uint32_t)r);
                    if (stamp) RW_STAMP(11);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    rw_cluster_sync();                       // no CTA leaves while a peer may still arrive on its barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------

bool rnn_wide_supported(const gasr_ctx *ctx, int H) {
    return ctx->cluster_ok && (H == 64 || H == 128 || H == 256 || H == 512);
}

size_t rnn_wide_smem_bytes(int H) { return 1024 + (size_t)(H / 64) * 2 * RW_W_TILE + RW_STAGES * RW_STAGE_BYTES + RW_EPI_WARPS * RW_EPI_STAGE + 256; }
size_t rnn_wide_plane_bytes(int T, int Npad, int H) { return align_up((size_t)T * Npad * H * 2, 1024); }   // one plane (hi or lo)

// W_hh -> W_hh^T bf16 hi/lo planes (xproj_tc_prepare_weights layout) + the four TMA descriptors of a layer
int rnn_wide_plan(gasr_ctx *ctx, RnnWidePlan &pl, const float *w_hh, int T, int N, int H, void *wbuf, void *planes, cudaStream_t st,
                  int pad_to) {
    GASR_CHECK(rnn_wide_supported(ctx, H), "rnn_wide: hidden size %d unsupported", H);
    GASR_CHECK(pad_to == 128 || pad_to == 256, "rnn_wide: groups are 128 or 256 utterances");
    pl.T = T; pl.N = N; pl.H = H; pl.Npad = ceil_div(N, pad_to) * pad_to;
    GASR_TRY(xproj_tc_prepare_weights(ctx, w_hh, H, H, wbuf, st));
    unsigned char *wb = static_cast<unsigned char *>(wbuf), *pb = static_cast<unsigned char *>(planes);
    pl.hi = reinterpret_cast<__nv_bfloat16 *>(pb);
    pl.lo = reinterpret_cast<__nv_bfloat16 *>(pb + rnn_wide_plane_bytes(T, pl.Npad, H));
    GASR_TRY(tc_make_map(&pl.maps[0], pl.hi, T * pl.Npad, H, RW_U));
    GASR_TRY(tc_make_map(&pl.maps[1], pl.lo, T * pl.Npad, H, RW_U));
    GASR_TRY(tc_make_map(&pl.maps[2], wb, H, H, RW_COLS));
    GASR_TRY(tc_make_map(&pl.maps[3], wb + xproj_tc_w_bytes(H, H) / 2, H, H, RW_COLS));
    return GASR_OK;
}

// Function attributes are set once per context and never while other kernels of a pipeline may be running.
int rnn_wide_prepare(gasr_ctx *ctx, int H) {
    if (ctx->attr_mask & 4096u) return GASR_OK;
    const int smem = (int)rnn_wide_smem_bytes(512);
    GASR_CUDA(cudaFuncSetAttribute(rnn_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    GASR_CUDA(cudaFuncSetAttribute(rnn_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ctx->attr_mask |= 4096u;
    (void)H;
    return GASR_OK;
}

int launch_rnn_wide(gasr_ctx *ctx, const RnnWidePlan &pl, const RnnWideRun &r, cudaStream_t st) {
    GASR_CHECK(r.xp && r.s0 >= 0 && r.s0 < r.s1 && r.s1 <= pl.T, "rnn_wide: bad step range [%d, %d)", r.s0, r.s1);
    GASR_CHECK(r.ldxp % 4 == 0 && (reinterpret_cast<uintptr_t>(r.xp) & 15) == 0, "rnn_wide: xproj must be 16-byte aligned");
    GASR_CHECK(r.out == nullptr || (r.ldo % 4 == 0 && r.col0 % 4 == 0 && (reinterpret_cast<uintptr_t>(r.out) & 15) == 0),
               "rnn_wide: output must be 16-byte aligned");
    GASR_TRY(rnn_wide_prepare(ctx, pl.H));
    RnnWideParams p;
    p.T = pl.T; p.N = pl.N; p.Npad = pl.Npad; p.KB = pl.H / 64; p.s0 = r.s0; p.s1 = r.s1;
    p.n_groups = pl.Npad / RW_U;
    p.G = r.groups_per_cluster >= 2 ? 2 : 1;
    p.xp = r.xp; p.ldxp = r.ldxp; p.xp_rpf = r.xp_rows_per_frame;
    p.out = r.out; p.ldo = r.ldo; p.col0 = r.col0; p.out_rpf = r.out_rows_per_frame;
    p.hi = pl.hi; p.lo = pl.lo;
    const int tasks = ceil_div(p.n_groups, p.G);
#ifdef GASR_RW_TRACE
    const int steps = r.s1 - r.s0;
    p.trace = nullptr;
    GASR_CUDA(cudaMalloc(&p.trace, sizeof(long long) * 16 * (size_t)steps));
    GASR_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(long long) * 16 * (size_t)steps, st));
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tasks * p.KB);
    cfg.blockDim = dim3(RW_THREADS);
    cfg.dynamicSmemBytes = rnn_wide_smem_bytes(pl.H);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.KB; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (r.multicast && p.KB > 1)
        GASR_CUDA(cudaLaunchKernelEx(&cfg, rnn_wide_kernel<true>, pl.maps[0], pl.maps[1], pl.maps[2], pl.maps[3], p));
    else
        GASR_CUDA(cudaLaunchKernelEx(&cfg, rnn_wide_kernel<false>, pl.maps[0], pl.maps[1], pl.maps[2], pl.maps[3], p));
    ctx->launches += 1;
#ifdef GASR_RW_TRACE
    {   // instrumented build: mean cycles between the stamps of cluster 0 / CTA 0 over the steps of this launch
        std::vector<long long> h((size_t)16 * steps);
        GASR_CUDA(cudaStreamSynchronize(st));
        GASR_CUDA(cudaMemcpy(h.data(), p.trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        cudaFree(p.trace);
        static const char *names[12] = {"P:top", "P:hrdy", "P:issued", "M:full0", "M:fullK", "M:commit", "E:top", "E:accf", "E:ld",
                                        "E:stored", "E:fenced", "E:arrived"};
        double sum[12] = {0}; int cnt = 0;
        for (int i = 2; i + 1 < steps; i++) {
            const long long *x = &h[(size_t)16 * i];
            if (!x[0] || !x[11]) continue;
            for (int k = 0; k < 12; k++) sum[k] += (double)(x[k] - x[0]);
            cnt++;
        }
        double step = 0;
        if (steps > 4) step = (double)(h[(size_t)16 * (steps - 2)] - h[(size_t)16 * 2]) / (steps - 4);
        fprintf(stderr, "[rw trace] tasks %d groups/cluster %d mc %d: %.0f cycles per step; offsets from P:top:", tasks, p.G, (int)(r.multicast && p.KB > 1), step);
        for (int k = 0; k < 12; k++) fprintf(stderr, " %s %.0f", names[k], cnt ? sum[k] / cnt : 0.0);
        fprintf(stderr, "\n");
    }
#endif
    return GASR_OK;
}

}  // namespace gasr
