// rnn_wide2.cu -- the wide tanh recurrence with CTA PAIRS: tcgen05.mma.cta_group::2, groups of 256 utterances.
//
// Same contract as rnn_wide.cu (h_t = tanh(xp_t + h_{t-1} * W_hh), reference RNN.cu:15-27 / RNN_Cell.cu:65-74; planes in,
// planes out), different shape.  rnn_wide.cu is bound by what an SM can ingest: every CTA of the cluster pulls the whole
// [128 x H] h_{t-1} of its group (256 KB per step as bf16 hi/lo planes at ~36 B/clk) to produce 128 x 64 outputs.  Here the
// two CTAs of a pair (cluster ranks 2p, 2p+1 -- one TPC) execute ONE M = 256 instruction: each CTA supplies the h rows of
// ITS 128 utterances (A) and ITS 64 columns of W_hh^T (half of B, N = 128), and receives the accumulator rows of its
// utterances for the 128 columns of the pair.  Per ingested byte a CTA now produces twice the outputs; the W_hh slice per
// CTA (resident, 128 KB), the h exchange through L2 + TMA and the per-warp release/acquire of a step are unchanged.
//
//   cluster = H/64 CTAs = H/128 pairs; a cluster owns one or two groups of 256 utterances (ping-pong) for a time chunk
//   leader (even rank): issues the MMAs (three per K step: hi*hi, hi*lo, lo*hi into the same 128 TMEM columns) and commits
//                       with a multicast arrive to both CTAs' barriers
//   both CTAs:          TMA producer for their own A tiles (the odd CTA's loads complete on the LEADER's "full" barrier),
//                       eight epilogue warps for their own 128 accumulator rows x 128 columns (two passes of 32 columns per warp)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "rnn_wide.cuh"
#include "rnn_wide_dev.cuh"
#include "tc_common.cuh"

namespace gasr {

struct RnnWide2Params {
    int T, N, Npad, KB;                           // frames, real utterances, plane rows per frame (multiple of 256), H / 64
    int s0, s1;
    int n_groups, G;                              // groups of 256 utterances; groups per cluster (1 or 2)
    const float *xp; int ldxp, xp_rpf;
    float *out; int ldo, col0, out_rpf;
    __nv_bfloat16 *hi, *lo;
    volatile unsigned *tlog;                      // instrumented build only
};

__global__ void __launch_bounds__(RW_THREADS, 1)
rnn_wide2_kernel(const __grid_constant__ CUtensorMap map_h_hi, const __grid_constant__ CUtensorMap map_h_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const RnnWide2Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t w_smem = (raw + 1023u) & ~1023u;
    const uint32_t ring = w_smem + (uint32_t)p.KB * 2u * RW_W_TILE;
    const uint32_t epi_stage = ring + RW_STAGES * RW_STAGE_BYTES;
    const uint32_t bars = epi_stage + RW_EPI_WARPS * RW_EPI_STAGE;
    const uint32_t full0 = bars, empty0 = bars + 8 * RW_STAGES, accf0 = bars + 16 * RW_STAGES, hrdy0 = accf0 + 16, wfull = hrdy0 + 16;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem_raw + (wfull + 8 - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CL = p.KB;                                                    // cluster size = H / 64 (even)
    const uint32_t cr = rw_cluster_rank();
    const uint32_t e = cr & 1u;                                             // which 128 utterances of a group are this CTA's
    const uint32_t leader = cr & ~1u;
    const uint16_t pair_mask = (uint16_t)(3u << leader);
    const int task = blockIdx.x / CL;
    const int g0 = task * p.G;
    const int ng = (p.n_groups - g0) < p.G ? (p.n_groups - g0) : p.G;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RW_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int g = 0; g < 2; g++) { mbar_init(accf0 + 8 * g, 1); mbar_init(hrdy0 + 8 * g, CL * RW_EPI_WARPS); }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    rw_cluster_sync();                                                      // the peer is running before the pair-collective alloc (see gemm_pair.cu)
    if (warp == 1) {                                                        // pair-collective: one warp of EACH CTA
        GASR_TLOG(p.tlog, 2, 1);
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(256) : "memory");
        GASR_TLOG(p.tlog, 2, 2);
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0 && lane == 0) {                                           // the resident W_hh^T slice of this CTA
        mbar_expect_tx(wfull, (uint32_t)p.KB * 2u * RW_W_TILE);
        for (int kb = 0; kb < p.KB; kb++) {
            tma_load_2d(w_smem + (uint32_t)(2 * kb) * RW_W_TILE, &map_w_hi, wfull, kb * TC_BK, (int)cr * RW_COLS);
            tma_load_2d(w_smem + (uint32_t)(2 * kb + 1) * RW_W_TILE, &map_w_lo, wfull, kb * TC_BK, (int)cr * RW_COLS);
        }
    }
    rw_wait(wfull, 0);                       // every thread: this CTA's slice is resident ...
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    rw_cluster_sync();                       // ... and so is the peer's (the leader's MMAs read both), and all barriers exist
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t own_base = *tmem_slot;    // what this CTA's warp allocated (and frees)
    uint32_t tmem_base;                      // the accumulator address the leader's MMAs write in BOTH CTAs
    {
        uint32_t leader_slot;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_slot) : "r"(smem_u32((const void *)tmem_slot)), "r"(leader));
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(tmem_base) : "r"(leader_slot) : "memory");
    }

    if (warp == 0) {
        // ===== TMA producer: this CTA's 128 rows of h_{t-1}; completion always on the leader's "full" barrier =====
        if (lane == 0) {
            uint32_t leader_full0;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_full0) : "r"(full0), "r"(leader));
            int it = 0;
            for (int t = p.s0; t < p.s1; t++) {
                if (t == 0) continue;                                        // h_{-1} = 0: nothing to multiply
                for (int g = 0; g < ng; g++) {
                    const int i = t - p.s0;
                    if (i >= 1) rw_wait_cluster(hrdy0 + 8 * g, (uint32_t)(i - 1) & 1u);
                    asm volatile("fence.proxy.async;" ::: "memory");
                    const int row = (t - 1) * p.Npad + (g0 + g) * 2 * RW_U + (int)e * RW_U;
                    for (int kb = 0; kb < p.KB; kb++, it++) {
                        const int s = it % RW_STAGES;
                        rw_wait(empty0 + 8 * s, ((uint32_t)(it / RW_STAGES) & 1u) ^ 1u);
                        const uint32_t st = ring + (uint32_t)s * RW_STAGE_BYTES;
                        if (e == 0) mbar_expect_tx(full0 + 8 * s, 2 * RW_STAGE_BYTES);      // both CTAs' tiles
                        rw2_tma_load_to_leader(st, &map_h_hi, leader_full0 + 8 * s, kb * TC_BK, row);
                        rw2_tma_load_to_leader(st + RW_A_TILE, &map_h_lo, leader_full0 + 8 * s, kb * TC_BK, row);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader of the pair only =====
        if (lane == 0 && e == 0) {
            // D = f32, A = B = bf16, both K-major, M = 256 (128 rows per CTA), N = 128 (64 W_hh columns per CTA)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * RW_COLS >> 3) << 17) | ((uint32_t)(2 * RW_U >> 4) << 24);
            int it = 0;
            for (int t = p.s0; t < p.s1; t++) {
                if (t == 0) continue;
                for (int g = 0; g < ng; g++) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(g * 2 * RW_COLS);
                    for (int kb = 0; kb < p.KB; kb++, it++) {
                        const int s = it % RW_STAGES;
                        rw_wait(full0 + 8 * s, (uint32_t)(it / RW_STAGES) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = ring + (uint32_t)s * RW_STAGE_BYTES;
                        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + RW_A_TILE);
                        const uint64_t b_hi = umma_desc_sw128(w_smem + (uint32_t)(2 * kb) * RW_W_TILE);
                        const uint64_t b_lo = umma_desc_sw128(w_smem + (uint32_t)(2 * kb + 1) * RW_W_TILE);
#pragma unroll
                        for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                            const uint64_t adv = (uint64_t)(k4 * 32 >> 4);
                            rw2_umma(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k4) != 0);
                            rw2_umma(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                            rw2_umma(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
                        }
                        rw2_commit_pair(empty0 + 8 * s, pair_mask);          // both CTAs' stage s is free once these MMAs retire
                    }
                    rw2_commit_pair(accf0 + 8 * g, pair_mask);
                }
            }
        }
    } else {
        // ===== epilogue: thread = utterance row of this CTA's accumulator half; 64 of the pair's 128 columns per warp, in two
        //       passes of 32 (warps 2-5: columns 0-63, warps 6-9: columns 64-127) =====
        const int q = warp & 3;
        const int hsel = (warp - 2) >> 2;
        const int H = p.KB * 64;
        const int t_first = p.s0 == 0 ? 1 : p.s0;
        const int lr = lane >> 3, lc = lane & 7;
        float4 *sx = reinterpret_cast<float4 *>(smem_raw + (epi_stage - raw) + (uint32_t)(warp - 2) * RW_EPI_STAGE);
        uint4 *sw = reinterpret_cast<uint4 *>(sx);
        for (int t = p.s0; t < p.s1; t++) {
            for (int g = 0; g < ng; g++) {
                const int ubase = (g0 + g) * 2 * RW_U + (int)e * RW_U + q * 32;      // first utterance of this warp's 32 rows
                const int u = ubase + lane;
                const bool live = u < p.N;
                // both passes' projection values are requested before the accumulator wait (coalesced: 4 rows x 128 B per instruction)
                float4 xg[2][8];
#pragma unroll
                for (int ps = 0; ps < 2; ps++) {
                    const int c0 = (int)leader * RW_COLS + hsel * 64 + ps * 32;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int uu = ubase + 4 * i + lr;
                        xg[ps][i] = uu < p.N ? __ldcs(reinterpret_cast<const float4 *>(p.xp + ((size_t)t * p.xp_rpf + uu) * p.ldxp + c0) + lc)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (t > 0) {
                    rw_wait(accf0 + 8 * g, (uint32_t)(t - t_first) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
#pragma unroll
                for (int ps = 0; ps < 2; ps++) {
                    const int c0 = (int)leader * RW_COLS + hsel * 64 + ps * 32;      // first hidden unit of this pass
                    uint32_t v[32];
                    if (t > 0) {
                        rw_tmem_ld32(v, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 2 * RW_COLS + hsel * 64 + ps * 32));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (ps == 1) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the next step's MMAs overwrite the accumulator
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = 0u;
                    }
                    float4 x4[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) { const int r = 4 * i + lr; sx[r * 8 + (lc ^ (r & 7))] = xg[ps][i]; }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; j++) x4[j] = sx[lane * 8 + (j ^ (lane & 7))];
                    __syncwarp();
                    float h[32];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        h[4 * j + 0] = rw_tanh(__uint_as_float(v[4 * j + 0]) + x4[j].x);
                        h[4 * j + 1] = rw_tanh(__uint_as_float(v[4 * j + 1]) + x4[j].y);
                        h[4 * j + 2] = rw_tanh(__uint_as_float(v[4 * j + 2]) + x4[j].z);
                        h[4 * j + 3] = rw_tanh(__uint_as_float(v[4 * j + 3]) + x4[j].w);
                    }
                    uint32_t ph[16], pl[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const __nv_bfloat162 th = __floats2bfloat162_rn(h[2 * j], h[2 * j + 1]);
                        const uint32_t hb = *reinterpret_cast<const uint32_t *>(&th);
                        const float f0 = __uint_as_float(hb << 16), f1 = __uint_as_float(hb & 0xffff0000u);
                        const __nv_bfloat162 tl = __floats2bfloat162_rn(h[2 * j] - f0, h[2 * j + 1] - f1);
                        ph[j] = hb;
                        pl[j] = *reinterpret_cast<const uint32_t *>(&tl);
                    }
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
                    if (p.out != nullptr && live) {
                        float4 *o = reinterpret_cast<float4 *>(p.out + ((size_t)t * p.out_rpf + u) * p.ldo + p.col0 + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) __stcs(o + j, make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
                    }
                }
                // publish: ONE cluster-scope release fence per warp, then a relaxed arrive on every CTA's "h ready" barrier
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.acq_rel.cluster;" ::: "memory");
                    for (int r = 0; r < CL; r++) rw_arrive_remote(hrdy0 + 8 * g, (uint32_t)r);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    rw_cluster_sync();                       // no CTA leaves while a peer may still arrive on its barriers / read its operands
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        GASR_TLOG(p.tlog, 2, 3);
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(own_base), "n"(256) : "memory");
        GASR_TLOG(p.tlog, 2, 4);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------

bool rnn_wide2_supported(const gasr_ctx *ctx, int H) { return ctx->cluster_ok && (H == 128 || H == 256 || H == 512); }

int rnn_wide2_prepare(gasr_ctx *ctx) {
    if (ctx->attr_mask & 8192u) return GASR_OK;
    GASR_CUDA(cudaFuncSetAttribute(rnn_wide2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rnn_wide_smem_bytes(512)));
    ctx->attr_mask |= 8192u;
    return GASR_OK;
}

// Same plan as the one-CTA kernel (the planes must have Npad % 256 == 0).
int launch_rnn_wide2(gasr_ctx *ctx, const RnnWidePlan &pl, const RnnWideRun &r, cudaStream_t st) {
    GASR_CHECK(rnn_wide2_supported(ctx, pl.H) && pl.Npad % 256 == 0, "rnn_wide2: needs H in {128, 256, 512} and whole groups of 256 utterances");
    GASR_CHECK(r.xp && r.s0 >= 0 && r.s0 < r.s1 && r.s1 <= pl.T, "rnn_wide2: bad step range [%d, %d)", r.s0, r.s1);
    GASR_CHECK(r.ldxp % 4 == 0 && (reinterpret_cast<uintptr_t>(r.xp) & 15) == 0, "rnn_wide2: xproj must be 16-byte aligned");
    GASR_CHECK(r.out == nullptr || (r.ldo % 4 == 0 && r.col0 % 4 == 0 && (reinterpret_cast<uintptr_t>(r.out) & 15) == 0),
               "rnn_wide2: output must be 16-byte aligned");
    GASR_TRY(rnn_wide2_prepare(ctx));
    RnnWide2Params p;
    p.tlog = nullptr;
#ifdef GASR_RW_TRACE
    p.tlog = trace_tmem_log();
#endif
    p.T = pl.T; p.N = pl.N; p.Npad = pl.Npad; p.KB = pl.H / 64; p.s0 = r.s0; p.s1 = r.s1;
    p.n_groups = pl.Npad / 256;
    p.G = r.groups_per_cluster >= 2 ? 2 : 1;
    p.xp = r.xp; p.ldxp = r.ldxp; p.xp_rpf = r.xp_rows_per_frame;
    p.out = r.out; p.ldo = r.ldo; p.col0 = r.col0; p.out_rpf = r.out_rows_per_frame;
    p.hi = pl.hi; p.lo = pl.lo;
    const int tasks = ceil_div(p.n_groups, p.G);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tasks * p.KB);
    cfg.blockDim = dim3(RW_THREADS);
    cfg.dynamicSmemBytes = rnn_wide_smem_bytes(pl.H);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.KB; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GASR_CUDA(cudaLaunchKernelEx(&cfg, rnn_wide2_kernel, pl.maps[0], pl.maps[1], pl.maps[2], pl.maps[3], p));
    ctx->launches += 1;
    return GASR_OK;
}

}  // namespace gasr
