// gemm_pair.cu -- the wave engine's projection GEMM on CTA PAIRS:  C[M, H] = A[M, K] * W[K, H] + bias,  A and W^T as bf16
// hi/lo planes, tcgen05.mma.cta_group::2 with 256 x 256 tiles.
//
// Stands behind the x * W_ih of RNN_Cell::forward (reference RNN_Cell.cu:66 via cuMatrix.cpp:46-60), batched over all frames
// of a time chunk.  The one-CTA tile engine (xproj_stream.cu) is bound by what an SM can ingest: a 128 x 128 tile of the
// fp32-grade 3-term product reads 512 KB of operands for 6144 cycles of MMA, twice what the ~36-40 B/clk per SM deliver.
// A pair of CTAs (one TPC) shares the B operand: each CTA loads ITS 128 rows of A and ITS 128 rows of W^T (half of N = 256),
// the leader issues M = 256, N = 256 instructions, each CTA receives the accumulator rows of its own 128 rows of A for all
// 256 columns -- the same 512 KB per CTA now feed 12 288 cycles of MMA.  With two 64 KB stages the kernel also leaves
// ~100 KB of shared memory per SM, so decoder CTAs (whose issue slots this kernel barely touches) run on the same SMs.
//
//   warp 0    TMA producer (both CTAs): A [128 x 64] hi/lo + W^T [128 x 64] hi/lo per K block; every load completes on the
//             LEADER's "full" barrier;
//   warp 1    MMA issuer (leader only): 3 x 4 instructions per K block, accumulator double buffered in TMEM (2 x 256 columns),
//             commits multicast to both CTAs;
//   warps 2-5 epilogue (both CTAs): tcgen05.ld -> + bias -> shared staging -> coalesced 128-byte row segments of C.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm_pair.cuh"
#include "rnn_wide_dev.cuh"
#include "tc_common.cuh"

namespace gasr {

constexpr int GP_STAGES = 2;
constexpr int GP_STAGE_BYTES = 4 * TC_TILE_BYTES;             // A hi, A lo, B hi, B lo: [128 x 64] bf16 each
constexpr int GP_THREADS = 192;
constexpr int GP_EPI_STAGE = 4 * 32 * 36 * 4;                 // [4 warps][32 rows][36 floats]
constexpr int GP_SMEM_BYTES = 1024 + GP_STAGES * GP_STAGE_BYTES + 256 + GP_EPI_STAGE;

struct GemmPairParams {
    int M, row0;                  // rows of this launch (multiple of 256), first row inside the A descriptor
    int n_ct, kblocks, terms;     // column tiles of 256, K / 64, 3 (fp32-grade) or 1 (bf16)
    float *C; int ldc;            // C already points at row row0
    const float *bias;
    volatile unsigned *trace;     // instrumented build only (make TRACE=1): progress words in mapped host memory
    volatile unsigned *tlog;      // instrumented build only: per-SM TMEM allocator event ring
    int variant;                  // instrumented build only: where the allocation permit is given up (0 = after the alloc, as in the product)
};

#ifdef GASR_RW_TRACE
// [slot][CTA][8]: word 0 setup (1 entered, 2 TMEM allocated, 3 pair synchronised, 9 exited), 1 producer (4 waits for an empty
// stage, 5 issued), 2 MMA issuer (6 waits for a drained accumulator, 7 waits for a full stage, 8 committed a tile),
// 3-6 epilogue warps (10 waits for the accumulator, 11 stored a tile); low 24 bits = iteration / tile counter
#define GP_MARK(role, code, it) do { if (p.trace) p.trace[blockIdx.x * 8 + (role)] = ((unsigned)(code) << 24) | ((unsigned)(it) & 0xffffffu); } while (0)
#else
#define GP_MARK(role, code, it) do { } while (0)
#endif

__device__ __forceinline__ void gp_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

__global__ void __launch_bounds__(GP_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const GemmPairParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles + GP_STAGES * GP_STAGE_BYTES;          // full[S], empty[S], tfull[2], tempty[2], tmem slot
    const uint32_t full0 = bars, empty0 = bars + 8 * GP_STAGES, tfull0 = bars + 16 * GP_STAGES, tempty0 = tfull0 + 16;
    unsigned char *gen = smem_raw + (tiles - raw);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + GP_STAGES * GP_STAGE_BYTES + 16 * GP_STAGES + 32);
    float *epi_stage = reinterpret_cast<float *>(gen + GP_STAGES * GP_STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t e = rw_cluster_rank();                              // cluster of two: 0 = leader
    const uint16_t pair_mask = 3;
    const int n_pairs = gridDim.x / 2, pair = blockIdx.x / 2;
    const int n_tiles = (p.M / 256) * p.n_ct;

    if (threadIdx.x == 0) {
        GP_MARK(0, 1, 0);
        for (int s = 0; s < GP_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8); }   // 4 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Both CTAs of the pair must be running before either issues the pair-collective tcgen05.alloc (as CUTLASS does: cluster
    // sync first).  A CTA that allocated while its peer was still being launched left the peer's own alloc blocked for good --
    // an intermittent stall of the whole pipeline (found with the progress words of `make TRACE=1`).
    __syncthreads();
    rw_cluster_sync();
    if (warp == 1) {                                                   // pair-collective allocation: one warp of each CTA
#ifdef GASR_RW_TRACE
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (lane == 0) GP_MARK(7, 20, smid);
#endif
        GASR_TLOG(p.tlog, 1, 1);
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(512) : "memory");
        GASR_TLOG(p.tlog, 1, 2);
#ifdef GASR_RW_TRACE
        if (lane == 0) GP_MARK(7, 21, smid);
        if (p.variant == 0) {
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
            if (lane == 0) GP_MARK(7, 22, smid);
        }
#else
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) GP_MARK(0, 2, 0);
    rw_cluster_sync();
    if (threadIdx.x == 0) GP_MARK(0, 3, 0);
#ifdef GASR_RW_TRACE
    if (warp == 1 && p.variant == 3) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
#endif
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t own_base = *tmem_slot;
    uint32_t tmem_base;
    {
        uint32_t leader_slot;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_slot) : "r"(smem_u32((const void *)tmem_slot)), "r"(0u));
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(tmem_base) : "r"(leader_slot) : "memory");
    }

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t leader_full0;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_full0) : "r"(full0), "r"(0u));
            const uint32_t per_cta = (p.terms == 3 ? 4u : 2u) * TC_TILE_BYTES;
            int it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int rb = tile / p.n_ct, ct = tile - rb * p.n_ct;
                const int arow = p.row0 + rb * 256 + (int)e * 128, brow = ct * 256 + (int)e * 128;
                for (int kb = 0; kb < p.kblocks; kb++, it++) {
                    const int s = it % GP_STAGES;
                    GP_MARK(1, 4, it);
                    rw_wait(empty0 + 8 * s, ((uint32_t)(it / GP_STAGES) & 1u) ^ 1u);
                    const uint32_t st = tiles + (uint32_t)s * GP_STAGE_BYTES;
                    if (e == 0) mbar_expect_tx(full0 + 8 * s, 2 * per_cta);
                    rw2_tma_load_to_leader(st, &map_a_hi, leader_full0 + 8 * s, kb * TC_BK, arow);
                    rw2_tma_load_to_leader(st + 2 * TC_TILE_BYTES, &map_b_hi, leader_full0 + 8 * s, kb * TC_BK, brow);
                    if (p.terms == 3) {
                        rw2_tma_load_to_leader(st + TC_TILE_BYTES, &map_a_lo, leader_full0 + 8 * s, kb * TC_BK, arow);
                        rw2_tma_load_to_leader(st + 3 * TC_TILE_BYTES, &map_b_lo, leader_full0 + 8 * s, kb * TC_BK, brow);
                    }
                    GP_MARK(1, 5, it);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader) =====
        if (lane == 0 && e == 0) {
            // D = f32, A = B = bf16, both K-major, M = 256 (128 rows per CTA), N = 256 (128 W^T rows per CTA)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            int it = 0, q = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, q++) {
                const int acc = q & 1;
                GP_MARK(2, 6, q);
                rw_wait(tempty0 + 8 * acc, (((uint32_t)q >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                for (int kb = 0; kb < p.kblocks; kb++, it++) {
                    const int s = it % GP_STAGES;
                    GP_MARK(2, 7, it);
                    rw_wait(full0 + 8 * s, (uint32_t)(it / GP_STAGES) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = tiles + (uint32_t)s * GP_STAGE_BYTES;
                    const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + TC_TILE_BYTES);
                    const uint64_t b_hi = umma_desc_sw128(st + 2 * TC_TILE_BYTES), b_lo = umma_desc_sw128(st + 3 * TC_TILE_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                        const uint64_t adv = (uint64_t)(k4 * 32 >> 4);
                        rw2_umma(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k4) != 0);
                        if (p.terms == 3) {
                            rw2_umma(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                            rw2_umma(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
                        }
                    }
                    rw2_commit_pair(empty0 + 8 * s, pair_mask);
                }
                rw2_commit_pair(tfull0 + 8 * acc, pair_mask);
                GP_MARK(2, 8, q);
            }
        }
    } else {
        // ===== epilogue (128 threads per CTA, thread = accumulator row = row of A owned by this CTA) =====
        const int qd = warp & 3;
        float4 *stg4 = reinterpret_cast<float4 *>(epi_stage + (warp - 2) * (32 * 36));
        int q = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs, q++) {
            const int rb = tile / p.n_ct, ct = tile - rb * p.n_ct;
            const int acc = q & 1;
            if (lane == 0) GP_MARK(warp + 1, 10, q);
            rw_wait(tfull0 + 8 * acc, ((uint32_t)q >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row0 = rb * 256 + (int)e * 128 + qd * 32;       // first row (inside this launch) of this warp's 32 rows
            const int n0 = ct * 256;
            uint32_t v[32];
            rw_tmem_ld32(v, tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)acc * 256u);
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
                float4 badd = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) badd = __ldg(reinterpret_cast<const float4 *>(p.bias + n0 + c * 32 + 4 * (lane & 7)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == 7) {
                    // accumulator fully read: hand it back to the leader's MMA issuer before the stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) gp_arrive_remote(tempty0 + 8 * acc, 0u);
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    stg4[lane * 9 + j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                if (c + 1 < 8) rw_tmem_ld32(v, tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)((c + 1) * 32));
                __syncwarp();
                const int c4 = lane & 7, rsub = lane >> 3;             // this lane: columns 4*c4..+3 of rows rsub + 4i
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = rsub + 4 * i;
                    float4 o = stg4[r * 9 + c4];
                    o.x += badd.x; o.y += badd.y; o.z += badd.z; o.w += badd.w;
                    __stcg(reinterpret_cast<float4 *>(p.C + (size_t)(row0 + r) * p.ldc + n0 + c * 32 + 4 * c4), o);
                }
                __syncwarp();
            }
            if (lane == 0) GP_MARK(warp + 1, 11, q);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) GP_MARK(0, 8, 0);
    rw_cluster_sync();                       // the peer may still arrive on this CTA's barriers / read its operand tiles
    if (threadIdx.x == 0) GP_MARK(0, 9, 0);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef GASR_RW_TRACE
        if (p.variant == 2) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
#endif
        GASR_TLOG(p.tlog, 1, 3);
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(own_base), "n"(512) : "memory");
        GASR_TLOG(p.tlog, 1, 4);
    }
}

#ifdef GASR_RW_TRACE
static unsigned *g_tlog_host = nullptr, *g_tlog_dev = nullptr;
unsigned *trace_tmem_log() {
    if (!getenv("GASR_GP_TRACE")) return nullptr;
    if (!g_tlog_host) {
        if (cudaHostAlloc((void **)&g_tlog_host, sizeof(unsigned) * 160 * 16, cudaHostAllocMapped) != cudaSuccess) return nullptr;
        cudaHostGetDevicePointer((void **)&g_tlog_dev, g_tlog_host, 0);
        for (int i = 0; i < 160 * 16; i++) g_tlog_host[i] = 0;
    }
    return g_tlog_dev;
}
void trace_tmem_log_dump() {
    if (!g_tlog_host) return;
    static const char *kn[] = {"?", "gemm_pair", "rnn_wide2", "xproj_stream", "rnn_wide"};
    static const char *en[] = {"?", "alloc..", "alloc ok", "dealloc..", "dealloc ok"};
    for (int sm = 0; sm < 160; sm++) {
        const unsigned *l = g_tlog_host + sm * 16;
        const unsigned n = l[8];
        if (!n) continue;
        const unsigned last = l[(n - 1) & 7];
        if (((last >> 24) & 15) == 4) continue;                    // the last event on this SM is a completed dealloc
        fprintf(stderr, "[tmem log] SM %3d (%u events):", sm, n);
        for (unsigned k = n >= 8 ? n - 8 : 0; k < n; k++) {
            const unsigned e = l[k & 7];
            fprintf(stderr, " %s#%u:%s", kn[(e >> 28) & 7], e & 0xffffffu, en[(e >> 24) & 15]);
        }
        fprintf(stderr, "\n");
    }
}
static unsigned *g_gp_trace_host = nullptr, *g_gp_trace_dev = nullptr;
static int g_gp_slot = 0;
static int g_gp_info[16][4];
constexpr int GP_TRACE_CTAS = 160;
void gemm_pair_trace_dump() {
    trace_tmem_log_dump();
    if (!g_gp_trace_host) return;
    for (int sl = 0; sl < 16; sl++) {
        const int ctas = g_gp_info[sl][0];
        int open = 0;
        for (int c = 0; c < ctas; c++) if ((g_gp_trace_host[(sl * GP_TRACE_CTAS + c) * 8] >> 24) != 9) open++;
        if (!open) continue;
        fprintf(stderr, "[gemm_pair trace] slot %d: %d CTAs, M=%d row0=%d tiles=%d -- %d CTAs not exited\n", sl, ctas, g_gp_info[sl][1],
                g_gp_info[sl][2], g_gp_info[sl][3], open);
        for (int c = 0; c < ctas; c++) {
            const unsigned *w = g_gp_trace_host + (sl * GP_TRACE_CTAS + c) * 8;
            if ((w[0] >> 24) == 9) continue;
            fprintf(stderr, "  cta %3d:", c);
            for (int r = 0; r < 8; r++) fprintf(stderr, " %u/%u", w[r] >> 24, w[r] & 0xffffffu);
            fprintf(stderr, "\n");
        }
    }
}
#else
void gemm_pair_trace_dump() {}
#endif

bool gemm_pair_supported(const gasr_ctx *ctx, int M, int H) { return ctx->cluster_ok && M >= 256 && M % 256 == 0 && H % 256 == 0; }

int gemm_pair_prepare(gasr_ctx *ctx) {
    if (ctx->attr_mask & 16384u) return GASR_OK;
    GASR_CUDA(cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM_BYTES));
    ctx->attr_mask |= 16384u;
    return GASR_OK;
}

// maps: A hi, A lo (box 128 rows), W^T hi, W^T lo (box 128 rows); C points at the first row of this launch
int launch_gemm_pair(gasr_ctx *ctx, const CUtensorMap maps[4], int row0, int M, int K, int H, float *C, int ldc, const float *bias,
                     int precision, cudaStream_t st) {
    GASR_CHECK(gemm_pair_supported(ctx, M, H), "gemm_pair: needs whole 256 x 256 tiles (M=%d, H=%d)", M, H);
    GASR_CHECK(ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_pair: output must be 16-byte aligned");
    GASR_TRY(gemm_pair_prepare(ctx));
    GemmPairParams p;
    p.M = M; p.row0 = row0; p.n_ct = H / 256; p.kblocks = ceil_div(K, TC_BK); p.terms = precision == GASR_PREC_BF16 ? 1 : 3;
    p.C = C; p.ldc = ldc; p.bias = bias; p.trace = nullptr; p.variant = 0; p.tlog = nullptr;
    const int n_tiles = (M / 256) * p.n_ct;
    int pairs = ctx->sm_count / 2;
    if (pairs > n_tiles) pairs = n_tiles;
#ifdef GASR_RW_TRACE
    if (const char *e = getenv("GASR_GP_VARIANT")) p.variant = atoi(e);
    p.tlog = trace_tmem_log();
    if (getenv("GASR_GP_TRACE")) {
        if (!g_gp_trace_host) {
            GASR_CUDA(cudaHostAlloc((void **)&g_gp_trace_host, sizeof(unsigned) * 16 * GP_TRACE_CTAS * 8, cudaHostAllocMapped));
            GASR_CUDA(cudaHostGetDevicePointer((void **)&g_gp_trace_dev, g_gp_trace_host, 0));
            for (int i = 0; i < 16 * GP_TRACE_CTAS * 8; i++) g_gp_trace_host[i] = 9u << 24;
        }
        const int sl = g_gp_slot++ & 15;
        g_gp_info[sl][0] = 2 * pairs; g_gp_info[sl][1] = M; g_gp_info[sl][2] = row0; g_gp_info[sl][3] = n_tiles;
        for (int i = 0; i < 2 * pairs * 8; i++) g_gp_trace_host[sl * GP_TRACE_CTAS * 8 + i] = 0;
        p.trace = g_gp_trace_dev + sl * GP_TRACE_CTAS * 8;
    }
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(GP_THREADS);
    cfg.dynamicSmemBytes = GP_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GASR_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel, maps[0], maps[1], maps[2], maps[3], p));
    ctx->launches += 1;
    return GASR_OK;
}

}  // namespace gasr
