// xproj_stream.cu -- persistent tcgen05 GEMM that feeds the streaming pipeline (stream_pipeline.cu): the input
// projections x*W_ih of layers 1..L-1 (reference RNN_Cell.cu:66, batched over frames) and the output layer
// Linear + log-softmax (Linear.cu:42-49 without the ReLU + baseline/model.py:49), computed block by block (128 rows =
// a few frames of the whole batch) AS SOON AS the recurrence kernel (rnn_stream.cu) has published the rows, and
// published back to it / to the decoder through progress counters in HBM.  No kernel boundary, no host in the loop.
//
// Same tile engine as xproj_gemm_tc.cu (TMA -> 3-stage smem ring -> tcgen05.mma, fp32 accumulator in TMEM, three
// bf16 hi/lo passes for fp32-grade results), made persistent:
//   warp 0    producer: waits for the source rows' progress counter, then cp.async.bulk.tensor loads of the
//             [128 x 64] A hi/lo tiles (the bf16 planes the recurrence kernel wrote) and the W^T tiles;
//   warp 1    MMA issuer; the accumulator is double buffered in TMEM (2 x 128 columns) so the next tile's MMAs
//             overlap the previous tile's epilogue;
//   warps 2-5 epilogue: tcgen05.ld -> + bias -> fp32 xproj rows, or (output layer, N = 32) a thread-local
//             log-softmax -- one thread owns one row of the accumulator -- then fence + one atomic per warp and tile.
// Every target owns a fixed range of CTAs and processes its (block, column tile) items in block order, so a target is
// never stuck behind another target's unmet dependency and the kernel cannot deadlock while all its CTAs are resident.
#include <math.h>

#include "common.cuh"
#include "stream.cuh"
#include "tc_common.cuh"

namespace gasr {

constexpr int XS_SMEM_EXTRA = 1024 /*alignment slack*/ + 128 /*barriers*/ + 4 * 32 * 36 * 4 /*epilogue staging*/;
constexpr int XS_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + XS_SMEM_EXTRA;
static inline int xs_stage_bytes(int wide) { return wide ? 2 * TC_TILE_BYTES + 2 * 256 * TC_BK * 2 : TC_STAGE_BYTES; }   // A hi/lo + B hi/lo (B up to 256 rows)
static inline int xs_smem_bytes(int stages, int wide) { return stages * xs_stage_bytes(wide) + XS_SMEM_EXTRA; }
constexpr unsigned long long XS_TIMEOUT_NS = 2000000000ull;

__device__ __forceinline__ bool xs_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long xs_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded mbarrier wait: returns false (and raises the abort flag) if the watchdog fires
// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread (asynchronous: tcgen05.wait::ld)
__device__ __forceinline__ void xs_tmem_ld32(uint32_t (&v)[32], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

__device__ __forceinline__ bool xs_wait(uint32_t bar, uint32_t parity, volatile int *abort_flag, volatile unsigned *gabort) {
    if (xs_mbar_try(bar, parity)) return true;
    const unsigned long long t0 = xs_now_ns();
    int spins = 0;
    while (!xs_mbar_try(bar, parity)) {
        if (*abort_flag) return false;
        if ((++spins & 1023) == 0 && (*gabort || xs_now_ns() - t0 > XS_TIMEOUT_NS)) { *abort_flag = 1; *gabort = 1u; return false; }
        __nanosleep(40);        // these CTAs share their SM with decoder CTAs: do not burn issue slots while waiting
    }
    return true;
}
__device__ __forceinline__ void xs_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Every target (a layer's projection, or the output layer) owns a fixed range of CTAs; its items (block, column tile)
// are dealt round-robin, in block order, to those CTAs.  A target only depends on the recurrence of the layer below,
// which in turn only depends on that layer's own target: no cross-target head-of-line blocking, no deadlock as long as
// all CTAs are resident.
struct XsIter {
    int tg, rank, nctas, idx, n_items, n_tiles;
    int blk, tile;
};
__device__ __forceinline__ bool xs_iter_init(const XsParams &p, XsIter &it) {
    it.tg = -1;
    for (int i = 0; i < p.n_targets; i++)
        if ((int)blockIdx.x >= p.target[i].cta0 && (int)blockIdx.x < p.target[i].cta0 + p.target[i].nctas) it.tg = i;
    if (it.tg < 0) return false;
    it.rank = (int)blockIdx.x - p.target[it.tg].cta0;
    it.nctas = p.target[it.tg].nctas;
    it.n_tiles = p.target[it.tg].n_tiles;
    it.n_items = p.n_blocks * it.n_tiles;
    it.idx = it.rank;
    it.blk = it.idx / it.n_tiles; it.tile = it.idx - it.blk * it.n_tiles;
    return it.idx < it.n_items;
}
__device__ __forceinline__ bool xs_iter_next(XsIter &it) {
    it.idx += it.nctas;
    it.blk = it.idx / it.n_tiles; it.tile = it.idx - it.blk * it.n_tiles;
    return it.idx < it.n_items;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
xproj_stream_kernel(const __grid_constant__ XsMaps maps, const XsParams p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ int abort_s;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const int NS = p.stages;                                             // ring depth: 3, or 2 (the wave engine: leaves shared memory for decoder CTAs on the same SM)
    const uint32_t SB = p.wide ? (uint32_t)(2 * TC_TILE_BYTES + 2 * 256 * TC_BK * 2) : (uint32_t)TC_STAGE_BYTES;   // stage bytes
    const uint32_t OFF_BLO = 2 * TC_TILE_BYTES + (p.wide ? 256u : 128u) * TC_BK * 2;                                  // B lo inside a stage
    const uint32_t ACC = p.wide ? 256u : (uint32_t)TC_BN;                                                             // TMEM columns per accumulator
    const uint32_t bars = tiles + NS * SB;                               // full[S], empty[S], tfull[2], tempty[2], tmem slot
    const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tfull0 = bars + 16 * TC_STAGES, tempty0 = tfull0 + 16;
    unsigned char *gen_tiles = smem_raw + (tiles - raw);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen_tiles + NS * SB + 16 * TC_STAGES + 32);
    volatile int *abort_flag = &abort_s;
    float *epi_stage = reinterpret_cast<float *>(gen_tiles + NS * SB + 128);     // [4 warps][32][36] floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        abort_s = 0;
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        GASR_TLOG(p.tlog, 3, 1);
        if (p.wide) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(512) : "memory");
        else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(2 * TC_BN) : "memory");
        GASR_TLOG(p.tlog, 3, 2);
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: progress counter -> TMA =====
        if (lane == 0) {
            int it = 0;                                                  // k-blocks issued so far (ring position)
            XsIter wi;
            for (bool have = xs_iter_init(p, wi); have && !*abort_flag; have = xs_iter_next(wi)) {
                const int blk = wi.blk, tg = wi.tg, tile = wi.tile;
                const XsTarget &t = p.target[tg];
                {   // the source rows of this block are complete (every CTA of the producing layer counted the block)
                    const volatile unsigned *flag = t.src_done + blk;
                    if (*flag < (unsigned)t.src_need) {
                        const unsigned long long t0 = xs_now_ns();
                        while (*flag < (unsigned)t.src_need) {
                            if (*abort_flag || *p.abort || xs_now_ns() - t0 > XS_TIMEOUT_NS) { *abort_flag = 1; *p.abort = 1u; if (p.error) { *reinterpret_cast<volatile int *>(p.error) = 3; __threadfence_system(); } break; }
                            __nanosleep(100);
                        }
                    }
                    __threadfence();                                     // acquire side of the counter
                    asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy writes -> async-proxy (TMA) reads
                }
                if (*abort_flag) break;
                const CUtensorMap *ma_hi = &maps.m[4 * tg + 0], *ma_lo = &maps.m[4 * tg + ((t.kind & 32) ? 0 : 1)];
                const CUtensorMap *mb_hi = &maps.m[4 * tg + 2], *mb_lo = &maps.m[4 * tg + ((t.kind & 32) ? 2 : 3)];
                const uint32_t btile = (uint32_t)t.bn * TC_BK * 2;
                const uint32_t bytes = (t.terms == 3 ? 2u : 1u) * (TC_TILE_BYTES + btile);
                for (int kb = 0; kb < t.kblocks; kb++, it++) {
                    const int s = it % NS;
                    if (!xs_wait(empty0 + 8 * s, ((it / NS) & 1) ^ 1, abort_flag, p.abort)) break;
                    const uint32_t st = tiles + s * SB;
                    mbar_expect_tx(full0 + 8 * s, bytes);
                    tma_load_2d(st, ma_hi, full0 + 8 * s, kb * TC_BK, p.row0 + blk * TC_BM);
                    tma_load_2d(st + 2 * TC_TILE_BYTES, mb_hi, full0 + 8 * s, kb * TC_BK, tile * t.bn);
                    if (t.terms == 3) {
                        tma_load_2d(st + TC_TILE_BYTES, ma_lo, full0 + 8 * s, kb * TC_BK, p.row0 + blk * TC_BM);
                        tma_load_2d(st + OFF_BLO, mb_lo, full0 + 8 * s, kb * TC_BK, tile * t.bn);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0, q = 0;
            XsIter wi;
            for (bool have = xs_iter_init(p, wi); have && !*abort_flag; have = xs_iter_next(wi), q++) {
                const int tg = wi.tg;
                const XsTarget &t = p.target[tg];
                const int acc = q & 1;
                if (!xs_wait(tempty0 + 8 * acc, ((q >> 1) & 1) ^ 1, abort_flag, p.abort)) break;       // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // instruction descriptor: D = f32, A = B = bf16, both K-major, N = bn, M = 128
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(t.bn >> 3) << 17) | ((TC_BM >> 4) << 24);
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC;
                bool ok = true;
                for (int kb = 0; kb < t.kblocks; kb++, it++) {
                    const int s = it % NS;
                    if (!xs_wait(full0 + 8 * s, (it / NS) & 1, abort_flag, p.abort)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = tiles + s * SB;
                    const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + TC_TILE_BYTES);
                    const uint64_t b_hi = umma_desc_sw128(st + 2 * TC_TILE_BYTES), b_lo = umma_desc_sw128(st + OFF_BLO);
#pragma unroll
                    for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                        const uint64_t adv = (uint64_t)(k4 * 32 >> 4);
                        umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k4) != 0);
                        if (t.terms == 3) {
                            umma_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                            umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
                        }
                    }
                    umma_commit(empty0 + 8 * s);
                }
                if (!ok) break;
                umma_commit(tfull0 + 8 * acc);
            }
        }
    } else {
        // ===== epilogue (128 threads, thread = accumulator row) =====
        const int qd = warp & 3;                                         // TMEM lane quadrant this warp may access
        int q = 0;
        XsIter wi;
        // The "rows stored" signal of a tile (fence + count) is deferred into the next tile: issued right after that
        // tile's first TMEM load, the fence finds this warp's stores already drained instead of waiting for them
        // (it was 16% of the epilogue's samples).  If the next accumulator is not ready yet, publish at once.
        unsigned *pend = nullptr;
        for (bool have = xs_iter_init(p, wi); have; have = xs_iter_next(wi), q++) {
            const int blk = wi.blk, tg = wi.tg, tile = wi.tile;
            const XsTarget &t = p.target[tg];
            const int acc = q & 1;
            if (pend && !xs_mbar_try(tfull0 + 8 * acc, (q >> 1) & 1)) {
                if (lane == 0) { __threadfence(); atomicAdd(pend, 1u); }
                pend = nullptr;
            }
            if (!xs_wait(tfull0 + 8 * acc, (q >> 1) & 1, abort_flag, p.abort)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = blk * TC_BM + qd * 32 + lane;
            const int n0 = tile * t.bn;
            const int nchunks = t.bn / 32;
            uint32_t v[32];
            xs_tmem_ld32(v, tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)acc * ACC);
            if (pend) {
                if (lane == 0) { __threadfence(); atomicAdd(pend, 1u); }
                pend = nullptr;
            }
            for (int c = 0; c < nchunks; c++) {
                // bias of this lane's store columns: requested before the TMEM load is waited for
                const float *bsrc = t.bias ? t.bias + n0 + c * 32 : nullptr;
                float4 badd = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((t.kind & 15) == XS_KIND_XPROJ && bsrc) badd = __ldg(reinterpret_cast<const float4 *>(bsrc + 4 * (lane & 7)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == nchunks - 1) {
                    // accumulator fully read: hand it back to the MMA issuer before the (slow) stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) xs_arrive(tempty0 + 8 * acc);
                }
                // registers (thread = row) -> shared staging tile [32 rows][36 floats] -> coalesced 128-byte row segments:
                // a scattered store (every lane its own row) costs 32 L2 transactions per instruction
                float4 *stg4 = reinterpret_cast<float4 *>(epi_stage + (warp - 2) * (32 * 36));
                if ((t.kind & 15) == XS_KIND_XPROJ) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        stg4[lane * 9 + j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                } else {
                    // output layer: bias + log-softmax over the first V columns, thread-local
                    float f[32];
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        f[j] = __uint_as_float(v[j]) + ((bsrc && j < t.V) ? bsrc[j] : 0.0f);
                        if (j < t.V) mx = fmaxf(mx, f[j]);
                    }
                    float sum = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; j++) if (j < t.V) sum += expf(f[j] - mx);
                    const float lse = logf(sum);
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = j < t.V ? (f[j] - mx) - lse : 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; j++) stg4[lane * 9 + j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                }
                // the registers are staged: fetch the next 32 columns while this chunk is being stored
                if (c + 1 < nchunks)
                    xs_tmem_ld32(v, tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)acc * ACC + (uint32_t)((c + 1) * 32));
                __syncwarp();
                if (!(t.kind & 16)) {
                    const int c4 = lane & 7, rsub = lane >> 3;                   // this lane: columns 4*c4..+3 of rows rsub + 4i
                    const int row0 = blk * TC_BM + qd * 32;
                    // output layer: kind bit 128 = the caller's matrix is exactly [rows, V] wide inside ldc (gasr_linear_forward):
                    // nothing beyond column V is touched; otherwise (the pipelines' own padded buffers) 32 columns, zeros beyond V
                    const int ncols = (t.kind & 15) == XS_KIND_XPROJ ? 32 : ((t.kind & 128) ? t.V : (t.ldc < 32 ? t.ldc : 32));
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = rsub + 4 * i;
                        float4 o = stg4[r * 9 + c4];
                        o.x += badd.x; o.y += badd.y; o.z += badd.z; o.w += badd.w;
                        if (row0 + r >= p.M || 4 * c4 >= ncols) continue;
                        float *dst = t.C + (size_t)(row0 + r) * t.ldc + n0 + c * 32 + 4 * c4;
                        if (4 * c4 + 4 <= ncols) __stcg(reinterpret_cast<float4 *>(dst), o);
                        else {                                                   // scalar tail of the last group of four
                            dst[0] = o.x;
                            if (4 * c4 + 1 < ncols) dst[1] = o.y;
                            if (4 * c4 + 2 < ncols) dst[2] = o.z;
                        }
                    }
                }
                __syncwarp();
            }
            // this warp's 32 rows are stored: fence + count (the consumer waits for XS_EPI_WARPS counts per tile), deferred
            __syncwarp();
            pend = t.dst_ready + blk;
        }
        if (pend && lane == 0) { __threadfence(); atomicAdd(pend, 1u); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        GASR_TLOG(p.tlog, 3, 3);
        if (p.wide) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TC_BN) : "memory");
        GASR_TLOG(p.tlog, 3, 4);
    }
}

int launch_xproj_stream(gasr_ctx *ctx, const XsMaps &maps, const XsParams &p, int ctas, cudaStream_t st) {
    GASR_CHECK(ctas >= 0 && p.n_targets >= 1 && p.n_targets <= XS_MAX_TARGETS, "xproj_stream: bad parameters");
    for (int i = 0; i < p.n_targets; i++) {
        const XsTarget &t = p.target[i];
        GASR_CHECK(t.C && t.src_done && t.dst_ready && t.kblocks >= 1 && (t.bn == 128 || t.bn == 32 || (t.bn == 256 && p.wide)) && (t.terms == 1 || t.terms == 3),
                   "xproj_stream: bad target %d", i);
        GASR_CHECK((t.kind & 15) == XS_KIND_XPROJ || (t.bn == 32 && t.n_tiles == 1 && t.V >= 1 && t.V <= 32), "xproj_stream: bad output-layer target");
        GASR_CHECK(ctas == 0 || (t.nctas >= 1 && t.cta0 >= 0 && t.cta0 + t.nctas <= ctas), "xproj_stream: bad CTA range of target %d", i);
        GASR_CHECK(t.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(t.C) & 15) == 0, "xproj_stream: output must be 16-byte aligned");
    }
    if (!(ctx->attr_mask & 8u)) {       // once per context, never while the pipeline's other kernels are running
        GASR_CUDA(cudaFuncSetAttribute(xproj_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xs_smem_bytes(2, 1) > XS_SMEM_BYTES ? xs_smem_bytes(2, 1) : XS_SMEM_BYTES));
        ctx->attr_mask |= 8u;
    }
    if (ctas == 0) return GASR_OK;      // preparation call
    GASR_CHECK(p.stages == 0 || p.stages == 2 || p.stages == 3, "xproj_stream: ring depth must be 2 or 3");
    XsParams q = p;
    q.tlog = nullptr;
#ifdef GASR_RW_TRACE
    q.tlog = trace_tmem_log();
#endif
    if (q.stages == 0) q.stages = TC_STAGES;
    if (q.wide) q.stages = 2;              // 96 KB stages: two fit
    xproj_stream_kernel<<<ctas, TC_THREADS, xs_smem_bytes(q.stages, q.wide), st>>>(maps, q);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

// ---- one-shot use of the tile engine: Linear (<= 32 outputs) + log-softmax over a large row count -----------------------
// The SIMT kernel of linear.cu keeps W in shared memory and reads x with broadcast loads; for wide inputs (cfg3: 1600) it
// has one CTA per SM and ~1 KB in flight per SM, 15 ms for 256 000 rows.  Here: x -> bf16 hi/lo planes (one pass), then the
// persistent GEMM above with a single output-layer target whose dependencies are preset -- TMA-fed, fp32-grade 3-term
// split, log-softmax in the epilogue.
bool linear_tc_supported(int rows, int in, int out, int ldy, const float *y, int act) {
    return act == GASR_ACT_LOGSOFTMAX && out >= 1 && out <= 32 && in >= 1024 && rows >= 4096 && ldy >= out && ldy % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(y) & 15) == 0;
}

int launch_linear_logsoftmax_tc(gasr_ctx *ctx, const float *x, int ldx, const float *W, const float *b, float *y, int ldy,
                                int rows, int in, int out, cudaStream_t st) {
    const int Kp = ceil_div(in, TC_BK) * TC_BK, nb = ceil_div(rows, TC_BM);
    size_t o = 0;
    const size_t off_wpad = o; o = align_up(o + sizeof(float) * (size_t)in * 32, 1024);
    const size_t off_bpad = o; o = align_up(o + 128, 1024);
    const size_t off_wbuf = o; o = align_up(o + xproj_tc_w_bytes(in, 32), 1024);
    const size_t off_abuf = o; o = align_up(o + xproj_tc_a_bytes(rows, in), 1024);
    const size_t off_flags = o; o = align_up(o + sizeof(unsigned) * (2 * (size_t)nb + 32), 1024);
    GASR_TRY(ws_reserve(ctx, ctx->ws_lin, o));
    unsigned char *base = static_cast<unsigned char *>(ctx->ws_lin.ptr);
    float *wpad = reinterpret_cast<float *>(base + off_wpad), *bpad = reinterpret_cast<float *>(base + off_bpad);
    unsigned *flags = reinterpret_cast<unsigned *>(base + off_flags);
    // W[in, out] -> [in, 32] (zero columns beyond out), bias -> 32 entries
    GASR_CUDA(cudaMemsetAsync(base + off_wpad, 0, off_wbuf - off_wpad, st));
    GASR_CUDA(cudaMemcpy2DAsync(wpad, 32 * sizeof(float), W, (size_t)out * sizeof(float), (size_t)out * sizeof(float), in,
                                cudaMemcpyDeviceToDevice, st));
    if (b) GASR_CUDA(cudaMemcpyAsync(bpad, b, sizeof(float) * out, cudaMemcpyDeviceToDevice, st));
    GASR_TRY(xproj_tc_prepare_weights(ctx, wpad, in, 32, base + off_wbuf, st));
    GASR_TRY(xproj_tc_split_rows(ctx, x, ldx, rows, in, base + off_abuf, st));
    GASR_CUDA(cudaMemsetAsync(flags, 0xff, sizeof(unsigned) * nb, st));                          // every source block complete
    GASR_CUDA(cudaMemsetAsync(flags + nb, 0, sizeof(unsigned) * ((size_t)nb + 32), st));         // tile counters, abort, error
    XsMaps maps;
    GASR_TRY(tc_make_map(&maps.m[0], base + off_abuf, rows, Kp, TC_BM));
    GASR_TRY(tc_make_map(&maps.m[1], base + off_abuf + xproj_tc_a_bytes(rows, in) / 2, rows, Kp, TC_BM));
    GASR_TRY(tc_make_map(&maps.m[2], base + off_wbuf, 32, Kp, 32));
    GASR_TRY(tc_make_map(&maps.m[3], base + off_wbuf + xproj_tc_w_bytes(in, 32) / 2, 32, Kp, 32));
    XsParams p = {};
    p.M = rows; p.n_blocks = nb; p.n_targets = 1;
    p.abort = flags + 2 * (size_t)nb;
    p.error = reinterpret_cast<int *>(flags + 2 * (size_t)nb + 8);
    XsTarget &t = p.target[0];
    t.kind = XS_KIND_LOGSOFTMAX | 128; t.cta0 = 0; t.nctas = nb < ctx->sm_count ? nb : ctx->sm_count;   // 128: columns >= out of y stay untouched
    // (every dependency of this launch is preset, so its bounded waits cannot expire; the abort / error words are not read back)
    t.n_tiles = 1; t.bn = 32; t.kblocks = Kp / TC_BK; t.terms = 3; t.V = out;
    t.C = y; t.ldc = ldy; t.bias = bpad;
    t.src_done = flags; t.src_need = 1; t.dst_ready = flags + nb;
    return launch_xproj_stream(ctx, maps, p, t.nctas, st);
}

}  // namespace gasr
