// xproj_gemm_tc.cu -- the batched input projection  C[M, N] = A[M, K] * W[K, N] + bias  on the 5th-generation tensor
// cores (tcgen05.mma, accumulator in TMEM, operands staged by TMA).  It replaces the per-timestep cublasSgemm of
// RNN_Cell::forward (reference RNN_Cell.cu:66 via cuMatrix.cpp:46-60) with ONE GEMM over all frames of a chunk.
//
// Precision (gasr.h GASR_PREC_*):
//   FP32: every fp32 operand is split into bf16 hi + bf16 lo (x = hi + lo + O(2^-17 |x|)); the kernel accumulates
//         hi*hi + hi*lo + lo*hi in the same fp32 TMEM accumulator -- three MMA passes per K block, fp32-grade result
//         (dropped lo*lo term ~2^-16 relative), still entirely on tensor cores;
//   BF16: hi*hi only (the "bf16 projection" of BASELINE.json cfg3).
//
// Kernel anatomy (one 128x128 output tile per CTA, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d of the A and W^T tiles ([128 rows x 64 bf16], 128-byte swizzle)
//            into a 3-stage shared-memory ring, completion on mbarriers (expect_tx);
//   warp 1   allocates 128 TMEM columns, then one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//            (M=128, N=128, K=16 per instruction; shared-memory matrix descriptors, K-major, SWIZZLE_128B) and
//            tcgen05.commit's each stage back to the producer and the finished tile to the epilogue;
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x32 (lane quadrant = warp % 4) -> + bias -> fp32 rows to HBM.
// Operand preparation (split_rows_kernel / split_transpose_kernel) converts fp32 to the padded bf16 hi/lo planes the
// TMA descriptors point at.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gasr {

struct TcParams {
    int M, N, kblocks, terms;     // terms = 3 (fp32-grade) or 1 (bf16)
    float *C; int ldc;
    const float *bias;
    int relu;                     // epilogue: max(. + bias, 0)  (Linear.cu:3-10)
};

__global__ void __launch_bounds__(TC_THREADS, 1)
xproj_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                     const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024-byte alignment
    // stage = A_hi, A_lo, B_hi, B_lo (fp32-grade) or just A_hi, B_hi (bf16): the bf16 mode's 96 KB ring lets two CTAs share
    // an SM, so one CTA's epilogue overlaps the other's main loop
    const uint32_t stage_bytes = (p.terms == 3 ? 4u : 2u) * TC_TILE_BYTES;
    const uint32_t off_b_hi = (p.terms == 3 ? 2u : 1u) * TC_TILE_BYTES;
    const uint32_t bars = tiles + TC_STAGES * stage_bytes;               // full[S], empty[S], tmem_full, tmem_slot
    const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tfull = bars + 16 * TC_STAGES;
    unsigned char *gen_tiles = smem_raw + (tiles - raw);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen_tiles + TC_STAGES * stage_bytes + 16 * TC_STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(TC_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < p.kblocks; kb++) {
                const int s = kb % TC_STAGES;
                mbar_wait(empty0 + 8 * s, ((kb / TC_STAGES) & 1) ^ 1);
                const uint32_t st = tiles + s * stage_bytes;
                mbar_expect_tx(full0 + 8 * s, stage_bytes);
                tma_load_2d(st, &map_a_hi, full0 + 8 * s, kb * TC_BK, m0);
                tma_load_2d(st + off_b_hi, &map_b_hi, full0 + 8 * s, kb * TC_BK, n0);
                if (p.terms == 3) {
                    tma_load_2d(st + TC_TILE_BYTES, &map_a_lo, full0 + 8 * s, kb * TC_BK, m0);
                    tma_load_2d(st + 3 * TC_TILE_BYTES, &map_b_lo, full0 + 8 * s, kb * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread issues for the CTA) =====
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 128, M = 128
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);
            for (int kb = 0; kb < p.kblocks; kb++) {
                const int s = kb % TC_STAGES;
                mbar_wait(full0 + 8 * s, (kb / TC_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = tiles + s * stage_bytes;
                const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + TC_TILE_BYTES);
                const uint64_t b_hi = umma_desc_sw128(st + off_b_hi), b_lo = umma_desc_sw128(st + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                    const uint64_t adv = (uint64_t)(k4 * 32 >> 4);     // 16 bf16 = 32 bytes along K inside the swizzle atom
                    umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k4) != 0);
                    if (p.terms == 3) {
                        umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
                    }
                }
                umma_commit(empty0 + 8 * s);                            // frees the stage when these MMAs retire
            }
            umma_commit(tfull);                                         // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias) -> HBM =====
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                                         // TMEM lane quadrant this warp may access
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; c++) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < p.M) {
                float *dst = p.C + (size_t)row * p.ldc + n0 + c * 32;
                const float *bsrc = p.bias ? p.bias + n0 + c * 32 : nullptr;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (n0 + c * 32 + j >= p.N) break;           // last column tile of an N that is not a multiple of 128 (N % 4 == 0)
                    float4 o;
                    o.x = __uint_as_float(v[j + 0]) + (bsrc ? bsrc[j + 0] : 0.0f);
                    o.y = __uint_as_float(v[j + 1]) + (bsrc ? bsrc[j + 1] : 0.0f);
                    o.z = __uint_as_float(v[j + 2]) + (bsrc ? bsrc[j + 2] : 0.0f);
                    o.w = __uint_as_float(v[j + 3]) + (bsrc ? bsrc[j + 3] : 0.0f);
                    if (p.relu) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                    *reinterpret_cast<float4 *>(dst + j) = o;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_BN) : "memory");
    }
}

// fp32 rows -> bf16 hi / lo planes [rows, Kp] (zero padded beyond K)
__global__ void split_rows_kernel(const float *__restrict__ x, int ldx, int rows, int K, int Kp,
                                  __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
    const size_t total = (size_t)rows * (Kp / 2);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / (Kp / 2)), c = (int)(i % (Kp / 2)) * 2;
        const float a = c < K ? x[(size_t)r * ldx + c] : 0.0f, b = c + 1 < K ? x[(size_t)r * ldx + c + 1] : 0.0f;
        const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
        const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
        reinterpret_cast<__nv_bfloat162 *>(hi)[i] = __halves2bfloat162(ah, bh);
        reinterpret_cast<__nv_bfloat162 *>(lo)[i] = __halves2bfloat162(al, bl);
    }
}

// W[K, N] (reference layout [in, out]) -> W^T hi / lo planes [N, Kp]
__global__ void split_transpose_kernel(const float *__restrict__ w, int K, int N, int Kp, __nv_bfloat16 *__restrict__ hi,
                                       __nv_bfloat16 *__restrict__ lo) {
    const size_t total = (size_t)N * Kp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / Kp), k = (int)(i % Kp);
        const float a = k < K ? w[(size_t)k * N + n] : 0.0f;
        const __nv_bfloat16 ah = __float2bfloat16_rn(a);
        hi[i] = ah;
        lo[i] = __float2bfloat16_rn(a - __bfloat162float(ah));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D bf16 tensor [rows, Kp] row-major, box [128 rows x 64 cols], 128-byte swizzle
int tc_make_map(CUtensorMap *map, const void *base, int rows, int Kp, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return GASR_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GASR_ERR_CUDA; }
    return GASR_OK;
}

bool xproj_tc_supported(int M, int K, int N) { return M >= 1 && K >= 1 && N >= TC_BN && N % 4 == 0; }

size_t xproj_tc_a_bytes(int M, int K) { return 2 * align_up((size_t)M * (size_t)ceil_div(K, TC_BK) * TC_BK * 2, 1024); }
size_t xproj_tc_w_bytes(int K, int N) { return 2 * align_up((size_t)N * (size_t)ceil_div(K, TC_BK) * TC_BK * 2, 1024); }

int xproj_tc_prepare_weights(gasr_ctx *ctx, const float *W, int K, int N, void *wbuf, cudaStream_t st) {
    const int Kp = ceil_div(K, TC_BK) * TC_BK;
    __nv_bfloat16 *hi = static_cast<__nv_bfloat16 *>(wbuf);
    __nv_bfloat16 *lo = reinterpret_cast<__nv_bfloat16 *>(static_cast<unsigned char *>(wbuf) + xproj_tc_w_bytes(K, N) / 2);
    const size_t total = (size_t)N * Kp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    split_transpose_kernel<<<blocks, 256, 0, st>>>(W, K, N, Kp, hi, lo);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

// fp32 A[M, K] -> bf16 hi/lo planes [M, Kp] in abuf (hi at abuf, lo at abuf + xproj_tc_a_bytes(M, K) / 2)
int xproj_tc_split_rows(gasr_ctx *ctx, const float *A, int lda, int M, int K, void *abuf, cudaStream_t st) {
    const int Kp = ceil_div(K, TC_BK) * TC_BK;
    __nv_bfloat16 *a_hi = static_cast<__nv_bfloat16 *>(abuf);
    __nv_bfloat16 *a_lo = reinterpret_cast<__nv_bfloat16 *>(static_cast<unsigned char *>(abuf) + xproj_tc_a_bytes(M, K) / 2);
    const size_t total = (size_t)M * (Kp / 2);
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    split_rows_kernel<<<blocks, 256, 0, st>>>(A, lda, M, K, Kp, a_hi, a_lo);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

// rows [row0, row0 + nrows) of fp32 A[M, K] -> the same rows of the bf16 hi/lo planes of a full-size abuf (M_total rows)
int xproj_tc_split_rows_range(gasr_ctx *ctx, const float *A, int lda, int M_total, int row0, int nrows, int K, void *abuf,
                              cudaStream_t st) {
    const int Kp = ceil_div(K, TC_BK) * TC_BK;
    __nv_bfloat16 *a_hi = static_cast<__nv_bfloat16 *>(abuf) + (size_t)row0 * Kp;
    __nv_bfloat16 *a_lo = reinterpret_cast<__nv_bfloat16 *>(static_cast<unsigned char *>(abuf) + xproj_tc_a_bytes(M_total, K) / 2) + (size_t)row0 * Kp;
    const size_t total = (size_t)nrows * (Kp / 2);
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
    split_rows_kernel<<<blocks, 256, 0, st>>>(A + (size_t)row0 * lda, lda, nrows, K, Kp, a_hi, a_lo);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

// A reusable launch plan: the four TMA descriptors of a fixed (A scratch, prepared W) pair.  Building descriptors costs
// microseconds of host time per call, which matters when the same GEMM runs once per timestep (GRU recurrence).
int xproj_tc_plan(XprojTcPlan &pl, int M, int K, int N, const void *wbuf, void *abuf) {
    GASR_CHECK(xproj_tc_supported(M, K, N), "xproj_tc: unsupported shape M=%d K=%d N=%d", M, K, N);
    const int Kp = ceil_div(K, TC_BK) * TC_BK;
    pl.M = M; pl.K = K; pl.N = N; pl.abuf = abuf;
    unsigned char *ab = static_cast<unsigned char *>(abuf);
    const unsigned char *wb = static_cast<const unsigned char *>(wbuf);
    GASR_TRY(tc_make_map(&pl.maps[0], ab, M, Kp, TC_BM));
    GASR_TRY(tc_make_map(&pl.maps[1], ab + xproj_tc_a_bytes(M, K) / 2, M, Kp, TC_BM));
    GASR_TRY(tc_make_map(&pl.maps[2], wb, N, Kp, TC_BM));
    GASR_TRY(tc_make_map(&pl.maps[3], wb + xproj_tc_w_bytes(K, N) / 2, N, Kp, TC_BM));
    return GASR_OK;
}

// C[M, N] = A[M, K] * W + bias through a plan: split A into the plan's bf16 planes, then the tcgen05 GEMM
int xproj_tc_run(gasr_ctx *ctx, const XprojTcPlan &pl, const float *A, int lda, const float *bias, float *C, int ldc,
                 int precision, cudaStream_t st, int relu) {
    GASR_CHECK(ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "xproj_tc: output must be 16-byte aligned");
    GASR_TRY(xproj_tc_split_rows(ctx, A, lda, pl.M, pl.K, pl.abuf, st));
    TcParams p;
    p.M = pl.M; p.N = pl.N; p.kblocks = ceil_div(pl.K, TC_BK); p.terms = precision == GASR_PREC_BF16 ? 1 : 3;
    p.C = C; p.ldc = ldc; p.bias = bias; p.relu = relu;
    if (!(ctx->attr_mask & 1024u)) {
        GASR_CUDA(cudaFuncSetAttribute(xproj_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        ctx->attr_mask |= 1024u;
    }
    dim3 grid(ceil_div(pl.N, TC_BN), ceil_div(pl.M, TC_BM));   // TMA zero-fills the rows of W^T beyond N
    const size_t smem = p.terms == 3 ? (size_t)TC_SMEM_BYTES : (size_t)TC_SMEM_BYTES - TC_STAGES * 2 * TC_TILE_BYTES;
    xproj_tcgen05_kernel<<<grid, TC_THREADS, smem, st>>>(pl.maps[0], pl.maps[1], pl.maps[2], pl.maps[3], p);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

// C[M, N] = A[M, K] * W + bias with W prepared by xproj_tc_prepare_weights; abuf is scratch of xproj_tc_a_bytes(M, K).
int launch_xproj_tc(gasr_ctx *ctx, const float *A, int lda, int M, int K, int N, const void *wbuf, void *abuf,
                    const float *bias, float *C, int ldc, int precision, cudaStream_t st, int relu) {
    XprojTcPlan pl;
    GASR_TRY(xproj_tc_plan(pl, M, K, N, wbuf, abuf));
    return xproj_tc_run(ctx, pl, A, lda, bias, C, ldc, precision, st, relu);
}

}  // namespace gasr
