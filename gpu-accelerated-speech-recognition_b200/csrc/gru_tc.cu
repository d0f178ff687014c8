// gru_tc.cu -- one GRU timestep of a (layer, direction) for a whole batch as ONE tcgen05 kernel:
//   hh = h_{t-1} * W_hh  (bf16 hi/lo split, three MMA terms, fp32 accumulation in TMEM)
//   r = sigmoid(xp_r + hh_r + b_r),  z = sigmoid(xp_z + hh_z + b_z),  n = tanh(xp_n + r * (hh_n + b_n)),
//   h_t = (1 - z) * n + z * h_{t-1}                                   (torch.nn.GRU, gate order r, z, n)
// with the gate math in the GEMM epilogue.  The reference has no GRU (torch CPU defines it); the
// time loop this replaces is the shape of RNN::forward (reference RNN.cu:9-30), one cell call per timestep.
//
// Why fused: the previous path ran three kernels per timestep (split h into bf16 planes, GEMM, gates), 29 us per step
// at cfg3 widths, most of it launch gaps and the 2 x 2.4 MB round trip of hh.  Here
//   * the columns of W_hh^T are permuted at preparation time so that an accumulator tile holds the r, z and n
//     pre-activations of the SAME 32 hidden units: tile = [r(32) | z(32) | n(32)] = 96 columns (UMMA N = 96);
//   * an epilogue thread of accumulator row n (utterance n; two threads per row, 16 units each) therefore has everything
//     it needs for its units of h_t: it adds the input projections and biases, applies the gates, writes h_t (fp32, the layer's output) AND the bf16
//     hi/lo planes of h_t that the next step's TMA loads read (ping-pong plane buffers) -- no split kernel, no hh;
//   * the epilogue threads request their xp / h_{t-1} values before they wait for the accumulator, so the scattered
//     thread-per-row loads complete while the MMAs run.
// Grid = ceil(H / 32) x ceil(N / 128) CTAs (25 x 2 at cfg3; both directions of a layer run concurrently on two streams).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gasr {

constexpr int GT_UNITS = 32;                 // hidden units per accumulator tile
constexpr int GT_BN = 3 * GT_UNITS;          // 96 accumulator columns: r | z | n
constexpr int GT_B_TILE_BYTES = GT_BN * TC_BK * 2;
constexpr int GT_THREADS = 320;              // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)

struct GruTcParams {
    int N, H, kblocks, Kp;
    const float *xp; int ldxp;               // x * W_ih + b_ih of this timestep, [N, >= 3H]
    const float *b_hh;                       // [3H]
    const float *hprev; int ldh;             // h_{t-1} fp32 (nullptr at the first step: h_0 = 0)
    float *out; int ldo;                     // h_t fp32
    __nv_bfloat16 *nhi, *nlo;                // bf16 planes of h_t for the next step, [N, Kp]
};

// Gate non-linearities on the SFU (ex2.approx, rcp.approx; absolute error ~1e-7, the parity bar is 1e-4): with libm's
// expf / division / tanhf the epilogue's 48 transcendental values per thread took 4.9 us of a 17 us step.
__device__ __forceinline__ float gt_sigmoid(float v) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return r;
}
__device__ __forceinline__ float gt_tanh(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ void gt_tmem_ld16(uint32_t (&v)[16], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

__global__ void __launch_bounds__(GT_THREADS, 1)
gru_tc_step_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const GruTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = tiles + TC_STAGES * TC_STAGE_BYTES;            // full[S], empty[S], tmem_full, tmem_slot
    const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tfull = bars + 16 * TC_STAGES;
    unsigned char *gen_tiles = smem_raw + (tiles - raw);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen_tiles + TC_STAGES * TC_STAGE_BYTES + 16 * TC_STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, m0 = blockIdx.y * TC_BM;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
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

    // Programmatic dependent launch: this grid may start while the previous timestep's grid is still running.  Let the
    // next one start as early as it can, do everything that does not depend on h_{t-1} (barriers, TMEM, the first W_hh
    // tiles) and only then wait for the previous step to complete.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 0) {
        // ===== TMA producer: h_{t-1} planes [128 utterances x 64] and permuted W_hh^T planes [96 columns x 64] =====
        if (lane == 0) {
            const int pre = p.kblocks < TC_STAGES ? p.kblocks : TC_STAGES;
            for (int kb = 0; kb < pre; kb++) {                            // weights: independent of the previous step
                const uint32_t st = tiles + kb * TC_STAGE_BYTES;
                mbar_expect_tx(full0 + 8 * kb, 2 * TC_TILE_BYTES + 2 * GT_B_TILE_BYTES);
                tma_load_2d(st + 2 * TC_TILE_BYTES, &map_b_hi, full0 + 8 * kb, kb * TC_BK, tile * GT_BN);
                tma_load_2d(st + 3 * TC_TILE_BYTES, &map_b_lo, full0 + 8 * kb, kb * TC_BK, tile * GT_BN);
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");             // the planes were written with generic stores
            for (int kb = 0; kb < pre; kb++) {
                const uint32_t st = tiles + kb * TC_STAGE_BYTES;
                tma_load_2d(st, &map_a_hi, full0 + 8 * kb, kb * TC_BK, m0);
                tma_load_2d(st + TC_TILE_BYTES, &map_a_lo, full0 + 8 * kb, kb * TC_BK, m0);
            }
            for (int kb = pre; kb < p.kblocks; kb++) {
                const int s = kb % TC_STAGES;
                mbar_wait(empty0 + 8 * s, ((kb / TC_STAGES) & 1) ^ 1);
                const uint32_t st = tiles + s * TC_STAGE_BYTES;
                mbar_expect_tx(full0 + 8 * s, 2 * TC_TILE_BYTES + 2 * GT_B_TILE_BYTES);
                tma_load_2d(st, &map_a_hi, full0 + 8 * s, kb * TC_BK, m0);
                tma_load_2d(st + 2 * TC_TILE_BYTES, &map_b_hi, full0 + 8 * s, kb * TC_BK, tile * GT_BN);
                tma_load_2d(st + TC_TILE_BYTES, &map_a_lo, full0 + 8 * s, kb * TC_BK, m0);
                tma_load_2d(st + 3 * TC_TILE_BYTES, &map_b_lo, full0 + 8 * s, kb * TC_BK, tile * GT_BN);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 96, M = 128
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((GT_BN >> 3) << 17) | ((TC_BM >> 4) << 24);
            for (int kb = 0; kb < p.kblocks; kb++) {
                const int s = kb % TC_STAGES;
                mbar_wait(full0 + 8 * s, (kb / TC_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = tiles + s * TC_STAGE_BYTES;
                const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + TC_TILE_BYTES);
                const uint64_t b_hi = umma_desc_sw128(st + 2 * TC_TILE_BYTES), b_lo = umma_desc_sw128(st + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                    const uint64_t adv = (uint64_t)(k4 * 32 >> 4);
                    umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k4) != 0);
                    umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                    umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(tfull);
        }
    } else {
        // ===== epilogue: thread = utterance (accumulator row); 16 hidden units of h_t per thread (warps 2-5: units 0-15,
        //       warps 6-9: units 16-31 of the tile; a warp may only read the TMEM lane quadrant warp % 4) =====
        asm volatile("griddepcontrol.wait;" ::: "memory");      // h_{t-1} is read and the other plane buffer written below
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = m0 + q * 32 + lane;
        const int j0 = tile * GT_UNITS;
        const bool live = row < p.N;
        const float *xr = p.xp + (size_t)(live ? row : 0) * p.ldxp;
        const float *hp = p.hprev ? p.hprev + (size_t)(live ? row : 0) * p.ldh : nullptr;
        // this thread's inputs (16 units x {xp_r, xp_z, xp_n, h_{t-1}}) do not depend on the MMAs: request them now, they
        // arrive while the main loop runs (thread-per-row accesses are scattered, their latency must not be exposed)
        const int jb = j0 + half * 16;
        float4 xr4[4], xz4[4], xn4[4], h4[4];
#pragma unroll
        for (int g4 = 0; g4 < 4; g4++) {
            const int j = jb + 4 * g4;
            const bool in = live && j < p.H;                             // H % 4 == 0: a group of four is all in or all out
            const int jc = in ? j : 0;
            xr4[g4] = *reinterpret_cast<const float4 *>(xr + jc);
            xz4[g4] = *reinterpret_cast<const float4 *>(xr + p.H + jc);
            xn4[g4] = *reinterpret_cast<const float4 *>(xr + 2 * p.H + jc);
            h4[g4] = hp ? *reinterpret_cast<const float4 *>(hp + jc) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            uint32_t ar[16], az[16], an[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16);
            gt_tmem_ld16(ar, taddr);
            gt_tmem_ld16(az, taddr + GT_UNITS);
            gt_tmem_ld16(an, taddr + 2 * GT_UNITS);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (live) {
#pragma unroll
                for (int g4 = 0; g4 < 4; g4++) {
                    const int j = jb + 4 * g4;
                    if (j >= p.H) break;
                    const float4 br4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + j));
                    const float4 bz4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + p.H + j));
                    const float4 bn4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + 2 * p.H + j));
                    const float xrv[4] = {xr4[g4].x, xr4[g4].y, xr4[g4].z, xr4[g4].w}, xzv[4] = {xz4[g4].x, xz4[g4].y, xz4[g4].z, xz4[g4].w};
                    const float xnv[4] = {xn4[g4].x, xn4[g4].y, xn4[g4].z, xn4[g4].w}, hv[4] = {h4[g4].x, h4[g4].y, h4[g4].z, h4[g4].w};
                    const float brv[4] = {br4.x, br4.y, br4.z, br4.w}, bzv[4] = {bz4.x, bz4.y, bz4.z, bz4.w};
                    const float bnv[4] = {bn4.x, bn4.y, bn4.z, bn4.w};
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        // first step: hh = b_hh exactly (h_0 = 0), as the unfused path computes it
                        const float gr = hp ? __uint_as_float(ar[4 * g4 + e]) + brv[e] : brv[e];
                        const float gz = hp ? __uint_as_float(az[4 * g4 + e]) + bzv[e] : bzv[e];
                        const float gn = hp ? __uint_as_float(an[4 * g4 + e]) + bnv[e] : bnv[e];
                        const float r = gt_sigmoid(xrv[e] + gr);
                        const float z = gt_sigmoid(xzv[e] + gz);
                        const float nn = gt_tanh(xnv[e] + r * gn);
                        o[e] = (1.0f - z) * nn + z * hv[e];
                    }
                    *reinterpret_cast<float4 *>(p.out + (size_t)row * p.ldo + j) = make_float4(o[0], o[1], o[2], o[3]);
                    // bf16 hi / lo planes of h_t: the next step's A operand
                    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        hi[e] = __float2bfloat16_rn(o[e]);
                        lo[e] = __float2bfloat16_rn(o[e] - __bfloat162float(hi[e]));
                    }
                    __nv_bfloat162 *dh = reinterpret_cast<__nv_bfloat162 *>(p.nhi + (size_t)row * p.Kp + j);
                    __nv_bfloat162 *dl = reinterpret_cast<__nv_bfloat162 *>(p.nlo + (size_t)row * p.Kp + j);
                    uint2 ph, pl;
                    __nv_bfloat162 t0 = __halves2bfloat162(hi[0], hi[1]), t1 = __halves2bfloat162(hi[2], hi[3]);
                    ph.x = *reinterpret_cast<uint32_t *>(&t0); ph.y = *reinterpret_cast<uint32_t *>(&t1);
                    t0 = __halves2bfloat162(lo[0], lo[1]); t1 = __halves2bfloat162(lo[2], lo[3]);
                    pl.x = *reinterpret_cast<uint32_t *>(&t0); pl.y = *reinterpret_cast<uint32_t *>(&t1);
                    *reinterpret_cast<uint2 *>(dh) = ph;
                    *reinterpret_cast<uint2 *>(dl) = pl;
                }
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

// W_hh[H, 3H] (reference layout [in, out], gates r | z | n) -> permuted W^T hi / lo planes [tiles * 96, Kp]:
// row tile * 96 + g * 32 + u  =  column g * H + tile * 32 + u of W_hh (zero beyond H)
__global__ void gru_perm_split_kernel(const float *__restrict__ w, int H, int Kp, int rows, __nv_bfloat16 *__restrict__ hi,
                                      __nv_bfloat16 *__restrict__ lo) {
    const size_t total = (size_t)rows * Kp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int prow = (int)(i / Kp), k = (int)(i % Kp);
        const int tile = prow / GT_BN, rem = prow - tile * GT_BN, g = rem / GT_UNITS, u = rem - g * GT_UNITS;
        const int j = tile * GT_UNITS + u;
        const float a = (k < H && j < H) ? w[(size_t)k * 3 * H + (size_t)g * H + j] : 0.0f;
        const __nv_bfloat16 ah = __float2bfloat16_rn(a);
        hi[i] = ah;
        lo[i] = __float2bfloat16_rn(a - __bfloat162float(ah));
    }
}

bool gru_tc_supported(int N, int H, int ldxp, int ldo, int col0) {
    return N >= 1 && H >= 32 && H % 4 == 0 && ldxp % 4 == 0 && ldo % 4 == 0 && col0 % 4 == 0;
}

static int gt_kp(int H) { return ceil_div(H, TC_BK) * TC_BK; }
static int gt_rows(int H) { return ceil_div(H, GT_UNITS) * GT_BN; }
size_t gru_tc_w_bytes(int H) { return 2 * align_up((size_t)gt_rows(H) * gt_kp(H) * 2, 1024); }
size_t gru_tc_plane_bytes(int N, int H) { return 2 * align_up((size_t)N * gt_kp(H) * 2, 1024); }     // hi + lo of one buffer

// Prepares the permuted weight planes, zeroes both h plane buffers (h_0 = 0 and the K padding) and builds the TMA
// descriptors: maps[b][0..1] = planes of buffer b (hi, lo), wmaps[0..1] = weights.
int gru_tc_prepare(gasr_ctx *ctx, GruTcPlan &pl, const float *w_hh, int N, int H, void *wbuf, void *planes, cudaStream_t st) {
    const int Kp = gt_kp(H), rows = gt_rows(H);
    pl.N = N; pl.H = H; pl.Kp = Kp;
    unsigned char *wb = static_cast<unsigned char *>(wbuf), *pb = static_cast<unsigned char *>(planes);
    const size_t whalf = gru_tc_w_bytes(H) / 2, phalf = gru_tc_plane_bytes(N, H) / 2;
    const size_t total = (size_t)rows * Kp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    gru_perm_split_kernel<<<blocks, 256, 0, st>>>(w_hh, H, Kp, rows, reinterpret_cast<__nv_bfloat16 *>(wb),
                                                  reinterpret_cast<__nv_bfloat16 *>(wb + whalf));
    GASR_CUDA(cudaGetLastError());
    GASR_CUDA(cudaMemsetAsync(planes, 0, 2 * gru_tc_plane_bytes(N, H), st));
    ctx->launches += 1;
    for (int b = 0; b < 2; b++) {
        pl.plane[b][0] = pb + (size_t)b * 2 * phalf;
        pl.plane[b][1] = pb + (size_t)b * 2 * phalf + phalf;
        GASR_TRY(tc_make_map(&pl.maps[b][0], pl.plane[b][0], N, Kp, TC_BM));
        GASR_TRY(tc_make_map(&pl.maps[b][1], pl.plane[b][1], N, Kp, TC_BM));
    }
    GASR_TRY(tc_make_map(&pl.wmaps[0], wb, rows, Kp, GT_BN));
    GASR_TRY(tc_make_map(&pl.wmaps[1], wb + whalf, rows, Kp, GT_BN));
    if (!(ctx->attr_mask & 2048u)) {
        GASR_CUDA(cudaFuncSetAttribute(gru_tc_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        ctx->attr_mask |= 2048u;
    }
    return GASR_OK;
}

// One timestep: reads the planes of buffer `src` (h_{t-1}), writes h_t to `out` and to the planes of buffer src ^ 1.
int gru_tc_step(gasr_ctx *ctx, const GruTcPlan &pl, int src, const float *xp, int ldxp, const float *b_hh, const float *hprev,
                int ldh, float *out, int ldo, bool overlap, cudaStream_t st) {
    GruTcParams p;
    p.N = pl.N; p.H = pl.H; p.Kp = pl.Kp; p.kblocks = pl.Kp / TC_BK;
    p.xp = xp; p.ldxp = ldxp; p.b_hh = b_hh; p.hprev = hprev; p.ldh = ldh; p.out = out; p.ldo = ldo;
    p.nhi = static_cast<__nv_bfloat16 *>(pl.plane[src ^ 1][0]); p.nlo = static_cast<__nv_bfloat16 *>(pl.plane[src ^ 1][1]);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(pl.H, GT_UNITS), ceil_div(pl.N, TC_BM));
    cfg.blockDim = dim3(GT_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;          // the first step of a sequence keeps the full dependency on the preparation
    GASR_CUDA(cudaLaunchKernelEx(&cfg, gru_tc_step_kernel, pl.maps[src][0], pl.maps[src][1], pl.wmaps[0], pl.wmaps[1], p));
    ctx->launches += 1;
    return GASR_OK;
}

}  // namespace gasr
