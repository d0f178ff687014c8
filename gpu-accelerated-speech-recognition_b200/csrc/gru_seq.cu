// gru_seq.cu -- the GRU recurrence of one (layer, direction) as ONE persistent kernel with W_hh resident in shared memory
// (GASR_PREC_BF16 mode; BASELINE.json cfg3: 5-layer bidirectional GRU, H = 800, 256 utterances).
//
// Stands behind the time loop of RNN::forward (reference RNN.cu:9-30, one cell call per timestep); the GRU equations are
// torch.nn.GRU's (gate order r, z, n; the test oracle restates them).  The per-timestep kernel (gru_tc.cu) costs 14 us per step
// at cfg3 widths: every launch re-reads its slice of W_hh^T (hi + lo planes, 320 KB per CTA) through L2 and pays the
// launch / dependency gap.  Here
//   * a CTA owns UNITS hidden units (the r, z and n columns of W_hh for them: a [3 UNITS x Kp] K-major slice of the permuted
//     W_hh^T, SINGLE-PLANE FP16) and keeps that slice in shared memory for all T steps;
//   * per step and row block (128 utterances) it pulls h_{t-1} ([128 x Kp] fp16, 208 KB at H = 800) through a TMA ring, runs
//     tcgen05.mma (M = 128, N = 3 UNITS, fp32 accumulator in TMEM), and its eight epilogue warps (thread = utterance row,
//     UNITS / 2 units) add the input projection and biases, apply the gates, keep h_{t-1} of their units in REGISTERS, and
//     write h_t as fp16 into the other plane buffer (first: the peers wait for it) and as fp32 (the layer output);
//   * the CTAs exchange h_t through those planes in L2: per step the epilogue warps meet at a named barrier, one thread
//     fences and counts on the row block's counter, and the TMA producers acquire the full count before they load;
//   * with two row blocks per CTA (pp = 2, the default where the batch has them) the row blocks alternate: while the gates
//     of one are computed and released, the ring and the MMAs work on the other, which hides the exchange latency.
// What bounds a step (globaltimer stamps, `make TRACE=1`): the TMA ring -- the W slice leaves 64 KB (UNITS = 32) for
// h tiles in flight, and a [128 x 64] tile takes ~1.4 us from request to arrival, so 208 KB need 5.9 us; UNITS = 16 halves
// the slice, doubles the ring (8 x 16 KB) and with it the ingest rate.
// Grid = ceil(H / UNITS) x ceil(row blocks / pp) CTAs (50 x 1 at cfg3), all co-resident (one per SM, checked by the launcher);
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

constexpr int GS_A_TILE = TC_BM * TC_BK * 2;       // 16 KB: [128 x 64] fp16
constexpr int GS_THREADS = 320;                    // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int GS_MAX_STAGES = 8;
constexpr int GS_MAX_SMEM = 227 * 1024;

struct GruSeqParams {
    int T, N, H, KB, Kp, reverse, nst;             // KB = Kp / 64 k-blocks, nst = ring stages
    int pp;                                        // row blocks per CTA (1 or 2)
    const float *xp; int ldxp;                     // x * W_ih + b_ih, [T * N, >= 3H]
    const float *b_hh;                             // [3H]
    float *out; int ldo;                           // h fp32, [T * N, ldo], already offset to this direction's columns
    __half *plane[2];                              // fp16 planes of h, [N, Kp] each (ping-pong by step parity)
    unsigned *cnt;                                 // [row blocks] CTAs that released their columns of h, zeroed before the launch
    unsigned long long *trace;                     // instrumented build only (make TRACE=1): cycle sums of CTA (0, 0)
    unsigned long long *stamps;                    // instrumented build only: globaltimer stamps of steps 500..503, [CTA][step][row block][6]
};

#ifdef GASR_RW_TRACE
// cycle sums per role of CTA (0, 0): 0 producer waits for the peers' h, 1 producer waits for a free ring stage, 2 MMA issuer waits
// for a full stage, 3 epilogue waits for the accumulator, 4 epilogue math + plane stores, 5 epilogue release (barrier + fence + atomic)
#define GS_T0(v) const long long v = clock64()
#define GS_T1(v, k) do { gs_acc[k] += (unsigned long long)(clock64() - v); } while (0)   /* per-thread sums, written once at the end */
#define GS_TDECL unsigned long long gs_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define GS_TFLUSH do { if (p.trace && blockIdx.x == 1 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) <= 2) for (int k_ = 0; k_ < 8; k_++) if (gs_acc[k_]) p.trace[k_] = gs_acc[k_]; } while (0)
// stamps: 0 released, 1 producer saw every release, 2 first h tile arrived, 3 accumulator committed (MMA warp), 4 epilogue
// starts on the accumulator, 5 gate math done
#define GS_STAMP(s, g, k) do { if (p.stamps && (s) >= 500 && (s) < 504 && (threadIdx.x & 31) == 0) p.stamps[((((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 + ((s) - 500)) * 2 + (g)) * 6 + (k)] = rw_now_ns(); } while (0)
#else
#define GS_STAMP(s, g, k) do { } while (0)
#define GS_T0(v) do { } while (0)
#define GS_T1(v, k) do { } while (0)
#define GS_TDECL do { } while (0)
#define GS_TFLUSH do { } while (0)
#endif

__device__ __forceinline__ float gs_sigmoid(float v) {       // SFU: ex2.approx + rcp.approx (absolute error ~1e-7)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return r;
}
__device__ __forceinline__ void gs_tmem_ld8(uint32_t (&v)[8], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ unsigned gs_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int UNITS>
__global__ void __launch_bounds__(GS_THREADS, 1)
gru_seq_kernel(const __grid_constant__ CUtensorMap map_h0, const __grid_constant__ CUtensorMap map_h1,
               const __grid_constant__ CUtensorMap map_w, const GruSeqParams p) {
    constexpr int BN = 3 * UNITS;                  // accumulator columns: r | z | n
    constexpr int W_TILE = BN * TC_BK * 2;         // [BN x 64] fp16
    constexpr int UT = UNITS / 2;                  // units per epilogue thread (two threads per accumulator row)
    // Back-to-back MMAs into ONE accumulator run at ~225 cycles each at these widths (N = 48 / 96: the dependent-issue latency, not
    // the 24 / 48 cycles of tensor work), 6 us for the 52 MMAs of a step.  The k-blocks therefore rotate over NACC independent
    // accumulators, which the epilogue adds up: 2 row blocks x NACC x ASTRIDE columns = all 512 TMEM columns.
    constexpr int ASTRIDE = UNITS == 16 ? 64 : 128, NACC = 256 / ASTRIDE;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t w_smem = (raw + 1023u) & ~1023u;                      // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t ring = w_smem + (uint32_t)p.KB * W_TILE;
    const uint32_t bars = ring + (uint32_t)p.nst * GS_A_TILE;            // full[8], empty[8], tfull[2], wfull, tmem slot
    const uint32_t full0 = bars, empty0 = bars + 64, tfull0 = bars + 128, wfull = bars + 144;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem_raw + (bars + 152 - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, rb0 = blockIdx.y * p.pp;
    const int n_tiles = gridDim.x;
    const int n_rb = (p.N + TC_BM - 1) / TC_BM;
    const int ng = (p.pp == 2 && rb0 + 1 < n_rb) ? 2 : 1;                // row blocks of this CTA (they alternate)
    const int NST = p.nst;
    GS_TDECL;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull0, 1); mbar_init(tfull0 + 8, 1); mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;                               // accumulator a of row block g: columns [256 g + ASTRIDE a, ... + BN)
    const int nacc = p.KB < NACC ? p.KB : NACC;

    if (warp == 0) {
        // ===== TMA producer: the resident W_hh^T slice once, then h_{t-1} of each row block for every step =====
        if (lane == 0) {
            mbar_expect_tx(wfull, (uint32_t)p.KB * W_TILE);
            for (int kb = 0; kb < p.KB; kb++) tma_load_2d(w_smem + (uint32_t)kb * W_TILE, &map_w, wfull, kb * TC_BK, tile * BN);
            int it = 0;
            for (int s = 1; s < p.T; s++) {                              // step 0: h_{-1} = 0, nothing to multiply
                const CUtensorMap *mh = ((s - 1) & 1) ? &map_h1 : &map_h0;
                for (int g = 0; g < ng; g++) {
                    // every CTA serving this row block has stored its columns of h_{s-1} (and so has finished reading h_{s-2})
                    const unsigned *cnt = p.cnt + rb0 + g;
                    const unsigned need = (unsigned)s * (unsigned)(n_tiles * 8);         // eight epilogue warps per CTA
                    GS_T0(tp);
                    if (gs_ld_acquire(cnt) < need) {
                        const unsigned long long t0 = rw_now_ns();
                        unsigned spins = 0;
                        while (gs_ld_acquire(cnt) < need)
                            if ((++spins & 1023u) == 0 && rw_now_ns() - t0 > RW_TIMEOUT_NS) __trap();
                    }
                    GS_T1(tp, 0);
                    GS_STAMP(s, g, 1);
                    asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy stores of the peers -> TMA reads
                    for (int kb = 0; kb < p.KB; kb++, it++) {
                        const int st = it % NST;
                        GS_T0(te);
                        rw_wait(empty0 + 8 * st, ((uint32_t)(it / NST) & 1u) ^ 1u);
                        GS_T1(te, 1);
                        mbar_expect_tx(full0 + 8 * st, GS_A_TILE);
                        tma_load_2d(ring + (uint32_t)st * GS_A_TILE, mh, full0 + 8 * st, kb * TC_BK, (rb0 + g) * TC_BM);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D_g[128 x BN] = h_{t-1}[row block g][128 x Kp] * W_slice^T, fp16 operands, fp32 accumulators in TMEM =====
        if (lane == 0) {
            // instruction descriptor: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N = BN, M = 128
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            rw_wait(wfull, 0);
            int it = 0;
            for (int s = 1; s < p.T; s++) {
                for (int g = 0; g < ng; g++) {
                    for (int kb = 0; kb < p.KB; kb++, it++) {
                        const int st = it % NST;
                        GS_T0(tf);
                        rw_wait(full0 + 8 * st, (uint32_t)(it / NST) & 1u);
                        GS_T1(tf, 2);
                        if (kb == 0) GS_STAMP(s, g, 2);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t da = umma_desc_sw128(ring + (uint32_t)st * GS_A_TILE);
                        const uint64_t db = umma_desc_sw128(w_smem + (uint32_t)kb * W_TILE);
#pragma unroll
                        for (int k4 = 0; k4 < TC_BK / 16; k4++) {
                            const uint64_t adv = (uint64_t)(k4 * 32 >> 4);
                            umma_bf16(tmem_base + (uint32_t)g * 256u + (uint32_t)(kb % NACC) * ASTRIDE, da + adv, db + adv, idesc,
                                      (kb >= NACC) || k4 != 0);                          // kind::f16; the descriptor selects fp16
                        }
                        umma_commit(empty0 + 8 * st);
                    }
                    umma_commit(tfull0 + 8 * g);
                    GS_STAMP(s, g, 3);
                }
            }
        }
    } else {
        // ===== epilogue: thread = utterance (accumulator row), UT hidden units; warps 2-5: the first UT units of the tile, warps
        //       6-9: the other UT (a warp may only read the TMEM lane quadrant warp % 4) =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int jb = tile * UNITS + half * UT;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * UT);
        float hst[2][UT];                                                 // h_{t-1} of this thread's units, per row block
#pragma unroll
        for (int e = 0; e < UT; e++) { hst[0][e] = 0.0f; hst[1][e] = 0.0f; }
        float4 xr4[UT / 4], xz4[UT / 4], xn4[UT / 4];
        // input projections of phase (s, g): they do not depend on the MMAs.  They are requested as soon as the previous phase
        // has released its h (the row-per-thread loads are uncoalesced: 32 lines per instruction, they must not sit in front
        // of the release) and arrive while the exchange / TMA ring / MMAs run
        auto request_x = [&](int s, int g) {
            const int t = p.reverse ? p.T - 1 - s : s;
            const int row = (rb0 + g) * TC_BM + q * 32 + lane;
            const float *xr = p.xp + ((size_t)t * p.N + (row < p.N ? row : 0)) * p.ldxp;
#pragma unroll
            for (int g4 = 0; g4 < UT / 4; g4++) {
                const int j = jb + 4 * g4;
                const int jc = j < p.H ? j : 0;                           // H % 4 == 0: a group of four is all in or all out
                xr4[g4] = __ldcs(reinterpret_cast<const float4 *>(xr + jc));
                xz4[g4] = __ldcs(reinterpret_cast<const float4 *>(xr + p.H + jc));
                xn4[g4] = __ldcs(reinterpret_cast<const float4 *>(xr + 2 * p.H + jc));
            }
        };
        request_x(0, 0);
        for (int s = 0; s < p.T; s++) {
            const int t = p.reverse ? p.T - 1 - s : s;
#pragma unroll
            for (int g = 0; g < 2; g++) {
                if (g >= ng) break;
                const int row = (rb0 + g) * TC_BM + q * 32 + lane;
                const bool live = row < p.N;
                GS_T0(ta);
                if (s > 0) {
                    rw_wait(tfull0 + 8 * g, (uint32_t)(s - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                GS_T1(ta, 3);
                if (warp == 2) GS_STAMP(s, g, 4);
                GS_T0(tc);
                // passes of 8 units: 24 accumulator registers live at a time
#pragma unroll
                for (int hp = 0; hp < UT / 8; hp++) {
                    uint32_t ar[8], az[8], an[8];
                    if (s > 0) {
                        const uint32_t ta0 = taddr + (uint32_t)g * 256u + (uint32_t)(8 * hp);
                        gs_tmem_ld8(ar, ta0);
                        gs_tmem_ld8(az, ta0 + UNITS);
                        gs_tmem_ld8(an, ta0 + 2 * UNITS);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        for (int a = 1; a < nacc; a++) {                  // the other partial accumulators
                            uint32_t br[8], bz[8], bn[8];
                            gs_tmem_ld8(br, ta0 + (uint32_t)a * ASTRIDE);
                            gs_tmem_ld8(bz, ta0 + (uint32_t)a * ASTRIDE + UNITS);
                            gs_tmem_ld8(bn, ta0 + (uint32_t)a * ASTRIDE + 2 * UNITS);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int e = 0; e < 8; e++) {
                                ar[e] = __float_as_uint(__uint_as_float(ar[e]) + __uint_as_float(br[e]));
                                az[e] = __float_as_uint(__uint_as_float(az[e]) + __uint_as_float(bz[e]));
                                an[e] = __float_as_uint(__uint_as_float(an[e]) + __uint_as_float(bn[e]));
                            }
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; e++) { ar[e] = 0u; az[e] = 0u; an[e] = 0u; }   // h_0 = 0: hh = b_hh exactly
                    }
#pragma unroll
                    for (int gq = 0; gq < 2; gq++) {
                        const int g4 = 2 * hp + gq;
                        const int j = jb + 4 * g4;
                        const int jc = j < p.H ? j : 0;
                        const float4 br4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + jc));
                        const float4 bz4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + p.H + jc));
                        const float4 bn4 = __ldg(reinterpret_cast<const float4 *>(p.b_hh + 2 * p.H + jc));
                        const float xrv[4] = {xr4[g4].x, xr4[g4].y, xr4[g4].z, xr4[g4].w}, xzv[4] = {xz4[g4].x, xz4[g4].y, xz4[g4].z, xz4[g4].w};
                        const float xnv[4] = {xn4[g4].x, xn4[g4].y, xn4[g4].z, xn4[g4].w};
                        const float brv[4] = {br4.x, br4.y, br4.z, br4.w}, bzv[4] = {bz4.x, bz4.y, bz4.z, bz4.w};
                        const float bnv[4] = {bn4.x, bn4.y, bn4.z, bn4.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float gr = __uint_as_float(ar[4 * gq + e]) + brv[e];
                            const float gz = __uint_as_float(az[4 * gq + e]) + bzv[e];
                            const float gn = __uint_as_float(an[4 * gq + e]) + bnv[e];
                            const float r = gs_sigmoid(xrv[e] + gr);
                            const float z = gs_sigmoid(xzv[e] + gz);
                            const float nn = rw_tanh(xnv[e] + r * gn);
                            hst[g][4 * g4 + e] = (1.0f - z) * nn + z * hst[g][4 * g4 + e];
                        }
                    }
                }
                if (s > 0) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM reads ordered before the release below
                // fp16 plane first (what the peers wait for), release, and only then the fp32 layer output
                if (live) {
                    __half *prow = ((s & 1) ? p.plane[1] : p.plane[0]) + (size_t)row * p.Kp;
#pragma unroll
                    for (int g8 = 0; g8 < UT / 8; g8++) {
                        const int j = jb + 8 * g8;
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const __half2 h2 = __floats2half2_rn(hst[g][8 * g8 + 2 * e], hst[g][8 * g8 + 2 * e + 1]);
                            w[e] = *reinterpret_cast<const uint32_t *>(&h2);
                        }
                        if (j + 8 <= p.H) *reinterpret_cast<uint4 *>(prow + j) = make_uint4(w[0], w[1], w[2], w[3]);
                        else if (j + 4 <= p.H) *reinterpret_cast<uint2 *>(prow + j) = make_uint2(w[0], w[1]);   // H % 4 == 0
                    }
                }
                GS_T1(tc, 4);
                if (warp == 2) GS_STAMP(s + 1, g, 5);
                if (s + 1 < p.T) {
                    // release this warp's 32 rows x UT columns of h_s (and its TMEM reads) to the row block's TMA producers
                    GS_T0(tr);
                    __syncwarp();
                    if (lane == 0) { __threadfence(); atomicAdd(p.cnt + rb0 + g, 1u); }
                    if (warp == 2) GS_STAMP(s + 1, g, 0);
                    GS_T1(tr, 5);
                }
                if (live) {
                    float *orow = p.out + ((size_t)t * p.N + row) * p.ldo;
#pragma unroll
                    for (int g4 = 0; g4 < UT / 4; g4++) {
                        const int j = jb + 4 * g4;
                        if (j >= p.H) break;
                        __stcs(reinterpret_cast<float4 *>(orow + j), make_float4(hst[g][4 * g4], hst[g][4 * g4 + 1], hst[g][4 * g4 + 2], hst[g][4 * g4 + 3]));
                    }
                }
                // the next phase's projections (the x registers are free since the gate math)
                if (g + 1 < ng) request_x(s, g + 1);
                else if (s + 1 < p.T) request_x(s + 1, 0);
            }
        }
    }
    GS_TFLUSH;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// W_hh[H, 3H] (reference layout [in, out], gates r | z | n) -> permuted W^T as one fp16 plane [tiles * 3 units, Kp]:
// row tile * 3 units + g * units + u  =  column g * H + tile * units + u of W_hh (zero beyond H)
__global__ void gru_seq_perm_kernel(const float *__restrict__ w, int H, int Kp, int rows, int units, __half *__restrict__ wt) {
    const size_t total = (size_t)rows * Kp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int prow = (int)(i / Kp), k = (int)(i % Kp);
        const int tile = prow / (3 * units), rem = prow - tile * 3 * units, g = rem / units, u = rem - g * units;
        const int j = tile * units + u;
        wt[i] = __float2half_rn((k < H && j < H) ? w[(size_t)k * 3 * H + (size_t)g * H + j] : 0.0f);
    }
}

static int gs_kp(int H) { return ceil_div(H, TC_BK) * TC_BK; }
static int gs_rows(int H, int units) { return ceil_div(H, units) * 3 * units; }
static size_t gs_w_bytes(int H, int units) { return align_up((size_t)gs_rows(H, units) * gs_kp(H) * 2, 1024); }
static size_t gs_plane_bytes(int N, int H) { return align_up((size_t)N * gs_kp(H) * 2, 1024); }
static int gs_stages(int H, int units) {
    const int left = GS_MAX_SMEM - 1024 - 256 - (gs_kp(H) / TC_BK) * 3 * units * TC_BK * 2;
    const int n = left / GS_A_TILE;
    return n > GS_MAX_STAGES ? GS_MAX_STAGES : n;
}

// Shape of the launch: units per CTA and row blocks per CTA.  Default: two row blocks per CTA where the batch has them and
// 16-unit slices (half the W footprint, twice the ring); GASR_GRU_UNITS / GASR_GRU_PP override.  Every CTA of both directions
// of a layer must be resident at once (they wait for each other's h columns).
static bool gs_shape(const gasr_ctx *ctx, int N, int H, int &units, int &pp) {
    const int n_rb = ceil_div(N, TC_BM);
    pp = ctx->opt.gru_pp == 1 ? 1 : (ctx->opt.gru_pp == 2 ? 2 : (n_rb >= 2 ? 2 : 1));
    units = ctx->opt.gru_units == 32 ? 32 : (ctx->opt.gru_units == 16 ? 16 : 16);
    for (int attempt = 0; attempt < 2; attempt++) {
        if (gs_stages(H, units) >= 2 && 2 * ceil_div(H, units) * ceil_div(n_rb, pp) <= ctx->sm_count) return true;
        if (ctx->opt.gru_units == 0 && units == 16) units = 32; else break;    // fewer, wider CTAs
    }
    return false;
}

bool gru_seq_supported(const gasr_ctx *ctx, int T, int N, int H, int ldxp, int ldo, int col0) {
    if (!(T >= 2 && N >= 1 && H >= 32 && H % 4 == 0 && ldxp % 4 == 0 && ldo % 4 == 0 && col0 % 4 == 0)) return false;
    int units, pp;
    return gs_shape(ctx, N, H, units, pp);
}

size_t gru_seq_ws_bytes(int N, int H) { return gs_w_bytes(H, 16) + gs_w_bytes(H, 32) + 2 * gs_plane_bytes(N, H) + 1024; }

// One launch for all T steps.  ws: gru_seq_ws_bytes(N, H) of device memory (weights plane, two h planes, counters).
int launch_gru_seq(gasr_ctx *ctx, const RnnLayerArgs &a, void *ws, cudaStream_t st) {
    const int H = a.H, N = a.N, Kp = gs_kp(H);
    int units = 16, pp = 1;
    GASR_CHECK(gs_shape(ctx, N, H, units, pp), "gru_seq: unsupported shape N=%d H=%d", N, H);
    const int rows = gs_rows(H, units);
    unsigned char *base = static_cast<unsigned char *>(ws);
    const size_t wbytes = gs_w_bytes(H, 16) > gs_w_bytes(H, 32) ? gs_w_bytes(H, 16) : gs_w_bytes(H, 32);
    __half *wt = reinterpret_cast<__half *>(base);
    __half *pl0 = reinterpret_cast<__half *>(base + wbytes);
    __half *pl1 = reinterpret_cast<__half *>(base + wbytes + gs_plane_bytes(N, H));
    unsigned *cnt = reinterpret_cast<unsigned *>(base + wbytes + 2 * gs_plane_bytes(N, H));
    {
        const size_t total = (size_t)rows * Kp;
        int blocks = (int)((total + 255) / 256);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        gru_seq_perm_kernel<<<blocks, 256, 0, st>>>(a.w_hh, H, Kp, rows, units, wt);
        GASR_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    // both planes zero (the K padding stays zero for the whole run), counters zero
    GASR_CUDA(cudaMemsetAsync(pl0, 0, 2 * gs_plane_bytes(N, H) + 1024, st));
    CUtensorMap mh0, mh1, mw;
    GASR_TRY(tc_make_map(&mh0, pl0, N, Kp, TC_BM));
    GASR_TRY(tc_make_map(&mh1, pl1, N, Kp, TC_BM));
    GASR_TRY(tc_make_map(&mw, wt, rows, Kp, 3 * units));
    GruSeqParams p;
    p.T = a.T; p.N = N; p.H = H; p.Kp = Kp; p.KB = Kp / TC_BK; p.reverse = a.reverse; p.nst = gs_stages(H, units); p.pp = pp;
    p.xp = a.xproj; p.ldxp = a.ldxp; p.b_hh = a.b_hh; p.out = a.out + a.col0; p.ldo = a.ldo;
    p.plane[0] = pl0; p.plane[1] = pl1; p.cnt = cnt; p.trace = nullptr; p.stamps = nullptr;
#ifdef GASR_RW_TRACE
    if (getenv("GASR_GS_TRACE")) {
        static unsigned long long *th[2] = {nullptr, nullptr}, *td[2] = {nullptr, nullptr};
        const int sl = a.reverse ? 1 : 0;
        if (!th[sl]) {
            GASR_CUDA(cudaHostAlloc((void **)&th[sl], 64, cudaHostAllocMapped));
            GASR_CUDA(cudaHostGetDevicePointer((void **)&td[sl], th[sl], 0));
        } else {
            GASR_CUDA(cudaStreamSynchronize(st));                         // the previous launch of this direction
            fprintf(stderr, "[gru_seq trace] dir %d units %d pp %d stages %d, cycles per step of CTA (1,0): wait-peers %.0f ring-free %.0f full-stage %.0f accumulator %.0f math+planes %.0f release %.0f\n",
                    sl, units, pp, p.nst, th[sl][0] / (double)a.T, th[sl][1] / (double)a.T, th[sl][2] / (double)a.T, th[sl][3] / (double)a.T, th[sl][4] / (double)a.T, th[sl][5] / (double)a.T);
        }
        for (int i = 0; i < 8; i++) th[sl][i] = 0;
        p.trace = td[sl];
        if (sl == 0) {
            static unsigned long long *sh = nullptr, *sd = nullptr;
            const int nc = ceil_div(H, units) * ceil_div(ceil_div(N, TC_BM), pp), n_rb = ceil_div(N, TC_BM);
            constexpr int CAP = 256 * 4 * 2 * 6;
            if (!sh) {
                GASR_CUDA(cudaHostAlloc((void **)&sh, sizeof(unsigned long long) * CAP, cudaHostAllocMapped));
                GASR_CUDA(cudaHostGetDevicePointer((void **)&sd, sh, 0));
            } else {
                // per row block (the row blocks only meet through their own counter): spread of every stamp over the CTAs serving it,
                // relative to the earliest "gate math done" of the step before
                const char *names[6] = {"released", "seen by producer", "first h tile", "accumulator done", "epilogue starts", "math (prev) done"};
                for (int rb = 0; rb < n_rb && rb < 2; rb++)
                    for (int q = 1; q < 3; q++) {
                        unsigned long long base = ~0ull;
                        auto at = [&](int c, int k) { return sh[((((size_t)c) * 4 + q) * 2 + (pp == 2 ? rb : 0)) * 6 + k]; };
                        auto serves = [&](int c) { return pp == 2 ? true : (c / ceil_div(H, units)) == rb; };
                        for (int c = 0; c < nc; c++) if (serves(c) && at(c, 5) && at(c, 5) < base) base = at(c, 5);
                        for (int k : {5, 0, 1, 2, 3, 4}) {
                            unsigned long long mn = ~0ull, mx = 0, sum = 0; int n = 0;
                            for (int c = 0; c < nc; c++) if (serves(c)) { const unsigned long long v = at(c, k) - base; mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; n++; }
                            fprintf(stderr, "[gru_seq stamps] row block %d step %d %-18s min %6llu mean %6llu max %6llu ns\n", rb, 500 + q, names[k], mn, sum / (n ? n : 1), mx);
                        }
                    }
            }
            for (int i = 0; i < CAP; i++) sh[i] = 0;
            p.stamps = sd;
        }
    }
#endif
    const size_t smem = 1024 + (size_t)p.KB * 3 * units * TC_BK * 2 + (size_t)p.nst * GS_A_TILE + 256;
    if (!(ctx->attr_mask & 32768u)) {
        GASR_CUDA(cudaFuncSetAttribute(gru_seq_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_MAX_SMEM));
        GASR_CUDA(cudaFuncSetAttribute(gru_seq_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_MAX_SMEM));
        ctx->attr_mask |= 32768u;
    }
    dim3 grid(ceil_div(H, units), ceil_div(ceil_div(N, TC_BM), pp));
    if (units == 16) gru_seq_kernel<16><<<grid, GS_THREADS, smem, st>>>(mh0, mh1, mw, p);
    else gru_seq_kernel<32><<<grid, GS_THREADS, smem, st>>>(mh0, mh1, mw, p);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

}  // namespace gasr
