// rnn_stream.cu -- persistent, multi-layer tanh recurrence (reference RNN.cu:9-30 / RNN_Cell.cu:5-13,65-74) built
// for step LATENCY: the whole layer stack runs as ONE launch for all T steps, layers coupled to the projection GEMMs
// through progress counters in HBM instead of kernel boundaries.
//
// Decomposition: one thread-block cluster of CS = H/64 CTAs per (layer, group of 16 utterances).  CTA r owns output
// columns [64r, 64r+64) of W_hh, resident in REGISTERS for the whole sequence as bf16 hi/lo MMA A-fragments of
// W_hh^T (warp (ch, kh): columns 16ch..16ch+15, K half kh; 128 registers per thread at H = 512).  h_{t-1} of the
// group lives in every CTA's shared memory as bf16 hi/lo planes [utterance][k] and is the MMA B operand (ldmatrix);
// one step is 3 x H/16 mma.sync.m16n8k16 per 16 columns (hi*hi + hi*lo + lo*hi, fp32 accumulate: fp32-grade).
//
// What makes it fast (measured on B200, tools/ubench/dsmem_exchange.cu):
//   * the all-gather of h_t inside the cluster is st.async (16-byte remote stores that complete_tx on the
//     DESTINATION CTA's mbarrier): no cluster barrier, no membar -- the consumer just waits for 16 KB of transactions;
//   * the 16 utterances are two independent sub-batches of 8 (MMA N = 8): while sub-batch A's h is in flight through
//     DSMEM (~1000 cycles), the tensor cores work on sub-batch B, so the exchange latency is hidden behind math;
//   * no CTA-wide barrier in the step: the only synchronisation is a 64-thread named barrier between the two
//     K-half warps of a column block, plus the mbarrier wait.
//   * progress flags: the projection GEMM's per-block counters are read with a plain volatile load issued one step
//     before the value is needed (no acquire stall; the xproj rows themselves are read with ld.global.cg, i.e. from
//     L2, the point of coherence); this layer's progress is published once per block by whichever warp finished the
//     block last (fence + one atomic per CTA and block).
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "rnn_stream.cuh"

namespace cg = cooperative_groups;

namespace gasr {

constexpr int RS_NB = 16;                  // utterances per cluster
constexpr int RS_SUB = 8;                  // utterances per sub-batch (MMA N)
constexpr int RS_HC = 64;                  // hidden columns per CTA
constexpr int RS_MATH_WARPS = 8;
constexpr int RS_SIGNAL_BLOCKS = 4;        // progress is published (and the warps are counted) once per this many blocks
constexpr int RS_THREADS = 32 * RS_MATH_WARPS;
constexpr unsigned long long RS_TIMEOUT_NS = 2000000000ull;   // watchdog for every spin on a flag

__device__ __forceinline__ uint32_t rs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t rs_mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void rs_mbar_init(uint32_t bar, uint32_t cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt));
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint32_t bar, uint32_t tx) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
}
__device__ __forceinline__ bool rs_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void rs_st_async_v4(uint32_t raddr, const uint4 &v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ void rs_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void rs_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void rs_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned rs_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long rs_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// tanh(x) = 1 - 2 / (2^(2x log2 e) + 1) on the SFU (ex2.approx, rcp.approx): absolute error ~1e-7 over the whole range
// (the fp32 libm tanhf costs 32 instructions per value on the step's critical path; the parity bar is 1e-4 absolute)
__device__ __forceinline__ float rs_tanh(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ void rs_split(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ void rs_split_pair(float x, float y, uint32_t &hi, uint32_t &lo) {
    __nv_bfloat16 xh, xl, yh, yl;
    rs_split(x, xh, xl);
    rs_split(y, yh, yl);
    const __nv_bfloat162 h2 = __halves2bfloat162(xh, yh), l2 = __halves2bfloat162(xl, yl);
    hi = *reinterpret_cast<const uint32_t *>(&h2);
    lo = *reinterpret_cast<const uint32_t *>(&l2);
}

// named barriers: 0 = __syncthreads, 1..4 = K-half pairs
constexpr int RS_BAR_PAIR0 = 1;

template <int H>
struct RsLayout {
    static constexpr int HSB = H * 2 + 16;                 // bytes of one bf16 h row (+16: ldmatrix rows hit distinct banks)
    static constexpr int PLANE = RS_SUB * HSB;             // one plane (hi or lo) of one sub-batch buffer
    static constexpr int HBUF = 2 * PLANE;                 // hi + lo
    static constexpr int OFF_H = 0;                        // [sub 2][parity 2][plane 2][8][HSB]
    static constexpr int OFF_XCH = OFF_H + 4 * HBUF;       // [sub 2][ch 4][kh 2][32] float2
    static constexpr int OFF_STG = OFF_XCH + 2 * 4 * 2 * 32 * 8;   // [warp 8][plane 2][utt 8][16 B]
    static constexpr int OFF_BAR = OFF_STG + RS_MATH_WARPS * 256;  // mbar[sub 2][parity 2], then control words
    static constexpr int OFF_CTL = OFF_BAR + 4 * 8;        // volatile int: [0], [1] math-warp arrivals by step parity, [2] abort
    static constexpr int BYTES = OFF_CTL + 16;
    static constexpr uint32_t TX = 2u * RS_SUB * H * 2u;   // bytes one step delivers into one sub-batch buffer
};

template <int H>
__global__ void __launch_bounds__(RS_THREADS, 1) rnn_stream_kernel(const RnnStreamParams p) {
    using LT = RsLayout<H>;
    constexpr int CS = H / RS_HC;
    constexpr int KSW = H / 32;                          // k-steps (of 16) per warp: half of K
    extern __shared__ __align__(128) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / CS;
    const int layer = cid / p.groups, group = cid % p.groups;
    const int n0 = group * RS_NB;
    const int colbase = rank * RS_HC;
    const RnnStreamLayer &L = p.layer[layer];
    volatile int *ctl = reinterpret_cast<volatile int *>(smem + LT::OFF_CTL);
    const uint32_t sbase = rs_smem_u32(smem);
    const int T = p.T, N = p.N, fpb = p.frames_per_block;

    // ---- init: zero h buffers (h_0 = 0, RNN.h:16-17), mbarriers armed for their first use ----------------------------
    for (int i = tid; i < (4 * LT::HBUF) / 16; i += RS_THREADS) reinterpret_cast<uint4 *>(smem + LT::OFF_H)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int b = 0; b < 4; b++) rs_mbar_init(sbase + LT::OFF_BAR + 8 * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ctl[0] = 0; ctl[1] = 0; ctl[2] = 0;
    }
    __syncthreads();
    if (tid == 0)
        for (int b = 0; b < 4; b++) rs_mbar_expect_tx(sbase + LT::OFF_BAR + 8 * b, LT::TX);

    // =================================== math warps ===================================
    const int ch = warp & 3, kh = warp >> 2;
    // resident weights: A fragments of W_hh^T, rows = columns colbase+16ch+{g, g+8}, k = kh*H/2 + 16ks + {2tg, 2tg+1, +8, +9}
    uint32_t ahi[KSW][4], alo[KSW][4];
    {
        const float *wc = L.w_hh + colbase + 16 * ch + g;
#pragma unroll
        for (int ks = 0; ks < KSW; ks++) {
            const int k = kh * (H / 2) + 16 * ks + 2 * tg;
            rs_split_pair(__ldg(wc + (size_t)k * H), __ldg(wc + (size_t)(k + 1) * H), ahi[ks][0], alo[ks][0]);
            rs_split_pair(__ldg(wc + (size_t)k * H + 8), __ldg(wc + (size_t)(k + 1) * H + 8), ahi[ks][1], alo[ks][1]);
            rs_split_pair(__ldg(wc + (size_t)(k + 8) * H), __ldg(wc + (size_t)(k + 9) * H), ahi[ks][2], alo[ks][2]);
            rs_split_pair(__ldg(wc + (size_t)(k + 8) * H + 8), __ldg(wc + (size_t)(k + 9) * H + 8), ahi[ks][3], alo[ks][3]);
        }
    }
    // ldmatrix lane address inside a sub-batch buffer: matrix m = lane >> 3 covers k-offset 8m, row = utterance lane & 7
    const uint32_t lm_off = (uint32_t)((lane & 7) * LT::HSB + (kh * (H / 2) + 8 * (lane >> 3)) * 2);
    // this thread finalises column jcol for utterances (2tg, 2tg+1) of each sub-batch
    const int jl = 16 * ch + 8 * kh + g;                 // column inside the CTA slice
    const int jcol = colbase + jl;
    // CTA-uniform bases + 32-bit element offsets that advance by one frame per step (T*N*ld < 2^31 is checked on the host)
    const float *const xp_base = L.xproj;
    float *const out_base = L.out;
    unsigned *const hdone = L.h_done;
    const int xstep = N * L.ldxp, ostep = N * L.ldo, pstep = N * L.ldp;
    const int nA0 = n0 + 2 * tg, nA1 = n0 + RS_SUB + 2 * tg;             // first utterance of this thread in sub-batch 0 / 1
    const unsigned vmask = (nA0 < N ? 1u : 0u) | (nA0 + 1 < N ? 2u : 0u) | (nA1 < N ? 4u : 0u) | (nA1 + 1 < N ? 8u : 0u);
    int xoff = nA0 * L.ldxp + jcol;                      // + ldxp: second utterance; + 8 ldxp: sub-batch 1
    int ooff = nA0 * L.ldo + jcol;
    // bf16 plane chunk of this lane: (plane cpl, utterance cu) -> 8 columns starting at colbase + 16ch + 8kh
    const int cpl = (lane >> 3) & 1, cu = lane & 7;
    __nv_bfloat16 *const plane_base = cpl ? L.out_lo : L.out_hi;
    int poff = (n0 + cu) * L.ldp + colbase + 16 * ch + 8 * kh;           // + 8 ldp: sub-batch 1
    const unsigned pmask = (lane < 16 && plane_base != nullptr) ? ((n0 + cu < N ? 1u : 0u) | (n0 + RS_SUB + cu < N ? 2u : 0u)) : 0u;
    const int ldxp8 = RS_SUB * L.ldxp, ldo8 = RS_SUB * L.ldo, ldp8 = RS_SUB * L.ldp, ldxp1 = L.ldxp, ldo1 = L.ldo;

    float xn00 = 0.f, xn01 = 0.f, xn10 = 0.f, xn11 = 0.f;                // prefetched xproj of the next step: [sub][utt]
    auto load_xp = [&]() {                               // frame = the one xoff points at
        if (vmask & 1u) xn00 = __ldcg(xp_base + xoff);
        if (vmask & 2u) xn01 = __ldcg(xp_base + xoff + ldxp1);
        if (vmask & 4u) xn10 = __ldcg(xp_base + xoff + ldxp8);
        if (vmask & 8u) xn11 = __ldcg(xp_base + xoff + ldxp8 + ldxp1);
        xoff += xstep;
    };
    // projection progress: blocks [0, ready_blocks) are known complete.  The counter of the next block is sampled with a
    // volatile load one step before it is needed (flag_next); a miss falls back to spinning on fresh loads.
    const volatile unsigned *xr = L.xp_ready;
    const int nblocks = (T + fpb - 1) / fpb;
    const int fsh = (fpb & (fpb - 1)) == 0 ? __ffs(fpb) - 1 : -1;        // frames per block is a power of two in the pipeline
    int ready_blocks = xr == nullptr ? nblocks : 0;
    unsigned flag_next = 0;
    auto sample_flag = [&]() { if (lane == 0) flag_next = xr[ready_blocks]; };
    auto advance_blocks = [&](int t) {                   // make sure the xproj rows of frame t are complete
        const int before = ready_blocks;
        // opportunistic: last step's sample of the next block's counter
        if (__shfl_sync(0xffffffffu, flag_next, 0) >= (unsigned)p.xp_need) ready_blocks++;
        const int need = (fsh >= 0 ? (t >> fsh) : t / fpb) + 1;
        while (ready_blocks < need) {                    // slow path: spin on fresh loads
            const unsigned long long t0 = rs_now_ns();
            unsigned v;
            do {
                v = 0;
                if (lane == 0) v = xr[ready_blocks];
                v = __shfl_sync(0xffffffffu, v, 0);
                if (v < (unsigned)p.xp_need) {
                    if (ctl[2] || (p.abort && *p.abort) || rs_now_ns() - t0 > RS_TIMEOUT_NS) {
                        if (p.abort) *p.abort = 1u; ctl[2] = 1; if (p.error) { *reinterpret_cast<volatile int *>(p.error) = 1; __threadfence_system(); } v = 0xffffffffu; }
                    else __nanosleep(40);
                }
            } while (v < (unsigned)p.xp_need);
            ready_blocks++;
        }
        if (ready_blocks != before) __threadfence();     // acquire side of the counters: the xproj rows are read after this fence
        flag_next = 0;
        if (ready_blocks < nblocks) sample_flag();       // issued BEFORE this step's xproj loads
    };
    unsigned char *const stg = smem + LT::OFF_STG + warp * 256;
    __nv_bfloat16 *const sg = reinterpret_cast<__nv_bfloat16 *>(stg) + (2 * tg) * 8 + g;    // [plane 64][utt 8][col]
    const uint4 *const chunk_src = reinterpret_cast<const uint4 *>(stg + (lane & 15) * 16);
    float2 *const xch_mine = reinterpret_cast<float2 *>(smem + LT::OFF_XCH) + (ch * 2 + kh) * 32 + lane;
    const float2 *const xch_peer = reinterpret_cast<const float2 *>(smem + LT::OFF_XCH) + (ch * 2 + (kh ^ 1)) * 32 + lane;
    // remote targets of this lane's chunk (sub-batch 0, parity 0): + HBUF per parity, + 2 HBUF per sub-batch
    const uint32_t dst_local0 = sbase + LT::OFF_H + cpl * LT::PLANE + cu * LT::HSB + (colbase + 16 * ch + 8 * kh) * 2;
    const uint32_t bar0 = sbase + LT::OFF_BAR;
    const int r0 = (lane >> 4) * (CS / 2);               // lanes 0-15 serve the first half of the cluster, 16-31 the second
    uint32_t phase_bits = 0;                             // bit (sub*2+par): parity of the next phase to wait for
    int pend_hi = -1;                                    // lane 0: highest block whose completion this warp still has to publish
    int fcnt = 0, blk = 0;                               // frames finished in the current block, current block

    cluster.sync();
    if (p.started != nullptr && tid == 0) {
        // residency handshake: the host launches the other persistent kernels only after every CTA of this grid runs
        if (atomicAdd(p.started, 1u) == gridDim.x - 1) { *p.host_go = p.epoch; __threadfence_system(); }
    }
    if (ready_blocks < nblocks) advance_blocks(0);
    load_xp();

    for (int s = 0; s < T; s++) {
        const int par = s & 1;
        const bool more = s + 1 < T;
        const float xc00 = xn00, xc01 = xn01, xc10 = xn10, xc11 = xn11;
        if (more) {
            if (ready_blocks < nblocks) advance_blocks(s + 1);
            load_xp();
        }
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            const int bi = sub * 2 + par;
            const uint32_t bar = bar0 + 8 * bi;
            if (s > 0) {
                const uint32_t ph = (phase_bits >> bi) & 1u;
                if (!rs_mbar_try(bar, ph)) {
                    const unsigned long long t0 = rs_now_ns();
                    int spins = 0;
                    while (!rs_mbar_try(bar, ph))
                        if (ctl[2] || ((++spins & 255) == 0 && ((p.abort && *p.abort) || rs_now_ns() - t0 > RS_TIMEOUT_NS))) {
                            ctl[2] = 1; if (p.abort) *p.abort = 1u; break;
                        }
                }
                phase_bits ^= 1u << bi;
                if (tid == 0) rs_mbar_expect_tx(bar, LT::TX);          // re-arm for the use two steps later
            }
            // ---- h_{s-1}[sub] * W_hh slice over this warp's K half -------------------------------------------------
            const uint32_t hb = sbase + LT::OFF_H + bi * LT::HBUF + lm_off;
            float cm0[4] = {0.f, 0.f, 0.f, 0.f}, cm1[4] = {0.f, 0.f, 0.f, 0.f};     // hi*hi (even / odd k-steps)
            float cs0[4] = {0.f, 0.f, 0.f, 0.f}, cs1[4] = {0.f, 0.f, 0.f, 0.f};     // hi*lo + lo*hi
#pragma unroll
            for (int kp = 0; kp < KSW / 2; kp++) {
                uint32_t bh[4], bl[4];
                rs_ldmatrix_x4(bh, hb + kp * 64);
                rs_ldmatrix_x4(bl, hb + LT::PLANE + kp * 64);
                rs_mma(cm0, ahi[2 * kp], bh[0], bh[1]);
                rs_mma(cm1, ahi[2 * kp + 1], bh[2], bh[3]);
                rs_mma(cs0, ahi[2 * kp], bl[0], bl[1]);
                rs_mma(cs1, ahi[2 * kp + 1], bl[2], bl[3]);
                rs_mma(cs0, alo[2 * kp], bh[0], bh[1]);
                rs_mma(cs1, alo[2 * kp + 1], bh[2], bh[3]);
            }
            float c[4];
#pragma unroll
            for (int i = 0; i < 4; i++) c[i] = (cm0[i] + cm1[i]) + (cs0[i] + cs1[i]);
            // ---- combine the two K halves: kh = 0 keeps rows g (c0, c1), kh = 1 keeps rows g + 8 (c2, c3) ----------
            xch_mine[sub * 256] = kh == 0 ? make_float2(c[2], c[3]) : make_float2(c[0], c[1]);
            rs_bar_sync(RS_BAR_PAIR0 + ch, 64);
            const float2 o = xch_peer[sub * 256];
            const float a0 = kh == 0 ? c[0] + o.x : o.x + c[2];
            const float a1 = kh == 0 ? c[1] + o.y : o.y + c[3];
            const float v0 = rs_tanh((sub ? xc10 : xc00) + a0), v1 = rs_tanh((sub ? xc11 : xc01) + a1);
            if (sub == 0 && pend_hi >= 0) {        // (lane 0 of one warp per CTA, once per RS_SIGNAL_BLOCKS blocks)
                __threadfence();
                for (int b2 = pend_hi - (RS_SIGNAL_BLOCKS - 1); b2 <= pend_hi; b2++) atomicAdd(hdone + b2, 1u);
                pend_hi = -1;
            }
            // ---- publish: fp32 output, bf16 planes, DSMEM all-gather ------------------------------------------------
            if (out_base != nullptr) {
                if (vmask & (sub ? 4u : 1u)) out_base[ooff + sub * ldo8] = v0;
                if (vmask & (sub ? 8u : 2u)) out_base[ooff + sub * ldo8 + ldo1] = v1;
            }
            if (more || plane_base != nullptr) {
                __nv_bfloat16 h0, l0, h1, l1;
                rs_split(v0, h0, l0);
                rs_split(v1, h1, l1);
                __syncwarp();                                    // previous sub-step's chunk reads are done
                sg[0] = h0; sg[8] = h1; sg[64] = l0; sg[72] = l1;
                __syncwarp();
                const uint4 chunk = *chunk_src;                  // (plane, utterance): 8 columns, 16 bytes
                if (pmask & (1u << sub)) *reinterpret_cast<uint4 *>(plane_base + poff + sub * ldp8) = chunk;
                if (more) {
                    const uint32_t nb_off = (uint32_t)(sub * 2 + (par ^ 1));
                    const uint32_t dst_local = dst_local0 + nb_off * LT::HBUF, bar_local = bar0 + 8 * nb_off;
#pragma unroll
                    for (int r = 0; r < CS / 2; r++)
                        rs_st_async_v4(rs_mapa(dst_local, r0 + r), chunk, rs_mapa(bar_local, r0 + r));
                }
            }
        }
        ooff += ostep;
        poff += pstep;
        // progress: the warp that completes a block last owes the consumers a fence + count for it.  The fence waits for
        // the CTA's outstanding stores, so it is DEFERRED to the middle of the next step (pend_hi), when those stores have
        // long been acknowledged, and issued once per RS_SIGNAL_BLOCKS blocks.
        if (++fcnt == fpb || !more) {
            if (hdone != nullptr && ((blk + 1) % RS_SIGNAL_BLOCKS == 0 || !more)) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    const int grp_par = (blk / RS_SIGNAL_BLOCKS) & 1;
                    const int last = (atomicAdd(const_cast<int *>(ctl + grp_par), 1) & (RS_MATH_WARPS - 1)) == RS_MATH_WARPS - 1;
                    if (last) {
                        pend_hi = blk;
                        if (!more) {
                            __threadfence();
                            for (int b2 = blk - blk % RS_SIGNAL_BLOCKS; b2 <= blk; b2++) atomicAdd(hdone + b2, 1u);
                            pend_hi = -1;
                        }
                    }
                }
            }
            fcnt = 0;
            blk++;
        }
    }
    cluster.sync();
}

bool rnn_stream_supported(const gasr_ctx *ctx, int H, int N, int L) {
    if (!ctx->cluster_ok) return false;
    if (H != 512 && H != 256 && H != 128) return false;
    if (N < 1 || L < 1 || L > RS_MAX_LAYERS) return false;
    const int groups = ceil_div(N, RS_NB);
    return groups * L <= 16;            // clusters of 8 that are co-resident on a B200 (2 per GPC)
}

template <int H>
static int launch_rs(gasr_ctx *ctx, const RnnStreamParams &p, cudaStream_t st) {
    size_t smem = RsLayout<H>::BYTES;
    if (smem < 200 * 1024) smem = 200 * 1024;   // one CTA per SM, nothing else co-resident: the step is latency critical
    // set once per context: cudaFuncSetAttribute may wait for running kernels, which would stall the streaming
    // pipeline (its kernels wait for each other)
    const unsigned bit = H == 512 ? 1u : H == 256 ? 2u : 4u;
    if (!(ctx->attr_mask & bit)) {
        GASR_CUDA(cudaFuncSetAttribute(rnn_stream_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_mask |= bit;
    }
    cudaLaunchConfig_t cfg = {};
    constexpr int CS = H / RS_HC;
    cfg.gridDim = dim3(p.L * p.groups * CS);
    cfg.blockDim = dim3(RS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GASR_CUDA(cudaLaunchKernelEx(&cfg, rnn_stream_kernel<H>, p));
    return GASR_OK;
}

// How many clusters of the recurrence kernel the device can hold at once (cudaOccupancyMaxActiveClusters): the streaming
// pipeline needs ALL of its clusters resident next to the GEMM and decoder CTAs, so gasr_asr_create asks before it picks the mode.
template <int H>
static int rs_max_clusters(int *out) {
    size_t smem = RsLayout<H>::BYTES;
    if (smem < 200 * 1024) smem = 200 * 1024;
    if (cudaFuncSetAttribute(rnn_stream_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return GASR_ERR_CUDA;
    cudaLaunchConfig_t cfg = {};
    constexpr int CS = H / RS_HC;
    cfg.gridDim = dim3(CS); cfg.blockDim = dim3(RS_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(out, rnn_stream_kernel<H>, &cfg) == cudaSuccess ? GASR_OK : GASR_ERR_CUDA;
}
int rnn_stream_max_clusters(gasr_ctx *ctx, int H, int *clusters, int *ctas_per_cluster) {
    (void)ctx;
    *ctas_per_cluster = H / RS_HC;
    if (H == 512) return rs_max_clusters<512>(clusters);
    if (H == 256) return rs_max_clusters<256>(clusters);
    if (H == 128) return rs_max_clusters<128>(clusters);
    *clusters = 0;
    return GASR_OK;
}

int rnn_stream_default_nsub(int N) { (void)N; return 0; }

int launch_rnn_stream(gasr_ctx *ctx, const RnnStreamParams &p, int H, cudaStream_t st) {
    GASR_CHECK(rnn_stream_supported(ctx, H, p.N, p.L), "rnn_stream: unsupported shape H=%d N=%d L=%d", H, p.N, p.L);
    GASR_CHECK(p.frames_per_block >= 1 && p.T >= 1, "rnn_stream: bad parameters");
    for (int l = 0; l < p.L; l++) {
        const long long rows = (long long)p.T * p.N;
        GASR_CHECK(rows * p.layer[l].ldxp < (1ll << 31) && rows * p.layer[l].ldo < (1ll << 31) && rows * p.layer[l].ldp < (1ll << 31),
                   "rnn_stream: sequence too large for 32-bit element offsets");
    }
    for (int l = 0; l < p.L; l++) {
        const RnnStreamLayer &y = p.layer[l];
        GASR_CHECK(y.xproj && y.w_hh && (y.out || y.out_hi), "rnn_stream: null layer operand");
        GASR_CHECK(!y.out_hi || (y.out_lo && y.ldp % 8 == 0 && (reinterpret_cast<uintptr_t>(y.out_hi) & 15) == 0 &&
                                 (reinterpret_cast<uintptr_t>(y.out_lo) & 15) == 0),
                   "rnn_stream: bf16 output planes must be 16-byte aligned");
    }
    GASR_CHECK(p.nsub == 0, "rnn_stream: sub-batched variant was removed");
    int rc;
    if (H == 512) rc = launch_rs<512>(ctx, p, st);
    else if (H == 256) rc = launch_rs<256>(ctx, p, st);
    else rc = launch_rs<128>(ctx, p, st);
    if (rc == GASR_OK) ctx->launches += 1;
    return rc;
}

}  // namespace gasr
