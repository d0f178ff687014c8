// rec_tcgen05.cu -- feasibility probe for the recurrence on tcgen05.
//
// One cluster of 8 CTAs runs  h_t = tanh(h_{t-1} W_hh + xp_t)  for NU utterances and H = 512 hidden units, 64 units per CTA:
//   * the CTA's W_hh^T slice [64 units x 512] sits in shared memory as bf16 hi / lo planes, K-major with the 128-byte
//     swizzle, and is the A operand of tcgen05.mma (cta_group::1, kind::f16, M = 64);
//   * h_{t-1} of all NU utterances [NU x 512] (bf16 hi / lo planes, same layout) is the B operand (N = NU) -- it is exactly
//     the buffer the other CTAs fill through DSMEM (st.async + mbarrier complete_tx), double-buffered by step parity;
//   * one thread issues the 96 MMAs of a step (8 K-blocks x 4 x {hi*hi, hi*lo, lo*hi}), the accumulator [64 x NU] fp32 lives
//     in TMEM (row i -> lane 32 (i / 16) + i % 16, see umma_m64_layout.cu);
//   * four epilogue warps (16 active lanes each = 16 hidden units, NU values per thread) read it back, add xp, apply tanh,
//     split to bf16 hi / lo, transpose through a small staging tile and all-gather 16-byte chunks into every CTA's B buffer.
// xp is a cheap integer hash (no memory traffic): the probe times the on-chip loop only.  The result after T_CHECK steps is
// compared with a double-precision host recurrence; then T_TIME steps are timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I gpu-accelerated-speech-recognition_b200/csrc -I include \
//        tools/ubench/rec_tcgen05.cu -o tools/ubench/rec_tcgen05 && tools/ubench/rec_tcgen05
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace cg = cooperative_groups;
using namespace gasr;

constexpr int H = 512, CS = 8, HC = 64, KB = H / 64;
constexpr int THREADS = 160;                         // warps 0-3: epilogue (TMEM lane quadrant = warp), warp 4: MMA issuer

__host__ __device__ inline float w_val(int k, int u) {      // W_hh[k][u], U(-1/sqrt(H), 1/sqrt(H))-like
    const unsigned x = (unsigned)(k * 7919 + u * 104729 + 12345) * 2654435761u;
    return ((float)(x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.0883883f;
}
__host__ __device__ inline float xp_val(int s, int j, int u) {
    const int v = (s * 131 + j * 31 + u * 17) % 97;
    return ((float)v * (1.0f / 97.0f) - 0.5f) * 0.5f;
}

__device__ __forceinline__ uint32_t sw128_off(int r, int k) {      // bf16 element (row r, column k) of a [rows x 64] K-major tile
    return (uint32_t)(r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, const uint4 &v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity, volatile int *giveup) {
    const long long t0 = clock64();
    int spins = 0;
    while (!mbar_try(bar, parity)) {
        if (*giveup) return false;
        if ((++spins & 255) == 0 && clock64() - t0 > 2000000000ll) { *giveup = 1; return false; }
    }
    return true;
}
__device__ __forceinline__ float fast_tanh(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

template <int NU>
struct Layout {
    static constexpr int A_PLANE = KB * 8192;                         // [64 rows x 128 B] per K-block
    static constexpr int B_KB = NU * 128;                             // [NU rows x 128 B] per K-block (NU % 8 == 0)
    static constexpr int B_PLANE = KB * B_KB;
    static constexpr int B_BUF = 2 * B_PLANE;                         // hi + lo
    static constexpr int OFF_A = 0, OFF_B = 2 * A_PLANE, OFF_STG = OFF_B + 2 * B_BUF;
    static constexpr int STG_WARP = 2 * NU * 16 * 2;                  // [plane][utt][16 units] bf16
    static constexpr int OFF_BAR = OFF_STG + 4 * STG_WARP;            // full[2], tfull, tmem slot, give-up flag
    static constexpr int BYTES = OFF_BAR + 64 + 1024;
    static constexpr uint32_t TX = (uint32_t)CS * NU * HC * 2 * 2;    // bytes one step delivers into one B buffer
};

template <int NU>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(THREADS, 1) rec_probe(int T, float *out, int *status) {
    using LT = Layout<NU>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *gen = smem_raw + (base - raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full0 = base + LT::OFF_BAR, tfull = full0 + 16;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + LT::OFF_BAR + 32);
    volatile int *giveup = reinterpret_cast<volatile int *>(gen + LT::OFF_BAR + 40);

    // ---- init: W slice as swizzled bf16 hi / lo planes, h_0 = 0 --------------------------------------------------------
    for (int i = tid; i < (2 * LT::B_BUF) / 16; i += THREADS) reinterpret_cast<uint4 *>(gen + LT::OFF_B)[i] = make_uint4(0, 0, 0, 0);
    for (int e = tid; e < HC * H; e += THREADS) {
        const int u = e / H, k = e % H;
        const float w = w_val(k, rank * HC + u);
        const __nv_bfloat16 hi = __float2bfloat16_rn(w), lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const uint32_t off = (uint32_t)(k >> 6) * 8192u + sw128_off(u, k & 63);
        *reinterpret_cast<__nv_bfloat16 *>(gen + LT::OFF_A + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(gen + LT::OFF_A + LT::A_PLANE + off) = lo;
    }
    if (tid == 0) {
        mbar_init(full0, 1); mbar_init(full0 + 8, 1); mbar_init(tfull, 1);
        *giveup = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) { mbar_expect_tx(full0, LT::TX); mbar_expect_tx(full0 + 8, LT::TX); }
    cluster.sync();

    if (warp == 4) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NU >> 3) << 17) | ((uint32_t)(HC >> 4) << 24);
            uint32_t phase_bits = 0;
            for (int s = 1; s < T; s++) {                                // step 0: h_0 = 0, nothing to multiply
                const int b = (s - 1) & 1;                               // buffer that receives h_{s-1}
                if (!wait_bounded(full0 + 8 * b, (phase_bits >> b) & 1u, giveup)) break;
                phase_bits ^= 1u << b;
                mbar_expect_tx(full0 + 8 * b, LT::TX);                   // re-arm for h_{s+1}
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = base + LT::OFF_A, a_lo = a_hi + LT::A_PLANE;
                const uint32_t b_hi = base + LT::OFF_B + b * LT::B_BUF, b_lo = b_hi + LT::B_PLANE;
#pragma unroll 1
                for (int kb = 0; kb < KB; kb++) {
                    const uint64_t dah = umma_desc_sw128(a_hi + kb * 8192), dal = umma_desc_sw128(a_lo + kb * 8192);
                    const uint64_t dbh = umma_desc_sw128(b_hi + kb * LT::B_KB), dbl = umma_desc_sw128(b_lo + kb * LT::B_KB);
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) {
                        const uint64_t adv = (uint64_t)(k4 * 32 >> 4);
                        umma_bf16(tmem_base, dah + adv, dbh + adv, idesc, (kb | k4) != 0);
                        umma_bf16(tmem_base, dah + adv, dbl + adv, idesc, 1);
                        umma_bf16(tmem_base, dal + adv, dbh + adv, idesc, 1);
                    }
                }
                umma_commit(tfull);
            }
        }
    } else {
        // ===== epilogue warps: lane l < 16 owns hidden unit 16 warp + l, all NU utterances =====
        const int u_loc = 16 * warp + (lane & 15), u_glob = rank * HC + u_loc;
        unsigned char *stg = gen + LT::OFF_STG + warp * LT::STG_WARP;
        float hfin[NU];
        for (int s = 0; s < T; s++) {
            uint32_t v[NU];
            if (s > 0) {
                if (!wait_bounded(tfull, (uint32_t)((s - 1) & 1), giveup)) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
                for (int c = 0; c < NU / 8; c++)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(v[8 * c]), "=r"(v[8 * c + 1]), "=r"(v[8 * c + 2]), "=r"(v[8 * c + 3]), "=r"(v[8 * c + 4]),
                                   "=r"(v[8 * c + 5]), "=r"(v[8 * c + 6]), "=r"(v[8 * c + 7])
                                 : "r"(taddr + 8 * c));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < NU; j++) v[j] = 0u;
            }
            __syncwarp();                                               // the previous step's chunk reads are done
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < NU; j++) {
                    const float h = fast_tanh(__uint_as_float(v[j]) + xp_val(s, j, u_glob));
                    hfin[j] = h;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(h), lo = __float2bfloat16_rn(h - __bfloat162float(hi));
                    reinterpret_cast<__nv_bfloat16 *>(stg)[(0 * NU + j) * 16 + lane] = hi;
                    reinterpret_cast<__nv_bfloat16 *>(stg)[(1 * NU + j) * 16 + lane] = lo;
                }
            }
            __syncwarp();
            if (s + 1 < T) {
                // all-gather: chunk (plane, utterance, 8-unit group) -> 16 bytes into every CTA's buffer for h_s
                const int nb = s & 1;
                const uint32_t bar_local = full0 + 8 * nb;
#pragma unroll
                for (int c = lane; c < 4 * NU; c += 32) {
                    const int pl = c / (2 * NU), j = (c >> 1) % NU, g2 = c & 1;
                    const uint4 chunk = *reinterpret_cast<const uint4 *>(stg + ((pl * NU + j) * 16 + g2 * 8) * 2);
                    const uint32_t dst_local = base + LT::OFF_B + nb * LT::B_BUF + pl * LT::B_PLANE + rank * LT::B_KB +
                                               (uint32_t)(j * 128 + (((2 * warp + g2) ^ (j & 7)) << 4));
#pragma unroll
                    for (int r = 0; r < CS; r++) st_async_v4(mapa(dst_local, r), chunk, mapa(bar_local, r));
                }
            }
        }
        if (lane < 16)
            for (int j = 0; j < NU; j++) out[(size_t)j * H + u_glob] = hfin[j];
    }
    if (*giveup && tid == 0) atomicExch(status, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster.sync();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
    }
}

template <int NU>
static int run(int t_check, int t_time) {
    using LT = Layout<NU>;
    float *d_out; int *d_status;
    cudaMalloc(&d_out, sizeof(float) * NU * H);
    cudaMalloc(&d_status, sizeof(int));
    cudaMemset(d_status, 0, sizeof(int));
    cudaFuncSetAttribute(rec_probe<NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT::BYTES);
    rec_probe<NU><<<CS, THREADS, LT::BYTES>>>(t_check, d_out, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    int status = 0;
    cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || status) { printf("NU=%d: failed (%s, watchdog %d)\n", NU, cudaGetErrorString(e), status); return 1; }
    std::vector<float> got(NU * H);
    cudaMemcpy(got.data(), d_out, sizeof(float) * NU * H, cudaMemcpyDeviceToHost);
    // host reference, double accumulation
    std::vector<double> h(NU * H, 0.0), hn(NU * H);
    std::vector<float> W((size_t)H * H);
    for (int k = 0; k < H; k++) for (int u = 0; u < H; u++) W[(size_t)k * H + u] = w_val(k, u);
    for (int s = 0; s < t_check; s++) {
        for (int j = 0; j < NU; j++)
            for (int u = 0; u < H; u++) {
                double acc = xp_val(s, j, u);
                if (s > 0) for (int k = 0; k < H; k++) acc += h[(size_t)j * H + k] * (double)W[(size_t)k * H + u];
                hn[(size_t)j * H + u] = std::tanh(acc);
            }
        h.swap(hn);
    }
    double err = 0.0;
    for (int i = 0; i < NU * H; i++) err = std::fmax(err, std::fabs((double)got[i] - h[i]));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 3; it++) {
        cudaEventRecord(e0);
        rec_probe<NU><<<CS, THREADS, LT::BYTES>>>(t_time, d_out, d_status);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    printf("NU=%2d utterances per cluster: max |h - host| after %d steps = %.2e; %d steps in %.3f ms = %.3f us/step  (smem %d B)\n",
           NU, t_check, err, t_time, best, 1e3 * best / t_time, LT::BYTES);
    return err < 1e-4 ? 0 : 2;
}

int main() {
    int rc = 0;
    rc |= run<8>(20, 2000);
    rc |= run<16>(20, 2000);      // (24 double-buffered utterances + staging exceed 227 KB: needs the single-buffer handshake)
    return rc;
}
