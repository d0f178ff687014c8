// rnn_wide_dev.cuh -- device helpers shared by the wide tcgen05 recurrence kernels (rnn_wide.cu: one CTA per 64 columns;
// rnn_wide2.cu: CTA pairs, tcgen05.mma.cta_group::2).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace gasr {

constexpr int RW_COLS = 64;                       // W_hh columns per CTA (UMMA N)
constexpr int RW_U = 128;                         // utterances per group (UMMA M)
constexpr int RW_STAGES = 2;                      // 2 x 32 KB in flight covers the TMA latency at the ~40 B/clk an SM ingests
constexpr int RW_EPI_WARPS = 8;
constexpr int RW_THREADS = 64 + 32 * RW_EPI_WARPS;   // TMA warp, MMA warp, epilogue warps
constexpr int RW_A_TILE = RW_U * TC_BK * 2;       // 16 KB: [128 x 64] bf16
constexpr int RW_STAGE_BYTES = 2 * RW_A_TILE;     // hi + lo
constexpr int RW_W_TILE = RW_COLS * TC_BK * 2;    // 8 KB: [64 x 64] bf16
constexpr int RW_EPI_STAGE = 4096;                // per epilogue warp: [32 rows x 64 B] hi + lo, to store whole 64-byte row segments
constexpr unsigned long long RW_TIMEOUT_NS = 4000000000ull;

__device__ __forceinline__ uint32_t rw_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void rw_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned long long rw_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded waits: a protocol bug must end as a launch failure, never as a hung GPU.
__device__ __forceinline__ void rw_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned long long t0 = 0;
    for (int spins = 0; ; spins++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if ((spins & 4095) == 4095) {
            const unsigned long long now = rw_now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > RW_TIMEOUT_NS) __trap();
        }
    }
}
__device__ __forceinline__ void rw_wait_cluster(uint32_t bar, uint32_t parity) {     // acquires the cluster's released stores
    uint32_t ok = 0;
    unsigned long long t0 = 0;
    for (int spins = 0; ; spins++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if ((spins & 4095) == 4095) {
            const unsigned long long now = rw_now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > RW_TIMEOUT_NS) __trap();
        }
    }
}
__device__ __forceinline__ void rw_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void rw_tma_load_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void rw_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ float rw_tanh(float x) {          // SFU: ex2.approx + rcp.approx, |error| ~ 1e-7 (parity bar 1e-4)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ void rw_tmem_ld32(uint32_t (&v)[32], uint32_t taddr) {
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

// ---- CTA pairs (tcgen05 cta_group::2) ----------------------------------------------------------------------------------
__device__ __forceinline__ void rw2_tma_load_to_leader(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    // the data lands in THIS CTA's shared memory, the transaction bytes complete on the pair leader's mbarrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void rw2_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void rw2_commit_pair(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

}  // namespace gasr
