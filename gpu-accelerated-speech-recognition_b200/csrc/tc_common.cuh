// tc_common.cuh -- tcgen05 / TMEM / TMA / mbarrier helpers shared by the projection GEMM kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gasr {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3, TC_THREADS = 192;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;              // 16 KB: one [128 x 64] bf16 tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;             // A_hi, A_lo, B_hi, B_lo
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor of a [rows x 64 bf16] tile (1024-byte aligned):
// start address >> 4 | LBO = 1 (16 B, unused for swizzled K-major) | SBO = 1024 B between 8-row groups |
// version 1 (Blackwell) | layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}


// 2-D bf16 tensor [rows, Kp] row-major, box [box_rows x 64 cols], 128-byte swizzle (xproj_gemm_tc.cu)
int tc_make_map(CUtensorMap *map, const void *base, int rows, int Kp, int box_rows);

}  // namespace gasr

// Instrumented build only (make TRACE=1): per-SM ring of the last eight TMEM allocator events in mapped host memory
// (kernel id << 28 | event << 24 | CTA): events 1 = before alloc, 2 = after alloc, 3 = before dealloc, 4 = after dealloc.
#ifdef GASR_RW_TRACE
namespace gasr { unsigned *trace_tmem_log(); void trace_tmem_log_dump(); }
#define GASR_TLOG(ptr, kid, ev) do { if ((ptr) && (threadIdx.x & 31) == 0) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); \
    volatile unsigned *l_ = (ptr) + sm_ * 16; const unsigned i_ = l_[8]; l_[i_ & 7] = ((unsigned)(kid) << 28) | ((unsigned)(ev) << 24) | (blockIdx.x & 0xffffffu); l_[8] = i_ + 1; } } while (0)
#else
#define GASR_TLOG(ptr, kid, ev) do { } while (0)
#endif

