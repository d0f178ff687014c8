// umma_m64_layout.cu -- where does tcgen05.mma (cta_group::1, kind::f16, M = 64, N = 24) put accumulator element (i, j) in
// TMEM?  Groundwork for a tcgen05 recurrence kernel (64 hidden units per CTA = UMMA M 64).
//   A[i][k] = (k == 0) ? i + 1 : 0,  B[j][k] = (k == 0) ? j + 1 : 0   =>   D[i][j] = (i + 1) * (j + 1)   (exact)
// Every warp dumps its 32 TMEM lanes x 32 columns; the host prints which (lane, column) holds which (i, j).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I gpu-accelerated-speech-recognition_b200/csrc -I include \
//        tools/ubench/umma_m64_layout.cu -o tools/ubench/umma_m64_layout && tools/ubench/umma_m64_layout
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

using namespace gasr;

constexpr int M = 64, N = 24;

__device__ __forceinline__ uint32_t sw128_off(int r, int k) {      // bf16 element (row r, column k) of a [rows x 64] K-major tile
    return (uint32_t)(r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
}

__global__ void __launch_bounds__(128, 1) probe(float *out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    unsigned char *gen = smem_raw + (tiles - raw);
    unsigned char *a_tile = gen, *b_tile = gen + 8192;
    const uint32_t bar = tiles + 8192 + 4096;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + 8192 + 4096 + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (8192 + 4096) / 4; i += 128) reinterpret_cast<uint32_t *>(gen)[i] = 0u;
    __syncthreads();
    if (threadIdx.x < M) *reinterpret_cast<__nv_bfloat16 *>(a_tile + sw128_off(threadIdx.x, 0)) = __float2bfloat16_rn((float)(threadIdx.x + 1));
    if (threadIdx.x < N) *reinterpret_cast<__nv_bfloat16 *>(b_tile + sw128_off(threadIdx.x, 0)) = __float2bfloat16_rn((float)(threadIdx.x + 1));
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // poison the accumulator columns first so untouched lanes are recognisable
    {
        uint32_t z = 0x7fc00000u;                                        // NaN
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 32; c++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
        umma_bf16(tmem_base, umma_desc_sw128(tiles), umma_desc_sw128(tiles + 8192), idesc, 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 32; c++) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[(warp * 32 + lane) * 32 + c] = __uint_as_float(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
}

int main() {
    float *d = nullptr;
    cudaMalloc(&d, 128 * 32 * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    probe<<<1, 128, 16384>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static float h[128 * 32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int found = 0, as_expected = 0;
    for (int lane = 0; lane < 128; lane++) {
        int first = -1, cnt = 0;
        for (int c = 0; c < 32; c++) if (h[lane * 32 + c] == h[lane * 32 + c]) { cnt++; if (first < 0) first = c; }
        if (cnt == 0) continue;
        // D[i][j] = (i+1)(j+1): column 0 value = i + 1 if columns are j
        const float v0 = h[lane * 32 + first], v1 = first + 1 < 32 ? h[lane * 32 + first + 1] : 0.f;
        printf("lane %3d: %2d written columns from %2d, first values %g %g -> row i = %g (if column = j)\n", lane, cnt, first, v0, v1,
               v0 / (first + 1) - 1);
        found++;
        bool ok = cnt == N;
        for (int c = 0; c < N && ok; c++) ok = h[lane * 32 + c] == (float)((lane + 1) * (c + 1));
        as_expected += ok;
    }
    printf("lanes with data: %d; lanes l < 64 holding row l in columns 0..%d: %d\n", found, N - 1, as_expected);
    return 0;
}
