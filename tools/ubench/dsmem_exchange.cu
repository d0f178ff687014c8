// Microbenchmark: per-step latency of an 8-CTA all-gather of 4 KB slices through distributed shared memory.
//   A: st.shared::cluster.v4 + barrier.cluster (release/acquire)
//   B: st.async.v4 + mbarrier complete_tx (no cluster barrier)
//   C: local staging + cp.async.bulk smem->remote smem + mbarrier complete_tx
// Also: mma.sync.m16n8k16 bf16 issue rate.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t tx) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void bulk_s2s(uint32_t rdst, uint32_t src, uint32_t bytes, uint32_t rbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}

template <int MODE, int CS, int BYTES_PER_THREAD>
__global__ void __launch_bounds__(256, 1) exchange_kernel(int steps, unsigned long long *cycles, uint32_t *sink, unsigned char *gscratch) {
    extern __shared__ __align__(128) unsigned char sm[];
    // layout: buf[2][CS][256 * BPT] | stage[256 * BPT] | mbar[2]
    constexpr int SLICE = 256 * BYTES_PER_THREAD;
    unsigned char *buf = sm;
    unsigned char *stage = sm + 2 * CS * SLICE;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stage + SLICE);
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, rank = cluster.block_rank();
    if (tid == 0) { mbar_init(smem_u32(&mbar[0]), 1); mbar_init(smem_u32(&mbar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (MODE != 0 && tid == 0) { mbar_expect_tx(smem_u32(&mbar[0]), CS * SLICE); mbar_expect_tx(smem_u32(&mbar[1]), CS * SLICE); }
    cluster.sync();
    uint32_t acc = tid;
    long long t0 = clock64();
    for (int s = 0; s < steps; s++) {
        const int b = s & 1;
        unsigned char *dstbuf = buf + (size_t)b * CS * SLICE + (size_t)rank * SLICE + tid * BYTES_PER_THREAD;
        uint4 v = make_uint4(acc, acc + 1, acc + 2, acc + 3);
        if (MODE == 0) {
            for (int r = 0; r < CS; r++)
                for (int q = 0; q < BYTES_PER_THREAD / 16; q++) st_cluster_v4(mapa(smem_u32(dstbuf + q * 16), r), v);
            asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        } else if (MODE == 1) {
            for (int r = 0; r < CS; r++)
                for (int q = 0; q < BYTES_PER_THREAD / 16; q++)
                    st_async_v4(mapa(smem_u32(dstbuf + q * 16), r), v, mapa(smem_u32(&mbar[b]), r));
            mbar_wait(smem_u32(&mbar[b]), (s >> 1) & 1);
            if (tid == 0) mbar_expect_tx(smem_u32(&mbar[b]), CS * SLICE);   // re-arm for the use two steps later
        } else if (MODE == 3) {
            // slice -> global scratch (double buffered by step parity), then ONE multicast bulk load delivers it to all CTAs
            unsigned char *g = gscratch + ((size_t)blockIdx.x * 2 + b) * SLICE;
            for (int q = 0; q < BYTES_PER_THREAD / 16; q++) *reinterpret_cast<uint4 *>(g + tid * BYTES_PER_THREAD + q * 16) = v;
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (tid == 0)
                bulk_g2s_mc(smem_u32(buf + (size_t)b * CS * SLICE + (size_t)rank * SLICE), g, SLICE, smem_u32(&mbar[b]), (uint16_t)((1u << CS) - 1));
            mbar_wait(smem_u32(&mbar[b]), (s >> 1) & 1);
            if (tid == 0) mbar_expect_tx(smem_u32(&mbar[b]), CS * SLICE);
        } else {
            for (int q = 0; q < BYTES_PER_THREAD / 16; q++) *reinterpret_cast<uint4 *>(stage + tid * BYTES_PER_THREAD + q * 16) = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < CS)
                bulk_s2s(mapa(smem_u32(buf + (size_t)b * CS * SLICE + (size_t)rank * SLICE), tid), smem_u32(stage), SLICE,
                         mapa(smem_u32(&mbar[b]), tid));
            mbar_wait(smem_u32(&mbar[b]), (s >> 1) & 1);
            if (tid == 0) mbar_expect_tx(smem_u32(&mbar[b]), CS * SLICE);
        }
        // consume: read one word of another CTA's slice (data dependency into the next step)
        acc += *reinterpret_cast<const uint32_t *>(buf + (size_t)b * CS * SLICE + (size_t)((rank + 1) % CS) * SLICE + tid * BYTES_PER_THREAD);
        if (MODE == 2) __syncthreads();   // stage reuse
    }
    long long t1 = clock64();
    cluster.sync();
    if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * 256 + tid] = acc;
}

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// CHAINS independent accumulators per warp, `iters` rounds
template <int CHAINS>
__global__ void mma_rate_kernel(int iters, unsigned long long *cycles, float *sink) {
    float d[CHAINS][4];
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) d[c][i] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
    uint32_t b0 = threadIdx.x * 3, b1 = threadIdx.x * 5;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) mma16816(d[c], a, b0 + c, b1);
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) s += d[c][i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE, int CS, int BPT>
void run_exchange(const char *name, int clusters) {
    constexpr int SLICE = 256 * BPT;
    size_t smem = 2 * CS * SLICE + SLICE + 64;
    auto kern = exchange_kernel<MODE, CS, BPT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long *cyc; uint32_t *sink; unsigned char *gs;
    CK(cudaMalloc(&cyc, 8 * 1024)); CK(cudaMalloc(&sink, 4 * 256 * 1024)); CK(cudaMalloc(&gs, 2 * SLICE * 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CS); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int steps = 2000;
    for (int rep = 0; rep < 2; rep++) CK(cudaLaunchKernelEx(&cfg, kern, steps, cyc, sink, gs));
    CK(cudaDeviceSynchronize());
    unsigned long long h[1024];
    CK(cudaMemcpy(h, cyc, 8 * clusters * CS, cudaMemcpyDeviceToHost));
    double mx = 0, mn = 1e30;
    for (int i = 0; i < clusters * CS; i++) { double c = (double)h[i] / steps; if (c > mx) mx = c; if (c < mn) mn = c; }
    printf("%-44s CS=%d slice=%5d B (per CTA in: %6d B) clusters=%2d: %7.1f .. %7.1f cycles/step\n", name, CS, SLICE, CS * SLICE, clusters, mn, mx);
    cudaFree(cyc); cudaFree(sink); cudaFree(gs);
}

template <int CHAINS>
void run_mma(int warps) {
    unsigned long long *cyc; float *sink;
    CK(cudaMalloc(&cyc, 8 * 1024)); CK(cudaMalloc(&sink, 4 * 1024 * 1024));
    const int iters = 4096;
    mma_rate_kernel<CHAINS><<<148, warps * 32>>>(iters, cyc, sink);
    mma_rate_kernel<CHAINS><<<148, warps * 32>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    unsigned long long h[148];
    CK(cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost));
    double c = (double)h[0] / iters;
    printf("mma.sync m16n8k16 bf16: warps/SM=%2d chains=%2d: %6.2f cycles per round (%5.2f cyc/MMA/warp, %6.1f FMA/cyc/SM)\n", warps, CHAINS, c,
           c / CHAINS, (double)warps * CHAINS * 2048.0 / c);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    run_exchange<0, 8, 16>("A st.shared::cluster + barrier.cluster", 1);
    run_exchange<0, 8, 16>("A st.shared::cluster + barrier.cluster", 12);
    run_exchange<1, 8, 16>("B st.async + mbarrier", 1);
    run_exchange<1, 8, 16>("B st.async + mbarrier", 12);
    run_exchange<2, 8, 16>("C staging + cp.async.bulk + mbarrier", 1);
    run_exchange<2, 8, 16>("C staging + cp.async.bulk + mbarrier", 12);
    run_exchange<0, 8, 32>("A 2x bytes", 12);
    run_exchange<1, 8, 32>("B 2x bytes", 12);
    run_exchange<2, 8, 32>("C 2x bytes", 12);
    run_exchange<0, 4, 16>("A CS=4", 12);
    run_exchange<1, 4, 16>("B CS=4", 12);
    run_exchange<2, 4, 16>("C CS=4", 12);
    run_exchange<1, 4, 32>("B CS=4 2x bytes", 12);
    run_exchange<2, 4, 32>("C CS=4 2x bytes", 12);
    run_exchange<1, 2, 32>("B CS=2 2x bytes", 12);
    run_exchange<2, 2, 32>("C CS=2 2x bytes", 12);
    run_exchange<3, 8, 16>("D global + multicast bulk load", 1);
    run_exchange<3, 8, 16>("D global + multicast bulk load", 12);
    run_exchange<3, 8, 32>("D 2x bytes", 12);
    run_exchange<3, 4, 16>("D CS=4", 12);
    run_exchange<3, 4, 32>("D CS=4 2x bytes", 12);
    run_mma<1>(1); run_mma<1>(4); run_mma<2>(4); run_mma<4>(4); run_mma<8>(4); run_mma<12>(4); run_mma<24>(4);
    run_mma<4>(8); run_mma<8>(8); run_mma<12>(8); run_mma<24>(8); run_mma<8>(16);
    return 0;
}
