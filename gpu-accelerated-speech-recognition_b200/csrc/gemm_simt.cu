// gemm_simt.cu -- exact-fp32 dense helpers behind matrixMul / matrixMulTA / matrixMulTB / matrixAdd
// (reference cuMatrix.cpp:33-168, there cublasSgemm / cublasSgeam + a host sync per call).
// FFMA register-tiled GEMM (128x128x16 CTA tile, 8x8 per thread), any shape, optional transposes, optional
// fused row-broadcast bias.  This is the fp32-exact path; the batched input projection of the recurrent stack
// uses the tcgen05 kernel in xproj_gemm_tc.cu.
#include "common.cuh"

namespace gasr {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;

template <bool TX, bool TY>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ y,
                                                       int ldy, float *__restrict__ z, int ldz, int m, int k, int n,
                                                       const float *__restrict__ bias) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    const int ty = tid / 16, tx = tid % 16;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < k; k0 += BK) {
        // A tile: x[row0 + r][k0 + c]  (op(x) is m x k)
#pragma unroll
        for (int e = 0; e < BM * BK / 256; e++) {
            const int lin = tid + e * 256;
            int r, c;
            if (TX) { r = lin % BM; c = lin / BM; } else { c = lin % BK; r = lin / BK; }
            const int gr = row0 + r, gc = k0 + c;
            float v = 0.0f;
            if (gr < m && gc < k) v = TX ? x[(size_t)gc * ldx + gr] : x[(size_t)gr * ldx + gc];
            As[c][r] = v;
        }
        // B tile: y[k0 + r][col0 + c]  (op(y) is k x n)
#pragma unroll
        for (int e = 0; e < BN * BK / 256; e++) {
            const int lin = tid + e * 256;
            int r, c;
            if (TY) { r = lin % BK; c = lin / BK; } else { c = lin % BN; r = lin / BN; }
            const int gr = k0 + r, gc = col0 + c;
            float v = 0.0f;
            if (gr < k && gc < n) v = TY ? y[(size_t)gc * ldy + gr] : y[(size_t)gr * ldy + gc];
            Bs[r][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gr = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gr >= m) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int gc = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gc < n) z[(size_t)gr * ldz + gc] = acc[i][j] + (bias ? bias[gc] : 0.0f);
        }
    }
}

int launch_matmul(gasr_ctx *ctx, const float *x, int ldx, int tx, const float *y, int ldy, int ty, float *z, int ldz,
                  int m, int k, int n, const float *bias, cudaStream_t st) {
    GASR_CHECK(x && y && z, "matmul: null operand");
    GASR_CHECK(m >= 0 && k >= 0 && n >= 0, "matmul: negative dimension");
    GASR_CHECK(ldx >= (tx ? m : k) && ldy >= (ty ? k : n) && ldz >= n, "matmul: leading dimension smaller than the row");
    if (m == 0 || n == 0) return GASR_OK;
    dim3 grid(ceil_div(n, BN), ceil_div(m, BM));
    GASR_CHECK(grid.y <= 65535, "matmul: too many rows for one launch (%d)", m);
    if (!tx && !ty) gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(x, ldx, y, ldy, z, ldz, m, k, n, bias);
    else if (tx && !ty) gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(x, ldx, y, ldy, z, ldz, m, k, n, bias);
    else if (!tx && ty) gemm_f32_kernel<false, true><<<grid, 256, 0, st>>>(x, ldx, y, ldy, z, ldz, m, k, n, bias);
    else gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(x, ldx, y, ldy, z, ldz, m, k, n, bias);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

__global__ void matadd_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ y, int ldy,
                              float *__restrict__ z, int ldz, int rows, int cols, float lambda) {
    const size_t total = (size_t)rows * cols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        z[(size_t)r * ldz + c] = fmaf(lambda, y[(size_t)r * ldy + c], x[(size_t)r * ldx + c]);
    }
}

int launch_matadd(gasr_ctx *ctx, const float *x, int ldx, const float *y, int ldy, float *z, int ldz, int rows,
                  int cols, float lambda, cudaStream_t st) {
    GASR_CHECK(x && y && z, "matadd: null operand");
    GASR_CHECK(rows >= 0 && cols >= 0 && ldx >= cols && ldy >= cols && ldz >= cols, "matadd: bad shape");
    if (rows == 0 || cols == 0) return GASR_OK;
    const size_t total = (size_t)rows * cols;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    matadd_kernel<<<blocks, 256, 0, st>>>(x, ldx, y, ldy, z, ldz, rows, cols, lambda);
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return GASR_OK;
}

}  // namespace gasr
