// gemm_pair.cuh -- projection GEMM on CTA pairs (gemm_pair.cu): 256 x 256 tiles, tcgen05.mma.cta_group::2.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace gasr {

bool gemm_pair_supported(const gasr_ctx *ctx, int M, int H);
int gemm_pair_prepare(gasr_ctx *ctx);            // function attributes (call before concurrent kernels run)
// C[M, H] = A * W + bias over rows [row0, row0 + M) of the A planes; maps = {A hi, A lo, W^T hi, W^T lo}, every box 128 rows
int launch_gemm_pair(gasr_ctx *ctx, const CUtensorMap maps[4], int row0, int M, int K, int H, float *C, int ldc, const float *bias,
                     int precision, cudaStream_t st);

void gemm_pair_trace_dump();                     // instrumented build (make TRACE=1, GASR_GP_TRACE=1): where every unfinished CTA stands

}  // namespace gasr
