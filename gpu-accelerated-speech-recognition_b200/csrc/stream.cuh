// stream.cuh -- parameter blocks of the streaming pipeline's persistent kernels (xproj_stream.cu, stream_pipeline.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace gasr {

constexpr int XS_MAX_TARGETS = 5;
constexpr int XS_KIND_XPROJ = 0, XS_KIND_LOGSOFTMAX = 1;
constexpr int XS_EPI_WARPS = 4;          // every epilogue warp counts a finished tile once

struct XsTarget {
    int kind;                 // XS_KIND_*
    int cta0, nctas;          // the CTAs that serve this target
    int n_tiles, bn;          // column tiles per block, tile width (128, or 32 for the output layer)
    int kblocks, terms;       // K / 64, 3 (fp32-grade split) or 1 (bf16)
    int V;                    // output layer: valid columns
    float *C; int ldc;
    const float *bias;
    const unsigned *src_done; int src_need;   // block b of the A operand is complete when src_done[b] >= src_need
    unsigned *dst_ready;                      // += XS_EPI_WARPS per finished tile of block b
};

struct XsParams {
    int M;                    // rows (T * N)
    int row0;                 // first row of the A operand in its TMA descriptor (the wave engine runs one time chunk per launch)
    int n_blocks;             // ceil(M / 128)
    int n_targets;
    int stages;               // shared-memory ring depth (0 = default 3)
    int wide;                 // 1: targets may use 256-column tiles (96 KB stages, two of them; 2 x 256 TMEM columns)
    int *error;               // mapped host word: watchdog code
    volatile unsigned *abort; // device word: any persistent kernel of the pipeline gave up -> everybody stops waiting
    volatile unsigned *tlog;  // instrumented build only (make TRACE=1)
    XsTarget target[XS_MAX_TARGETS];
};

struct XsMaps { CUtensorMap m[4 * XS_MAX_TARGETS]; };   // per target: A hi, A lo, W^T hi, W^T lo

int launch_xproj_stream(gasr_ctx *ctx, const XsMaps &maps, const XsParams &p, int ctas, cudaStream_t st);

}  // namespace gasr
