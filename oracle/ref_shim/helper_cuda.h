/* Stand-in for the CUDA-samples header the reference's Makefile points at
 * (-I/usr/local/cuda/samples/common/inc, reference Makefile:2); that directory does not exist in
 * this image.  Only the two macros the reference uses are provided. */
#ifndef GASR_REF_SHIM_HELPER_CUDA_H
#define GASR_REF_SHIM_HELPER_CUDA_H
#include <cuda_runtime.h>
#include <stdio.h>
#define checkCudaErrors(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define getLastCudaError(msg) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
    fprintf(stderr, "%s: CUDA error %s at %s:%d\n", msg, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#endif
