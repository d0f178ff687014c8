/* cuMatrix.h -- mirror of the reference's tensor class (cuMatrix.h:13-229) and dense helpers (cuMatrix.h:231-240,
 * cuMatrix.cpp:33-168) as inline wrappers over the C ABI (gasr.h).  Same constructors, methods and public fields:
 * row-major [rows, cols, channels], lazily allocated pinned host + device copies (zero-filled), blocking toGpu/toCpu,
 * shallow offset views that never free and refuse toGpu/toCpu.  No CUDA headers are needed to compile a caller. */
#ifndef _CU_MATRIX_H_
#define _CU_MATRIX_H_
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>

#include "MemoryMonitor.h"
#include "gasr_cxx.h"

/* the CUDA runtime's opaque stream handle, declared here so that callers need no CUDA headers (an identical typedef in
 * <driver_types.h> is a legal redeclaration) */
struct CUstream_st;
typedef struct CUstream_st *cudaStream_t;

template <class T>
class cuMatrix {
public:
    /* deep copy of host data (cuMatrix.h:18-23) */
    cuMatrix(T *_data, int _n, int _m, int _c)
        : cols(_m), rows(_n), channels(_c), hostData(NULL), devData(NULL), isShallow(false) {
        mallocHost();
        memcpy(hostData, _data, sizeof(*hostData) * (size_t)cols * rows * channels);
    }
    cuMatrix(int _n, int _m, int _c) : cols(_m), rows(_n), channels(_c), hostData(NULL), devData(NULL), isShallow(false) {}
    /* "slicing": a view of `other` at element offset (cuMatrix.h:33-46); whole rows keep 16-byte alignment */
    cuMatrix(cuMatrix<T> *other, int offset, int _n, int _m, int _c) : cols(_m), rows(_n), channels(_c), isShallow(true) {
        if (other->hostData == NULL || other->devData == NULL) {
            printf("Error: offset constructor from uninitialized matrix");
            hostData = NULL; devData = NULL;
        } else if (offset + _n * _m * _c > other->getLen()) {
            printf("Error: offset constructor out of bound");
            hostData = NULL; devData = NULL;
        } else {
            hostData = other->hostData + offset;
            devData = other->devData + offset;
        }
    }
    cuMatrix(int _n, int _m, int _c, T *hostPtr, T *devPtr)
        : cols(_m), rows(_n), channels(_c), hostData(hostPtr), devData(devPtr), isShallow(true) {}

    void freeCudaMem() {
        if (isShallow) return;
        if (devData) { MemoryMonitor::instance()->freeGpuMemory(devData); devData = NULL; }
    }
    ~cuMatrix() {
        if (!isShallow) {
            if (hostData) MemoryMonitor::instance()->freeCpuMemory(hostData);
            if (devData) MemoryMonitor::instance()->freeGpuMemory(devData);
        }
    }
    void toCpu() {
        if (isShallow) { printf("Error: attempting to manipulate memory of a shallow copy."); return; }
        mallocDev(); mallocHost();
        gasr_cxx::check(gasr_memcpy_d2h(gasr_cxx::ctx(), hostData, devData, bytes()), "cuMatrix::toCPU data download failed");
    }
    void toGpu() {
        if (isShallow) { printf("Error: attempting to manipulate memory of a shallow copy."); return; }
        mallocDev(); mallocHost();
        gasr_cxx::check(gasr_memcpy_h2d(gasr_cxx::ctx(), devData, hostData, bytes()), "cuMatrix::toGPU data upload failed");
    }
    /* asynchronous upload on a caller's stream (reference cuMatrix.h:108-115 takes a cudaStream_t); the handle is passed to
     * the library as an opaque pointer, so a caller needs no CUDA headers unless it creates streams itself */
    void toGpu(cudaStream_t stream) {
        if (isShallow) { printf("Error: attempting to manipulate memory of a shallow copy."); return; }
        mallocDev(); mallocHost();
        gasr_cxx::check(gasr_memcpy_h2d_on_stream(gasr_cxx::ctx(), devData, hostData, bytes(), (void *)stream), "cuMatrix::toGpu(stream)");
    }
    /* the same on the context's own stream */
    void toGpuAsync() { toGpu((cudaStream_t)0); }
    void gpuClear() {
        if (isShallow) { printf("Error: attempting to manipulate memory of a shallow copy."); return; }
        mallocDev();
        gasr_cxx::check(gasr_memset_device(gasr_cxx::ctx(), devData, 0, bytes()), "device memory cudaMemset failed");
    }
    void cpuClear() {
        if (isShallow) { printf("Error: attempting to manipulate memory of a shallow copy."); return; }
        mallocHost();
        memset(hostData, 0, bytes());
    }
    void set(int i, int j, int k, T v) { mallocHost(); hostData[(i * cols + j) + cols * rows * k] = v; }
    T get(int i, int j, int k) { mallocHost(); return hostData[(i * cols + j) + cols * rows * k]; }
    int getLen() { return rows * cols * channels; }
    int getArea() { return rows * cols; }
    int getRows() { return rows; }
    int getCols() { return cols; }
    T *&getHost() { mallocHost(); return hostData; }
    T *&getDev() { mallocDev(); return devData; }

    int cols;
    int rows;
    int channels;

private:
    T *hostData;
    T *devData;
    bool isShallow;
    size_t bytes() const { return sizeof(T) * (size_t)cols * rows * channels; }
    void mallocHost() {
        if (NULL == hostData) hostData = (T *)MemoryMonitor::instance()->cpuMalloc(bytes());   /* pinned, zero-filled */
    }
    void mallocDev() {
        if (NULL == devData) {
            void *p = NULL;
            gasr_cxx::check(MemoryMonitor::instance()->gpuMalloc(&p, bytes()), "cuMatrix::cuMatrix device memory allocation failed");
            devData = (T *)p;                                                                   /* zero-filled */
        }
    }
};

inline void printMatrixInfo(cuMatrix<float> *mat) {
    std::cout << "shape: (" << mat->rows << ", " << mat->cols << ")" << std::endl;
    for (int i = 0; i < mat->rows; i++) {
        for (int j = 0; j < mat->cols; j++) std::cout << mat->getHost()[i * mat->cols + j] << "\t";
        std::cout << std::endl;
    }
}
/* z = x * y (cuMatrix.cpp:33-70) */
inline void matrixMul(cuMatrix<float> *x, cuMatrix<float> *y, cuMatrix<float> *z) {
    if (x->channels != 1 || y->channels != 1 || z->channels != 1) { printf("matrix mul channels != 1\n"); exit(1); }
    if (x->cols != y->rows || z->rows != x->rows || z->cols != y->cols) { printf("matrix mul dimension mismatch\n"); exit(1); }
    gasr_cxx::check(gasr_matmul(gasr_cxx::ctx(), x->getDev(), x->cols, 0, y->getDev(), y->cols, 0, z->getDev(), z->cols,
                                x->rows, x->cols, y->cols), "matrixMul");
}
/* z = T(x) * y (cuMatrix.cpp:73-108) */
inline void matrixMulTA(cuMatrix<float> *x, cuMatrix<float> *y, cuMatrix<float> *z) {
    if (x->rows != y->rows || z->rows != x->cols || z->cols != y->cols) { printf("matrix mul dimension mismatch\n"); exit(1); }
    gasr_cxx::check(gasr_matmul(gasr_cxx::ctx(), x->getDev(), x->cols, 1, y->getDev(), y->cols, 0, z->getDev(), z->cols,
                                x->cols, x->rows, y->cols), "matrixMulTA");
}
/* z = x * T(y) (cuMatrix.cpp:111-145) */
inline void matrixMulTB(cuMatrix<float> *x, cuMatrix<float> *y, cuMatrix<float> *z) {
    if (x->cols != y->cols || z->rows != x->rows || z->cols != y->rows) { printf("matrix mul dimension mismatch\n"); exit(1); }
    gasr_cxx::check(gasr_matmul(gasr_cxx::ctx(), x->getDev(), x->cols, 0, y->getDev(), y->cols, 1, z->getDev(), z->cols,
                                x->rows, x->cols, y->rows), "matrixMulTB");
}
/* z = x + lambda * y (cuMatrix.cpp:147-168) */
inline void matrixAdd(cuMatrix<float> *x, cuMatrix<float> *y, cuMatrix<float> *z, float lambda) {
    gasr_cxx::check(gasr_matadd(gasr_cxx::ctx(), x->getDev(), x->cols, y->getDev(), y->cols, z->getDev(), z->cols, x->rows,
                                x->cols, lambda), "matrixAdd");
}
#endif
