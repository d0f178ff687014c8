/* gasr_cxx.h -- shared plumbing of the C++ module mirror (cuMatrix.h, Linear.h, RNN_Cell.h, RNN.h, CTCBeamSearch.h):
 * a per-thread default gasr_ctx and the reference's error convention (print + exit) on top of the C ABI's codes. */
#ifndef GASR_CXX_H
#define GASR_CXX_H
#include <stdio.h>
#include <stdlib.h>

#include "gasr.h"

namespace gasr_cxx {

/* One context per host thread (the reference has process-global singletons, cuMatrix.cpp:18-30; a multi-GPU driver
 * calls use_device(d) in each worker thread before constructing modules). */
inline gasr_ctx *&ctx_slot() {
    static thread_local gasr_ctx *ctx = nullptr;
    return ctx;
}
inline void use_device(int device) {
    if (ctx_slot()) gasr_ctx_destroy(ctx_slot());
    ctx_slot() = nullptr;
    if (gasr_ctx_create(device, &ctx_slot()) != GASR_OK) {
        printf("gasr: %s\n", gasr_last_error());
        exit(1);
    }
}
inline gasr_ctx *ctx() {
    if (!ctx_slot()) use_device(0);
    return ctx_slot();
}
/* The reference prints and exit(0)s on any error (cuMatrix.cpp:37-42); the wrappers print and exit(1). */
inline void check(int status, const char *what) {
    if (status != GASR_OK) {
        printf("%s: %s\n", what, gasr_last_error());
        exit(1);
    }
}

}  // namespace gasr_cxx
#endif
