/* MemoryMonitor.h -- mirror of the reference's allocation-accounting singleton (MemoryMonitor.h:9-28,
 * MemoryMonitor.cpp:9-51) over the C ABI.  Sizes are size_t (the reference's `int size` caps a buffer at 2 GiB). */
#ifndef __MEMORY_MONITOR_H__
#define __MEMORY_MONITOR_H__
#include <stddef.h>
#include <stdio.h>

#include "gasr_cxx.h"

class MemoryMonitor {
public:
    static MemoryMonitor *instance() {
        static MemoryMonitor *monitor = new MemoryMonitor();
        return monitor;
    }
    void *cpuMalloc(size_t size) {
        void *p = NULL;
        gasr_cxx::check(gasr_malloc_host(gasr_cxx::ctx(), size, &p), "MemoryMonitor::cpuMalloc");
        return p;
    }
    int gpuMalloc(void **devPtr, size_t size) { return gasr_malloc_device(gasr_cxx::ctx(), size, devPtr); }
    void freeGpuMemory(void *ptr) { gasr_free_device(gasr_cxx::ctx(), ptr); }
    void freeCpuMemory(void *ptr) { gasr_free_host(gasr_cxx::ctx(), ptr); }
    void printCpuMemory() {
        size_t d = 0, h = 0;
        gasr_memory_stats(gasr_cxx::ctx(), &d, &h);
        printf("total malloc cpu memory %fMb\n", h / 1024.0 / 1024.0);
    }
    void printGpuMemory() {
        size_t d = 0, h = 0;
        gasr_memory_stats(gasr_cxx::ctx(), &d, &h);
        printf("total malloc gpu memory %fMb\n", d / 1024.0 / 1024.0);
    }
};
#endif
