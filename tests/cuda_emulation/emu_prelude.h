// Host emulation of the small CUDA subset the region-geometry kernels use, so their SOURCE TEXT (the
// anonymous namespace of csrc/yam_regiongeom.cu, pasted below this prelude by the test) runs on the CPU:
// one OS thread per CUDA thread of a block, pthread barrier = __syncthreads, blocks one after another
// (so function-local `static` storage stands in for __shared__).  Test infrastructure only -- it checks
// kernel logic (indexing, arithmetic, barrier placement), not launch configuration or performance.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <thread>
#include <vector>

#define YAM_PROPS_STRIDE 8
#define __global__
#define __device__
#define __forceinline__ inline
#define __constant__ static const
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static pthread_barrier_t emu_barrier;

static inline void __syncthreads() { pthread_barrier_wait(&emu_barrier); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}

template <typename F>
static void emu_launch(unsigned grid, unsigned block, F&& kernel) {
    gridDim = dim3(grid);
    blockDim = dim3(block);
    for (unsigned b = 0; b < grid; b++) {
        pthread_barrier_init(&emu_barrier, nullptr, block);
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; t++)
            threads.emplace_back([&, t, b]() {
                threadIdx = dim3(t);
                blockIdx = dim3(b);
                kernel();
            });
        for (auto& th : threads) th.join();
        pthread_barrier_destroy(&emu_barrier);
    }
}
