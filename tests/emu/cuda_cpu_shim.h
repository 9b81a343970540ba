// cuda_cpu_shim.h -- just enough of the CUDA execution model to run one thread block of a kernel on the CPU.
//
// Test infrastructure only (tests/test_tile_cpu_emulation.py builds tests/emu/tile_emu.cpp with g++ and this header
// force-included).  A block is executed by blockDim.x OS threads; __syncthreads() is a pthread barrier; `__shared__`
// variables become function-local statics (blocks run one after another, so one copy is enough); the dynamic shared-memory
// window is a heap buffer.  Nothing here models warps, memory ordering or timing: the emulation checks indexing, control
// flow and arithmetic of the kernel body, not its performance or its behaviour under races.
#pragma once

#include <cuda_runtime.h> // vector types (double2, uint4), make_double2; host-side declarations only under g++

#include <cmath>
#include <cstdint>
#include <pthread.h>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)

namespace spz_emu {
extern unsigned char *dyn_smem;
extern unsigned block_threads;
#ifdef SPZ_EMU_TSAN
// ThreadSanitizer build (race check of the shared-memory protocol, i.e. "is a __syncthreads() missing?").
// pthread_barrier_wait is useless for that: TSan models it as release/acquire on ONE sync object, so a fast thread that has
// already arrived at the NEXT barrier publishes its later writes to slow threads still leaving the previous one, and the
// race disappears from the happens-before graph.  This barrier synchronises through relaxed atomics (no happens-before
// edges of their own) and annotates each GENERATION with its own tag, which makes the graph exact.
} // namespace spz_emu
#include <atomic>
#include <sanitizer/tsan_interface.h>
#include <sched.h>
namespace spz_emu {
extern std::atomic<unsigned> bar_count, bar_gen;
extern char bar_tags[4096];
inline void barrier() {
    const unsigned g = bar_gen.load(std::memory_order_relaxed);
    __tsan_release(&bar_tags[g & 4095u]);
    if (bar_count.fetch_add(1u, std::memory_order_relaxed) + 1u == block_threads) {
        bar_count.store(0u, std::memory_order_relaxed);
        bar_gen.store(g + 1u, std::memory_order_relaxed);
    } else {
        while (bar_gen.load(std::memory_order_relaxed) == g) sched_yield();
    }
    __tsan_acquire(&bar_tags[g & 4095u]);
}
#else
extern pthread_barrier_t block_barrier;
inline void barrier() { pthread_barrier_wait(&block_barrier); }
#endif
} // namespace spz_emu

static thread_local uint3 threadIdx;
static thread_local uint3 blockIdx;
static thread_local uint3 blockDim;
static thread_local uint3 gridDim;

namespace spz_emu {
// Kernels without barriers: no OS threads needed, just the two loops of the grid.
template <class F>
inline void run_flat(unsigned grid, unsigned threads, F f) {
    gridDim.x = grid; gridDim.y = 1; gridDim.z = 1;
    blockDim.x = threads; blockDim.y = 1; blockDim.z = 1;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < threads; ++t) {
            blockIdx.x = b; blockIdx.y = 0; blockIdx.z = 0;
            threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
            f();
        }
}
} // namespace spz_emu

// cache-policy loads/stores: plain accesses
static inline double2 __ldcs(const double2 *p) { return *p; }
static inline double2 __ldcg(const double2 *p) { return *p; }
static inline void __stcs(double2 *p, double2 v) { *p = v; }
static inline void __stcg(double2 *p, double2 v) { *p = v; }

// the dispatch code checks for launch errors; there is nothing to fail here
#define cudaGetLastError() cudaSuccess

static inline void __syncthreads() { spz_emu::barrier(); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
// round-to-nearest intrinsics: plain IEEE operations (the emulation is compiled with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
