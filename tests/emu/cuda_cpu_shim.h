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

#include <atomic>
#include <sched.h>
#ifdef SPZ_EMU_TSAN
#include <sanitizer/tsan_interface.h>
#endif

namespace spz_emu {
extern unsigned char *dyn_smem;

// The barrier of one thread block.  Plain build: a pthread barrier.  ThreadSanitizer build (race check of the shared-memory
// protocol, i.e. "is a __syncthreads() missing?"): pthread_barrier_wait is useless for that -- TSan models it as
// release/acquire on ONE sync object, so a fast thread that has already arrived at the NEXT barrier publishes its later
// writes to slow threads still leaving the previous one, and the race disappears from the happens-before graph.  The TSan
// barrier synchronises through relaxed atomics (no happens-before edges of their own) and annotates each GENERATION with
// its own tag, which makes the graph exact.
struct CtaBarrier {
    unsigned n = 0;
    std::atomic<int> or_acc{0};
#ifdef SPZ_EMU_TSAN
    std::atomic<unsigned> count{0}, gen{0};
    char tags[4096];
    void init(unsigned threads) { n = threads; count.store(0); gen.store(0); }
    void destroy() {}
    void wait() {
        const unsigned g = gen.load(std::memory_order_relaxed);
        __tsan_release(&tags[g & 4095u]);
        if (count.fetch_add(1u, std::memory_order_relaxed) + 1u == n) {
            count.store(0u, std::memory_order_relaxed);
            gen.store(g + 1u, std::memory_order_relaxed);
        } else {
            while (gen.load(std::memory_order_relaxed) == g) sched_yield();
        }
        __tsan_acquire(&tags[g & 4095u]);
    }
#else
    pthread_barrier_t b;
    void init(unsigned threads) { n = threads; pthread_barrier_init(&b, nullptr, threads); }
    void destroy() { pthread_barrier_destroy(&b); }
    void wait() { pthread_barrier_wait(&b); }
#endif
    // __syncthreads_or: two barriers keep it simple (nobody can start the next round before everyone has read this one)
    int sync_or(int pred) {
        if (pred) or_acc.store(1, std::memory_order_relaxed);
        wait();
        const int r = or_acc.load(std::memory_order_relaxed);
        wait();
        or_acc.store(0, std::memory_order_relaxed);
        return r;
    }
};
extern CtaBarrier default_cta;                 // harnesses that run one block at a time
static thread_local CtaBarrier *cta = nullptr;  // harnesses with concurrent blocks point every thread at its block's barrier
inline void barrier() { (cta ? cta : &default_cta)->wait(); }
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
static inline int __syncthreads_or(int pred) { return (spz_emu::cta ? spz_emu::cta : &spz_emu::default_cta)->sync_or(pred); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
// round-to-nearest intrinsics: plain IEEE operations (the emulation is compiled with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
