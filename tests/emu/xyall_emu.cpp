// xyall_emu.cpp -- the tile body of the batched <X> / <Y> pass (spinoza_b200/csrc/kernels_xyall.cuh) on the CPU: one OS
// thread per CUDA thread, a pthread barrier for __syncthreads, a heap buffer for the staging arrays.
//
// Test infrastructure only (tests/test_reduce_cpu_emulation.py).  k_xy_all is this body followed by the block reduction
// every reduction of kernels_reduce.cu uses; here the per-thread sums are added up on the host instead.
#define SPZ_CPU_EMULATION 1
#include "cuda_cpu_shim.h"

#include <thread>
#include <vector>

#include "../../spinoza_b200/csrc/kernels_xyall.cuh"

namespace spz_emu {
unsigned char *dyn_smem = nullptr;
CtaBarrier default_cta;
} // namespace spz_emu

// out[b] = 2 * sum over the pairs of tile bit b, b < 12 (only the bits of tmask).  grid CTAs of `threads` threads.
extern "C" int emu_xy_tile(int n, const double *re, const double *im, int L, int H, const int *high, unsigned tmask, int obs, int grid,
                           int threads, double *out) {
    if (L + H > spz::kXYBits || L + H > n || H > 6) return 1;
    spz::XYArgs a{};
    a.re = re; a.im = im; a.L = L; a.H = H; a.tmask = tmask; a.obs = obs;
    for (int k = 0; k < H; ++k) a.high[k] = (unsigned char)high[k];
    a.n_tiles = (1ll << n) >> (L + H);
    const unsigned len = 1u << (L + H);
    std::vector<double> sre(len), sim(len);
    std::vector<double> sums((size_t)threads * spz::kXYBits, 0.0);
    static unsigned long long hoff[64];
    spz_emu::default_cta.init((unsigned)threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            threadIdx.x = (unsigned)t; blockDim.x = (unsigned)threads; gridDim.x = (unsigned)grid;
            for (int b = 0; b < grid; ++b) {
                blockIdx.x = (unsigned)b;
                double acc[spz::kXYBits] = {0.0};
                spz::xy_prepare(a, hoff);
                for (long long tile = b; tile < a.n_tiles; tile += grid) spz::xy_tile_accumulate(a, tile, hoff, sre.data(), sim.data(), acc);
                for (int k = 0; k < spz::kXYBits; ++k) sums[(size_t)t * spz::kXYBits + k] += acc[k];
                spz_emu::barrier();
            }
        });
    }
    for (auto &th : pool) th.join();
    spz_emu::default_cta.destroy();
    for (int k = 0; k < spz::kXYBits; ++k) {
        double s = 0.0;
        for (int t = 0; t < threads; ++t) s += sums[(size_t)t * spz::kXYBits + k];
        out[k] = 2.0 * s;
    }
    return 0;
}
