// zall_emu.cpp -- the per-thread body of the all-qubit <Z> pass (spinoza_b200/csrc/kernels_zall.cuh) executed on the CPU.
//
// Test infrastructure only (tests/test_reduce_cpu_emulation.py).  The kernel k_z_all is this body followed by the block
// reduction every other reduction of kernels_reduce.cu uses; what can go wrong in new code is the body's index arithmetic --
// which vector a thread reads, which bits of the index a vector's mass is credited to, the two-vectors-per-trip tail --
// and that is what runs here, for a grid of virtual threads whose partial sums are added up on the host.
#define SPZ_CPU_EMULATION 1
#include "cuda_cpu_shim.h"

#include "../../spinoza_b200/csrc/kernels_zall.cuh"

template <int NB>
static void run(const double *re, const double *im, long long nvec, int n, int grid, int threads, double *out) {
    for (int i = 0; i <= n; ++i) out[i] = 0.0;
    const long long stride = (long long)grid * threads;
    for (long long first = 0; first < stride; ++first) {
        double total = 0.0, s1[NB];
        for (int t = 0; t < NB; ++t) s1[t] = 0.0;
        spz::z_all_accumulate<NB>(re, im, nvec, n, first, stride, total, s1);
        out[0] += total;
        for (int t = 0; t < n; ++t) out[1 + t] += s1[t];
    }
}

// out: n + 1 doubles (total, then the mass at indices with bit t set); nb selects the instantiation as reduce_z_all does
extern "C" int emu_z_all(int n, const double *re, const double *im, int grid, int threads, double *out) {
    if (n < 2 || n > spz::kZMaxBits || threads != 256) return 1; // (the body relies on CTAs of 256 threads, see kZThreadBits)
    const long long nvec = (1ll << n) / 4;
    if (n <= 16) run<16>(re, im, nvec, n, grid, threads, out);
    else if (n <= 24) run<24>(re, im, nvec, n, grid, threads, out);
    else if (n <= 32) run<32>(re, im, nvec, n, grid, threads, out);
    else run<spz::kZMaxBits>(re, im, nvec, n, grid, threads, out);
    return 0;
}
