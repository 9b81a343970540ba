// bw_probe.cu -- tuning probe for the one-gate-per-pass kernels (kernels_direct.cuh) on a real B200.
// Sweeps vector width W, unroll U, CTA size and cache policy for RX at several targets of an n-qubit state and
// prints achieved GB/s (32 * 2^n bytes per pass) next to two ceilings measured the same way: cudaMemcpy D2D and an
// in-place scale kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spinoza_b200/csrc
//        tools/bw_probe.cu -o tools/bw_probe      Run: tools/bw_probe [n_qubits]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels_direct.cuh"

namespace spz {
void set_error(const char *, ...) {}
int cuda_fail(cudaError_t e, const char *what, const char *, int) { fprintf(stderr, "CUDA %s: %s\n", what, cudaGetErrorString(e)); exit(1); }
void count_launch(int) {}
}
using namespace spz;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int W, int POL>
__global__ void __launch_bounds__(256) k_scale(double *re, double *im, long long nvec) {
    const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    if (v >= nvec) return;
    Vec<W> r = ldv<W, POL>(re + v * W), m = ldv<W, POL>(im + v * W);
#pragma unroll
    for (int l = 0; l < W; ++l) { r.v[l] *= 0.999; m.v[l] *= 1.001; }
    stv<W, POL>(re + v * W, r); stv<W, POL>(im + v * W, m);
}

static cudaEvent_t e0, e1;
template <typename F>
static double time_ms(F f, int reps = 5) {
    f();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

static PairArgs make_args(double *re, double *im, int n, int t, int logw) {
    PairArgs a{};
    a.re = re; a.im = im;
    const bool low = t < logw;
    a.nins = low ? 0 : 1;
    a.nvec = 1ll << (n - logw - a.nins);
    a.tbit = low ? 0 : (1ull << t);
    a.tlow = low ? t : 0;
    a.pos[0] = (unsigned char)t;
    a.s[0] = 0.8775825618903728; a.s[1] = -0.479425538604203;
    return a;
}

template <int KIND, int W, int U, int THREADS, int POL>
static double run_cfg(double *re, double *im, int n, int t) {
    PairArgs a = make_args(re, im, n, t, LogW<W>::v);
    const long long per = (long long)THREADS * U;
    const unsigned grid = (unsigned)((a.nvec + per - 1) / per);
    if (t < LogW<W>::v) return time_ms([&] { k_pair_low<KIND, 0, W, U, THREADS, POL><<<grid, THREADS>>>(a); });
    return time_ms([&] { k_pair_vec<KIND, 1, W, U, THREADS, POL><<<grid, THREADS>>>(a); });
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 30;
    const long long len = 1ll << n;
    double *re, *im, *tmp;
    CK(cudaMalloc(&re, 8 * len)); CK(cudaMalloc(&im, 8 * len));
    CK(cudaMemset(re, 0, 8 * len)); CK(cudaMemset(im, 0, 8 * len));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double gb = 32.0 * len / 1e9;
    printf("# n=%d  bytes per pass = %.2f GB\n", n, gb);
    if (cudaMalloc(&tmp, 8 * len) == cudaSuccess) {
        double ms = time_ms([&] { cudaMemcpyAsync(tmp, re, 8 * len, cudaMemcpyDeviceToDevice); });
        printf("ceiling cudaMemcpy D2D (r+w bytes)      : %8.1f GB/s\n", 16.0 * len / 1e9 / (ms * 1e-3));
        cudaFree(tmp);
    }
    { double ms = time_ms([&] { k_scale<2, 0><<<(unsigned)(len / 2 / 256), 256>>>(re, im, len / 2); });
      printf("ceiling in-place scale W=2 pol0         : %8.1f GB/s\n", gb / (ms * 1e-3)); }
    { double ms = time_ms([&] { k_scale<4, 0><<<(unsigned)(len / 4 / 256), 256>>>(re, im, len / 4); });
      printf("ceiling in-place scale W=4 pol0         : %8.1f GB/s\n", gb / (ms * 1e-3)); }
    { double ms = time_ms([&] { k_scale<4, 1><<<(unsigned)(len / 4 / 256), 256>>>(re, im, len / 4); });
      printf("ceiling in-place scale W=4 pol1(.cs)    : %8.1f GB/s\n", gb / (ms * 1e-3)); }
    { double ms = time_ms([&] { k_scale<4, 2><<<(unsigned)(len / 4 / 256), 256>>>(re, im, len / 4); });
      printf("ceiling in-place scale W=4 pol2(.cg)    : %8.1f GB/s\n", gb / (ms * 1e-3)); }

    std::vector<int> ts = {0, 1, 2, 3, 5, 8, 12, 16, 20, 24, n - 3, n - 2, n - 1};
    printf("%-28s", "config \\ target");
    for (int t : ts) printf("%8d", t);
    printf("\n");
#define ROW(KIND, W, U, TH, POL)                                                        \
    do {                                                                                \
        printf("%-4s W=%d U=%d T=%-3d pol=%d      ", #KIND + 9, W, U, TH, POL);         \
        for (int t : ts) printf("%8.0f", gb / (run_cfg<KIND, W, U, TH, POL>(re, im, n, t) * 1e-3)); \
        printf("\n"); fflush(stdout);                                                   \
    } while (0)
    ROW(SPZ_GATE_RX, 4, 1, 256, 0);
    ROW(SPZ_GATE_RX, 4, 2, 256, 0);
    ROW(SPZ_GATE_RX, 4, 4, 256, 0);
    ROW(SPZ_GATE_RX, 4, 1, 512, 0);
    ROW(SPZ_GATE_RX, 4, 2, 512, 0);
    ROW(SPZ_GATE_RX, 4, 2, 128, 0);
    ROW(SPZ_GATE_RX, 2, 1, 256, 0);
    ROW(SPZ_GATE_RX, 2, 2, 256, 0);
    ROW(SPZ_GATE_RX, 2, 4, 256, 0);
    ROW(SPZ_GATE_RX, 2, 4, 512, 0);
    ROW(SPZ_GATE_RX, 4, 2, 256, 1);
    ROW(SPZ_GATE_RX, 4, 2, 256, 2);
    ROW(SPZ_GATE_RX, 4, 1, 256, 1);
    ROW(SPZ_GATE_RX, 2, 2, 256, 1);
    ROW(SPZ_GATE_H, 4, 2, 256, 0);
    ROW(SPZ_GATE_RZ, 4, 2, 256, 0);
    ROW(SPZ_GATE_P, 4, 2, 256, 0);
    ROW(SPZ_GATE_X, 4, 2, 256, 0);
    return 0;
}
