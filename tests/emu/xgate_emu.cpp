// xgate_emu.cpp -- the fused exchange + gate kernel (spinoza_b200/csrc/kernels_xgate.cuh) on the CPU: TWO ranks, each
// running a small persistent grid whose thread blocks are all alive at the same time, because the kernel's protocol is a
// conversation between block b of one rank and block b of the other.
//
// Test infrastructure only (tests/test_xgate_cpu_emulation.py).  One OS thread per CUDA thread; the system-scope
// release / acquire of the flags become C++ release / acquire atomics, so ThreadSanitizer checks exactly the claim the
// kernel's header makes: with those flags, no slot is overwritten while the partner still reads it.
#define SPZ_CPU_EMULATION 1
#include "cuda_cpu_shim.h"

#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

namespace spz {
static inline unsigned long long xg_timer_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline void xg_release(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long xg_acquire(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void xg_pause() { sched_yield(); }
} // namespace spz

#include "../../spinoza_b200/csrc/kernels_xgate.cuh"

namespace spz_emu {
unsigned char *dyn_smem = nullptr;
CtaBarrier default_cta;
} // namespace spz_emu

namespace spz {
// gate scalars as the product computes them
void set_error(const char *, ...) {}
#include "../../spinoza_b200/csrc/gate_resolve.inl"
} // namespace spz

namespace {

constexpr int W = 4, U = 4; // as dist_exchange_gate launches it

template <int KIND, int THREADS>
void run_pair(double *re0, double *im0, double *re1, double *im1, int n_local, int lq, const double *s7, int grid, unsigned long long flag_base,
              unsigned long long *flags0, unsigned long long *flags1, unsigned long long *err, int break_protocol) {
    spz::XGArgs a[2];
    double *re[2] = {re0, re1}, *im[2] = {im0, im1};
    unsigned long long *flags[2] = {flags0, flags1};
    for (int r = 0; r < 2; ++r) {
        a[r].mine_re = re[r]; a[r].mine_im = im[r]; a[r].peer_re = re[1 - r]; a[r].peer_im = im[1 - r];
        a[r].nvec = ((long long)1 << (n_local - 1)) / W;
        a[r].lq = lq; a[r].my_bit = r;
        a[r].peer_flag = flags[1 - r]; a[r].my_flag = flags[r];
        a[r].flag_base = flag_base;
        a[r].err = err + r;
        a[r].timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
        for (int k = 0; k < 7; ++k) a[r].s[k] = s7[k];
    }
    if (break_protocol) a[1].flag_base = flag_base - 1000; // rank 1 believes every step has been acknowledged already
    std::vector<spz_emu::CtaBarrier> bars(2 * grid);
    for (auto &b : bars) b.init(THREADS);
    std::vector<std::thread> pool;
    for (int r = 0; r < 2; ++r)
        for (int b = 0; b < grid; ++b)
            for (int t = 0; t < THREADS; ++t)
                pool.emplace_back([&, r, b, t]() {
                    threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)b;
                    blockDim.x = THREADS; gridDim.x = (unsigned)grid;
                    spz_emu::cta = &bars[r * grid + b];
                    spz::k_exchange_gate<KIND, W, U, THREADS>(a[r]);
                });
    for (auto &th : pool) th.join();
    for (auto &b : bars) b.destroy();
}

} // namespace

// Both shards in host memory; returns 0.  kind: spz_gate_kind of an uncontrolled non-diagonal gate.
extern "C" int emu_xgate(int n_local, int lq, int kind, const double *params, double *re0, double *im0, double *re1, double *im1, int grid,
                         int break_protocol) {
    spz::GateK g;
    if (int rc = spz::resolve_gate(kind, params, &g)) return rc;
    if (lq < 2 || lq >= n_local || grid < 1 || grid > spz::kMaxXgCtas) return -1;
    std::vector<unsigned long long> f0(spz::kMaxXgCtas, 0), f1(spz::kMaxXgCtas, 0);
    unsigned long long err[2] = {0, 0};
    constexpr int T = 32; // a small block keeps the OS thread count sane: 2 ranks x grid x 32
    const unsigned long long base = 4000; // as if earlier launches had used the flags
    switch (kind) {
    case SPZ_GATE_H: run_pair<SPZ_GATE_H, T>(re0, im0, re1, im1, n_local, lq, g.s, grid, base, f0.data(), f1.data(), err, break_protocol); break;
    case SPZ_GATE_X: run_pair<SPZ_GATE_X, T>(re0, im0, re1, im1, n_local, lq, g.s, grid, base, f0.data(), f1.data(), err, break_protocol); break;
    case SPZ_GATE_Y: run_pair<SPZ_GATE_Y, T>(re0, im0, re1, im1, n_local, lq, g.s, grid, base, f0.data(), f1.data(), err, break_protocol); break;
    case SPZ_GATE_RX: run_pair<SPZ_GATE_RX, T>(re0, im0, re1, im1, n_local, lq, g.s, grid, base, f0.data(), f1.data(), err, break_protocol); break;
    case SPZ_GATE_RY: run_pair<SPZ_GATE_RY, T>(re0, im0, re1, im1, n_local, lq, g.s, grid, base, f0.data(), f1.data(), err, break_protocol); break;
    case SPZ_GATE_U: run_pair<SPZ_GATE_U, T>(re0, im0, re1, im1, n_local, lq, g.s, grid, base, f0.data(), f1.data(), err, break_protocol); break;
    default: return -2;
    }
    return (err[0] || err[1]) ? 3 : 0;
}

#ifdef SPZ_EMU_MAIN
// tile-emu style stand-alone driver for the ThreadSanitizer run:
//   xgate_emu <n_local> <lq> <kind> <grid> <break 0|1> <state.bin: re0 im0 re1 im1, each 2^n_local f64>   (rewritten in place)
#include <cstdio>
#include <cstdlib>
int main(int argc, char **argv) {
    if (argc != 7) return 64;
    const int n_local = std::atoi(argv[1]), lq = std::atoi(argv[2]), kind = std::atoi(argv[3]), grid = std::atoi(argv[4]), brk = std::atoi(argv[5]);
    const size_t len = (size_t)1 << n_local;
    std::vector<double> st(4 * len);
    FILE *f = std::fopen(argv[6], "rb");
    if (!f || std::fread(st.data(), sizeof(double), 4 * len, f) != 4 * len) return 65;
    std::fclose(f);
    const double params[3] = {0.3, 0.5, 0.7};
    const int rc = emu_xgate(n_local, lq, kind, params, st.data(), st.data() + len, st.data() + 2 * len, st.data() + 3 * len, grid, brk);
    if (rc != 0) return 70 + rc;
    f = std::fopen(argv[6], "wb");
    if (!f || std::fwrite(st.data(), sizeof(double), 4 * len, f) != 4 * len) return 65;
    std::fclose(f);
    return 0;
}
#endif
