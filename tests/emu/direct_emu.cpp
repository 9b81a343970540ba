// direct_emu.cpp -- the one-gate-per-pass path (spinoza_b200/csrc/kernels_direct.cu: launch_gate / launch_swap and the
// k_pair_* / k_swap_* kernels they dispatch to) executed on the CPU.
//
// Test infrastructure only: built by tests/test_direct_cpu_emulation.py with g++ and tests/emu/cuda_cpu_shim.h force-included.
// The dispatch code and the kernel bodies are #included unchanged (SPZ_LAUNCH runs the grid as two nested loops; the
// 256-bit ld/st helpers fall back to plain loads), so argument construction, zero-bit insertion, lane masks for low
// controls, the low-target in-register path and the scalar fallback are all the product's own code.
#define SPZ_CPU_EMULATION 1
#include "cuda_cpu_shim.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "../../spinoza_b200/csrc/kernels_direct.cu"

namespace spz {
static char g_err[512];
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t, const char *, const char *, int) { return SPZ_ERR_CUDA; }
void count_launch(int) {}
int dist_join(spz_state *) { return SPZ_OK; }
int arrival_join(spz_state *) { return SPZ_OK; }
#include "../../spinoza_b200/csrc/gate_resolve.inl"
} // namespace spz

static spz_state make_state(int n, double *re, double *im) {
    spz_state st;
    st.n = n;
    st.len = (int64_t)1 << n;
    st.re = re;
    st.im = im;
    return st;
}

extern "C" int emu_apply(int n, double *re, double *im, int kind, const double *params, unsigned long long ctrl_mask, int target) {
    spz_state st = make_state(n, re, im);
    spz::GateK g;
    if (int rc = spz::resolve_gate(kind, params, &g)) return rc;
    return spz::launch_gate(&st, g, ctrl_mask, target);
}

extern "C" int emu_apply_signed(int n, double *re, double *im, int kind, const double *params, unsigned long long ctrl_mask,
                                unsigned long long neg_mask, int target) {
    spz_state st = make_state(n, re, im);
    spz::GateK g;
    if (int rc = spz::resolve_gate(kind, params, &g)) return rc;
    return spz::launch_gate_signed(&st, g, ctrl_mask, neg_mask, target);
}

extern "C" int emu_swap(int n, double *re, double *im, int t0, int t1) {
    spz_state st = make_state(n, re, im);
    return spz::launch_swap(&st, t0, t1);
}

extern "C" const char *emu_last_error() { return spz::g_err; }
