// kernels_direct.cuh -- one-gate-per-pass kernels (the unfused path behind spz_apply / spz_c_apply /
// spz_mc_apply and the bandwidth sweep of BASELINE config 2).
//
// Replaces the per-gate CPU loops of gates.rs:322-1386.  One launch reads and writes each touched
// amplitude exactly once: algorithmic traffic 32 * 2^n bytes for a full gate (H/X/Y/RX/RY/RZ/U),
// 16 * 2^n for Z/P (s1 only), divided by 2^k under k controls.  HBM-bound; no tensor cores.
//
// Layout: re[] and im[] are separate f64 arrays (SoA, core.rs:20-24).  A thread owns U "vectors" of
// W consecutive amplitudes (W = 4 -> one 256-bit LDG/STG per array per pair side, W = 2 -> 128-bit),
// so every warp-level access is a run of 32*W*8 contiguous bytes on each of the (up to) four streams
// re[s0], re[s1], im[s0], im[s1].  All loads of the U vectors are issued before any arithmetic.
//
//   k_pair_vec : target bit >= log2(W).  vector index v -> amplitude index by inserting zero bits at the
//                target and control positions (controls then set to 1 -- or left 0 for a negative control, the
//                spz_mc_apply_signed extension); s1 = s0 | 2^t.
//   k_pair_low : target bit <  log2(W): the pair lives inside one vector, updated in registers.
//   k_pair_scalar: fully general scalar fallback (tiny states, n < log2(W)+1).
// Controls below log2(W) become a lane predicate (lane_cmask / lane_cval).
#pragma once

#include "gate_math.cuh"

namespace spz {

constexpr int kMaxIns = 40;

struct PairArgs {
    double *re;
    double *im;
    long long nvec;              // number of W-vectors to visit
    unsigned long long setmask;  // control bits >= log2(W), OR-ed into the index
    unsigned long long tbit;     // 1 << target (k_pair_vec), unused in k_pair_low
    int nins;                    // number of zero-bit insertions
    int lane_cmask;              // control bits < log2(W)
    int lane_cval;               // the value those bits must have (= lane_cmask unless some are negative controls)
    int tlow;                    // k_pair_low: the target bit (0 or 1)
    unsigned char pos[kMaxIns];  // ascending insertion positions (amplitude-index bit numbers)
    double s[7];
};

// ---- W-wide vector load/store with a cache policy --------------------------------------------------
// POL 0: default ld/st.  POL 1: streaming (.cs, evict-first).  POL 2: .cg (L2 only).
template <int W>
struct Vec {
    double v[W];
};

template <int W, int POL>
__device__ __forceinline__ Vec<W> ldv(const double *p) {
    Vec<W> r;
#ifdef SPZ_CPU_EMULATION // tests/emu/: the same kernel bodies compiled with g++; cache policies mean nothing there
    for (int l = 0; l < W; ++l) r.v[l] = p[l];
#else
    if constexpr (W == 4) {
        if constexpr (POL == 1)
            asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
        else if constexpr (POL == 2)
            asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
        else
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    } else if constexpr (W == 2) {
        double2 t;
        if constexpr (POL == 1) t = __ldcs(reinterpret_cast<const double2 *>(p));
        else if constexpr (POL == 2) t = __ldcg(reinterpret_cast<const double2 *>(p));
        else t = *reinterpret_cast<const double2 *>(p);
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
        r.v[0] = *p;
    }
#endif
    return r;
}

template <int W, int POL>
__device__ __forceinline__ void stv(double *p, const Vec<W> &r) {
#ifdef SPZ_CPU_EMULATION
    for (int l = 0; l < W; ++l) p[l] = r.v[l];
#else
    if constexpr (W == 4) {
        if constexpr (POL == 1)
            asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};"
                         :: "l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
        else if constexpr (POL == 2)
            asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};"
                         :: "l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
        else
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                         :: "l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
    } else if constexpr (W == 2) {
        double2 t = make_double2(r.v[0], r.v[1]);
        if constexpr (POL == 1) __stcs(reinterpret_cast<double2 *>(p), t);
        else if constexpr (POL == 2) __stcg(reinterpret_cast<double2 *>(p), t);
        else *reinterpret_cast<double2 *>(p) = t;
    } else {
        *p = r.v[0];
    }
#endif
}

template <int W> struct LogW;
template <> struct LogW<1> { static constexpr int v = 0; };
template <> struct LogW<2> { static constexpr int v = 1; };
template <> struct LogW<4> { static constexpr int v = 2; };

// NINS >= 0: compile-time number of insertions; NINS == -1: runtime a.nins.
template <int NINS>
__device__ __forceinline__ unsigned long long expand_index(unsigned long long x, const PairArgs &a) {
    if constexpr (NINS >= 0) {
#pragma unroll
        for (int k = 0; k < NINS; ++k) x = insert_zero(x, a.pos[k]);
    } else {
        for (int k = 0; k < a.nins; ++k) x = insert_zero(x, a.pos[k]);
    }
    return x | a.setmask;
}

// ---- target >= log2(W) -------------------------------------------------------------------------------
template <int KIND, int NINS, int W, int U, int THREADS, int POL>
__global__ void __launch_bounds__(THREADS) k_pair_vec(const PairArgs a) {
    constexpr bool S0 = GateTraits<KIND>::touches_s0;
    const long long v0 = (long long)blockIdx.x * (THREADS * U) + threadIdx.x;
    unsigned long long i0[U];
    Vec<W> r0[U], m0[U], r1[U], m1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long v = v0 + (long long)u * THREADS;
        if (v < a.nvec) {
            i0[u] = expand_index<NINS>((unsigned long long)v << LogW<W>::v, a);
            const unsigned long long i1 = i0[u] | a.tbit;
            if constexpr (S0) {
                r0[u] = ldv<W, POL>(a.re + i0[u]);
                m0[u] = ldv<W, POL>(a.im + i0[u]);
            }
            r1[u] = ldv<W, POL>(a.re + i1);
            m1[u] = ldv<W, POL>(a.im + i1);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long v = v0 + (long long)u * THREADS;
        if (v < a.nvec) {
#pragma unroll
            for (int l = 0; l < W; ++l) {
                if ((l & a.lane_cmask) == a.lane_cval) {
                    double x0 = 0.0, y0 = 0.0;
                    if constexpr (S0) { x0 = r0[u].v[l]; y0 = m0[u].v[l]; }
                    pair_update<KIND>(a.s, x0, y0, r1[u].v[l], m1[u].v[l]);
                    if constexpr (S0) { r0[u].v[l] = x0; m0[u].v[l] = y0; }
                }
            }
            const unsigned long long i1 = i0[u] | a.tbit;
            if constexpr (S0) {
                stv<W, POL>(a.re + i0[u], r0[u]);
                stv<W, POL>(a.im + i0[u], m0[u]);
            }
            stv<W, POL>(a.re + i1, r1[u]);
            stv<W, POL>(a.im + i1, m1[u]);
        }
    }
}

// ---- target < log2(W): pair inside the vector ---------------------------------------------------------
template <int KIND, int NINS, int W, int U, int THREADS, int POL>
__global__ void __launch_bounds__(THREADS) k_pair_low(const PairArgs a) {
    const long long v0 = (long long)blockIdx.x * (THREADS * U) + threadIdx.x;
    unsigned long long i0[U];
    Vec<W> r[U], m[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long v = v0 + (long long)u * THREADS;
        if (v < a.nvec) {
            i0[u] = expand_index<NINS>((unsigned long long)v << LogW<W>::v, a);
            r[u] = ldv<W, POL>(a.re + i0[u]);
            m[u] = ldv<W, POL>(a.im + i0[u]);
        }
    }
    const int tb = 1 << a.tlow;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long v = v0 + (long long)u * THREADS;
        if (v < a.nvec) {
#pragma unroll
            for (int l = 0; l < W; ++l) {
                // lanes with target bit clear drive the pair (l, l | tb)
                if (!(l & tb) && ((l & a.lane_cmask) == a.lane_cval)) {
                    // W is 2 or 4 and tb in {1,2}: resolve l|tb with compile-time-indexable selects
                    if (tb == 1) pair_update<KIND>(a.s, r[u].v[l], m[u].v[l], r[u].v[(l | 1) % W], m[u].v[(l | 1) % W]);
                    else         pair_update<KIND>(a.s, r[u].v[l], m[u].v[l], r[u].v[(l | 2) % W], m[u].v[(l | 2) % W]);
                }
            }
            stv<W, POL>(a.re + i0[u], r[u]);
            stv<W, POL>(a.im + i0[u], m[u]);
        }
    }
}

// ---- fully general scalar fallback (tiny registers) ---------------------------------------------------
struct ScalarArgs {
    double *re;
    double *im;
    long long npairs;
    unsigned long long setmask;
    unsigned long long tbit;
    int nins;
    int kind;
    unsigned char pos[kMaxIns];
    double s[7];
};

static __global__ void k_pair_scalar(const ScalarArgs a) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.npairs;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)p;
        for (int k = 0; k < a.nins; ++k) x = insert_zero(x, a.pos[k]);
        x |= a.setmask;
        const unsigned long long y = x | a.tbit;
        double ra = a.re[x], ia = a.im[x], rb = a.re[y], ib = a.im[y];
        pair_update_rt(a.kind, a.s, ra, ia, rb, ib);
        a.re[x] = ra; a.im[x] = ia; a.re[y] = rb; a.im[y] = ib;
    }
}

// ---- SWAP (swap_apply gates.rs:1376-1386): amp[lo=1,hi=0] <-> amp[lo=0,hi=1] ------------------------------
struct SwapArgs {
    double *re;
    double *im;
    long long nvec;
    int lo, hi; // lo < hi
};

template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) k_swap_vec(const SwapArgs a) {
    const long long v = (long long)blockIdx.x * THREADS + threadIdx.x;
    if (v >= a.nvec) return;
    unsigned long long x = (unsigned long long)v << LogW<W>::v;
    x = insert_zero(x, a.lo);
    x = insert_zero(x, a.hi);
    const unsigned long long ia = x | (1ull << a.lo), ib = x | (1ull << a.hi);
    Vec<W> ra = ldv<W, 0>(a.re + ia), ma = ldv<W, 0>(a.im + ia);
    Vec<W> rb = ldv<W, 0>(a.re + ib), mb = ldv<W, 0>(a.im + ib);
    stv<W, 0>(a.re + ia, rb); stv<W, 0>(a.im + ia, mb);
    stv<W, 0>(a.re + ib, ra); stv<W, 0>(a.im + ib, ma);
}

static __global__ void k_swap_scalar(const SwapArgs a) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.nvec;
         p += (long long)gridDim.x * blockDim.x) {
        unsigned long long x = insert_zero(insert_zero((unsigned long long)p, a.lo), a.hi);
        const unsigned long long ia = x | (1ull << a.lo), ib = x | (1ull << a.hi);
        double t = a.re[ia]; a.re[ia] = a.re[ib]; a.re[ib] = t;
        t = a.im[ia]; a.im[ia] = a.im[ib]; a.im[ib] = t;
    }
}

} // namespace spz
