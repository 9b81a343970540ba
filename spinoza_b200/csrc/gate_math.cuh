// gate_math.cuh -- per-pair amplitude updates, one per reference gate.
//
// Notation follows SURVEY.md 8(a): s0 = (a, b) = (re, im) at target bit 0, s1 = (c, d) at target bit 1.
// Each update performs the SAME IEEE-754 double operations in the SAME order as the reference's
// serial body (file:line cited per gate): __dmul_rn/__dadd_rn/__dsub_rn where the reference writes
// `*`,`+`,`-` (these intrinsics are never contracted into FMAs by nvcc) and __fma_rn where it calls
// `mul_add`.  With host-computed scalars this makes every unfused gate bit-identical to the CPU
// oracle, which is what tests/test_gpu_parity.py asserts.
#pragma once

#include "engine.h"

namespace spz {

#define SPZ_SQRT_ONE_HALF 0.70710678118654752440 /* math.rs:5 */

template <int KIND>
struct GateTraits {
    // Z and P only touch s1 (gates.rs:1048-1057, 829-840): s0 is neither loaded nor stored.
    static constexpr bool touches_s0 = !(KIND == SPZ_GATE_Z || KIND == SPZ_GATE_P);
};

template <int KIND>
__device__ __forceinline__ void pair_update(const double *__restrict__ s, double &a, double &b, double &c,
                                            double &d) {
    if constexpr (KIND == SPZ_GATE_H) {
        // h_apply_strat2 gates.rs:551-560
        const double a1 = __dmul_rn(SPZ_SQRT_ONE_HALF, a), b1 = __dmul_rn(SPZ_SQRT_ONE_HALF, b);
        const double c1 = __dmul_rn(SPZ_SQRT_ONE_HALF, c), d1 = __dmul_rn(SPZ_SQRT_ONE_HALF, d);
        a = __dadd_rn(a1, c1);
        b = __dadd_rn(b1, d1);
        c = __dsub_rn(a1, c1);
        d = __dsub_rn(b1, d1);
    } else if constexpr (KIND == SPZ_GATE_X) {
        // x_apply_target gates.rs:329-334
        double t = a; a = c; c = t;
        t = b; b = d; d = t;
    } else if constexpr (KIND == SPZ_GATE_Y) {
        // y_proc_chunk gates.rs:455-464: s0 <- (d, -c), s1 <- (-b, a)
        const double na = d, nb = -c, nc = -b, nd = a;
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (KIND == SPZ_GATE_Z) {
        // z_proc_chunk gates.rs:1051-1055
        c = -c;
        d = -d;
    } else if constexpr (KIND == SPZ_GATE_P) {
        // p_proc_chunk gates.rs:835-838: re' = z_re.mul_add(cos, -z_im * sin); im' = z_im.mul_add(cos, z_re * sin)
        const double cs = s[0], sn = s[1];
        const double zr = c, zi = d;
        c = __fma_rn(zr, cs, -__dmul_rn(zi, sn));
        d = __fma_rn(zi, cs, __dmul_rn(zr, sn));
    } else if constexpr (KIND == SPZ_GATE_RX) {
        // rx_apply_target gates.rs:720-724
        const double cs = s[0], ns = s[1];
        const double na = __dsub_rn(__dmul_rn(a, cs), __dmul_rn(d, ns));
        const double nb = __dadd_rn(__dmul_rn(b, cs), __dmul_rn(c, ns));
        const double nc = __dadd_rn(__dmul_rn(b, -ns), __dmul_rn(c, cs));
        const double nd = __dadd_rn(__dmul_rn(d, cs), __dmul_rn(a, ns));
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (KIND == SPZ_GATE_RY) {
        // ry_apply_strategy2 gates.rs:1116-1119
        const double sn = s[0], cs = s[1];
        const double na = __dsub_rn(__dmul_rn(a, cs), __dmul_rn(c, sn));
        const double nb = __dsub_rn(__dmul_rn(b, cs), __dmul_rn(d, sn));
        const double nc = __dadd_rn(__dmul_rn(a, sn), __dmul_rn(c, cs));
        const double nd = __dadd_rn(__dmul_rn(b, sn), __dmul_rn(d, cs));
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (KIND == SPZ_GATE_RZ) {
        // rz_apply_strategy1 gates.rs:941-945 with m = d0 = (cos, -sin) on s0, d1 = (cos, sin) on s1
        const double cs = s[0], sn = s[1];
        const double na = __fma_rn(a, cs, -__dmul_rn(b, -sn));
        const double nb = __fma_rn(a, -sn, __dmul_rn(b, cs));
        const double nc = __fma_rn(c, cs, -__dmul_rn(d, sn));
        const double nd = __fma_rn(c, sn, __dmul_rn(d, cs));
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (KIND == SPZ_GATE_U) {
        // u_apply_target gates.rs:1246-1259 (c,d,m,n) = (a,b,c,d) here
        const double ga = s[0], k = s[1], l = s[2], q = s[3], r = s[4], ss = s[5], t = s[6];
        const double t0 = __fma_rn(ga, a, __fma_rn(k, c, -__dmul_rn(l, d)));
        const double t1 = __fma_rn(ga, b, __fma_rn(k, d, __dmul_rn(l, c)));
        const double t2 = __fma_rn(q, a, __fma_rn(-r, b, __fma_rn(ss, c, -__dmul_rn(t, d))));
        const double t3 = __fma_rn(q, b, __fma_rn(r, a, __fma_rn(ss, d, __dmul_rn(t, c))));
        a = t0; b = t1; c = t2; d = t3;
    }
}

// Runtime-dispatched variant for the fused tile kernel (kind is uniform across the CTA).
__device__ __forceinline__ void pair_update_rt(int kind, const double *__restrict__ s, double &a, double &b,
                                               double &c, double &d) {
    switch (kind) {
    case SPZ_GATE_H: pair_update<SPZ_GATE_H>(s, a, b, c, d); break;
    case SPZ_GATE_X: pair_update<SPZ_GATE_X>(s, a, b, c, d); break;
    case SPZ_GATE_Y: pair_update<SPZ_GATE_Y>(s, a, b, c, d); break;
    case SPZ_GATE_Z: pair_update<SPZ_GATE_Z>(s, a, b, c, d); break;
    case SPZ_GATE_P: pair_update<SPZ_GATE_P>(s, a, b, c, d); break;
    case SPZ_GATE_RX: pair_update<SPZ_GATE_RX>(s, a, b, c, d); break;
    case SPZ_GATE_RY: pair_update<SPZ_GATE_RY>(s, a, b, c, d); break;
    case SPZ_GATE_RZ: pair_update<SPZ_GATE_RZ>(s, a, b, c, d); break;
    case SPZ_GATE_U: pair_update<SPZ_GATE_U>(s, a, b, c, d); break;
    default: break;
    }
}

// One amplitude of a diagonal gate: `hi` says whether this amplitude has target bit 1.
// (RZ s0 factor d0, s1 factor d1; Z/P act on target-bit-1 amplitudes only.)
__device__ __forceinline__ void diag_update_rt(int kind, const double *__restrict__ s, bool hi, double &x,
                                               double &y) {
    double dummy_a = 0.0, dummy_b = 0.0;
    if (hi) {
        pair_update_rt(kind, s, dummy_a, dummy_b, x, y);
    } else if (kind == SPZ_GATE_RZ) {
        double c = 0.0, d = 0.0;
        pair_update<SPZ_GATE_RZ>(s, x, y, c, d);
    }
}

// ---- index helpers --------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t insert_zero(uint64_t x, int pos) {
    const uint64_t lo = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | lo;
}

} // namespace spz
