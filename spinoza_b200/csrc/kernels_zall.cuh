// kernels_zall.cuh -- <Z_t> (and prob0) of EVERY qubit of the register in one read pass.
//
// xyz_expectation_value('z', &state, targets) (core.rs:222-264) costs the reference two clones and a gate per target; the
// one-target kernel here (k_reduce<4>) costs one read pass per target.  <Z_t> = total - 2 S1[t] with S1[t] the probability
// mass of the amplitudes whose index has bit t set, so ONE pass that keeps a running S1 for every bit serves any list of
// targets: 16 * 2^n bytes read once instead of once per target (30 qubits: 2.4 ms instead of 30 x 2.4 ms).
//
// A thread reads vectors of four consecutive amplitudes (two 128-bit loads per array).  Inside a vector bits 0 and 1 vary
// (two partial sums); every higher bit is constant over the vector, so the vector's mass goes to S1[t] whole or not at all:
// one select + one add per bit and vector, ~8 FP64 adds per amplitude -- far below the time the loads take.
#pragma once

#include "gate_math.cuh"

namespace spz {

// (kZMaxBits, the largest register the pass serves: engine.h)

// One thread's share of the pass: vectors first, first + stride, ... below nvec.  NB: compile-time bound on the qubit count.
// The stride must be a multiple of 2^kZThreadBits vectors (CTAs of 256 threads): then the low kZThreadBits bits of a thread's
// vector numbers -- index bits 2 .. 2 + kZThreadBits - 1 -- never change, their share of the thread's mass is decided once after
// the loop, and only the bits above them need a running sum (2 * (NB - 10) + 6 accumulator registers).
constexpr int kZThreadBits = 8;
template <int NB>
__device__ __forceinline__ void z_all_accumulate(const double *__restrict__ re, const double *__restrict__ im, long long nvec,
                                                 int n, long long first, long long stride, double &total, double (&s1)[NB]) {
    constexpr int T0 = 2 + kZThreadBits; // first index bit that varies along a thread's loop
    const double2 *r2 = reinterpret_cast<const double2 *>(re);
    const double2 *m2 = reinterpret_cast<const double2 *>(im);
    for (long long v = first; v < nvec; v += 2 * stride) {
        // two vectors per trip, all eight loads in flight before the arithmetic
        const long long w = v + stride;
        const bool two = w < nvec;
        const double2 ra = r2[2 * v], rb = r2[2 * v + 1], ma = m2[2 * v], mb = m2[2 * v + 1];
        double2 rc = make_double2(0.0, 0.0), rd = rc, mc = rc, md = rc;
        if (two) { rc = r2[2 * w]; rd = r2[2 * w + 1]; mc = m2[2 * w]; md = m2[2 * w + 1]; }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double2 x0 = h ? rc : ra, x1 = h ? rd : rb, y0 = h ? mc : ma, y1 = h ? md : mb;
            const unsigned long long vv = (unsigned long long)(h ? w : v);
            const double p0 = x0.x * x0.x + y0.x * y0.x, p1 = x0.y * x0.y + y0.y * y0.y;
            const double p2 = x1.x * x1.x + y1.x * y1.x, p3 = x1.y * x1.y + y1.y * y1.y;
            const double ps = (p0 + p1) + (p2 + p3);
            total += ps;
            s1[0] += p1 + p3;
            s1[1] += p2 + p3;
#pragma unroll
            for (int t = T0; t < NB; ++t)
                if (t < n) s1[t] += ((vv >> (t - 2)) & 1ull) ? ps : 0.0; // (the second vector of a trip past the end holds zeros)
        }
    }
#pragma unroll
    for (int t = 2; t < T0 && t < NB; ++t) s1[t] = (((unsigned long long)first >> (t - 2)) & 1ull) ? total : 0.0;
}

} // namespace spz
