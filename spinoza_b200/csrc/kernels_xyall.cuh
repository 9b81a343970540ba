// kernels_xyall.cuh -- <X_t> or <Y_t> of up to twelve qubits in one read pass.
//
// xyz_expectation_value('x' | 'y', &state, targets) (core.rs:222-264) pairs the amplitudes across the target bit:
// <X_t> = 2 sum (a c + b d), <Y_t> = 2 sum (a d - b c) over the pairs s0 = (a, b), s1 = s0 | 2^t = (c, d).  The one-target
// kernel (k_reduce<2|3>) reads the whole state once per target.  Here a CTA stages a TILE of the state in shared memory --
// 2^L contiguous amplitudes times the 2^H combinations of up to six arbitrary higher qubits, L + H <= 12, like the tiles of the
// fused gate kernels -- and every qubit of the tile gets its pair sum from that one copy: 16 * 2^n bytes of HBM for up to
// twelve targets instead of for each.  The pair sums are shared-memory bound (64 KB of shared-memory reads per target and
// tile): ~2 x the streaming time for twelve targets, against 12 x for twelve passes.
#pragma once

#include "gate_math.cuh"

namespace spz {

constexpr int kXYBits = 12;                 // tile bits
constexpr int kXYThreads = 256;
constexpr unsigned kXYTile = 1u << kXYBits; // amplitudes per full tile: 32 KB per array

struct XYArgs {
    const double *re;
    const double *im;
    int L, H;                // the tile: index bits 0..L-1 and qubits high[0..H) (ascending, all >= L)
    unsigned char high[8];
    unsigned tmask;          // tile bits whose pair sums are wanted
    int obs;                 // 0 = X, 1 = Y
    long long n_tiles;
    double *partials;        // [gridDim.x][kXYBits]
};

// hoff[s]: offset of the tile's segment s -- the combination s of the high tile qubits, deposited at their positions.  Built
// once per CTA (2^H <= 64 entries of shared memory) by xy_prepare, before the first tile.
__device__ __forceinline__ void xy_prepare(const XYArgs &a, unsigned long long *hoff) {
    for (unsigned s = threadIdx.x; s < (1u << a.H); s += blockDim.x) {
        unsigned long long o = 0;
        for (int k = 0; k < a.H; ++k) o |= (unsigned long long)((s >> k) & 1u) << a.high[k];
        hoff[s] = o;
    }
    __syncthreads();
}
// One thread's share of one tile: stage it (all threads, then a barrier), accumulate the pair sums of the wanted tile bits.
// sre / sim: 2^(L+H) doubles each.  acc[b] += this thread's part of sum (a c + b d) or sum (a d - b c) over the pairs of tile bit b.
__device__ __forceinline__ void xy_tile_accumulate(const XYArgs &a, long long tile, const unsigned long long *hoff, double *sre, double *sim,
                                                   double (&acc)[kXYBits]) {
    const int tb = a.L + a.H;
    const unsigned len = 1u << tb, lmask = (1u << a.L) - 1u;
    // first amplitude of the tile: the tile number fills the index bits that are not tile bits
    unsigned long long base = (unsigned long long)tile << a.L;
    for (int k = 0; k < a.H; ++k) base = insert_zero(base, a.high[k]);
    for (unsigned j = threadIdx.x; j < len; j += blockDim.x) {
        const unsigned long long g = base + hoff[j >> a.L] + (j & lmask);
        sre[j] = a.re[g];
        sim[j] = a.im[g];
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < kXYBits; ++b) {
        if (((a.tmask >> b) & 1u) && b < tb) { // (uniform over the CTA)
            double s = 0.0;
            const unsigned low = (1u << b) - 1u;
            for (unsigned p = threadIdx.x; p < len / 2; p += blockDim.x) {
                const unsigned s0 = ((p & ~low) << 1) | (p & low), s1 = s0 | (1u << b); // a zero bit inserted at position b
                const double x0 = sre[s0], y0 = sim[s0], x1 = sre[s1], y1 = sim[s1];
                s += a.obs == 0 ? (x0 * x1 + y0 * y1) : (x0 * y1 - y0 * x1);
            }
            acc[b] += s;
        }
    }
    __syncthreads(); // the next tile overwrites the staging arrays
}

} // namespace spz
