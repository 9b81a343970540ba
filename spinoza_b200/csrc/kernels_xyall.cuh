// kernels_xyall.cuh -- <X_t> or <Y_t> of up to twelve qubits in one read pass.
//
// xyz_expectation_value('x' | 'y', &state, targets) (core.rs:222-264) pairs the amplitudes across the target bit:
// <X_t> = 2 sum (a c + b d), <Y_t> = 2 sum (a d - b c) over the pairs s0 = (a, b), s1 = s0 | 2^t = (c, d).  The one-target
// kernel (k_reduce<2|3>) reads the whole state once per target.  Here a CTA stages a TILE of the state in shared memory --
// 2^L contiguous amplitudes times the 2^H combinations of up to six arbitrary higher qubits, L + H <= 12, like the tiles of the
// fused gate kernels -- and every qubit of the tile gets its pair sum from that one copy: 16 * 2^n bytes of HBM for up to
// twelve targets instead of for each.  The sums are taken from registers, four tile bits per shared-memory read of the tile.
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
    // Staging: two amplitudes (128 bits) per load, and the eight loads of a trip are all issued before the first
    // shared-memory store -- a load that waits for the store of the one before it keeps 8 bytes per thread in flight, and the
    // pass then runs at a fifth of the HBM rate (measured: 13.8 ms for twelve targets at 30 qubits against 2.4 ms of streaming).
    const double2 *r2 = reinterpret_cast<const double2 *>(a.re);
    const double2 *m2 = reinterpret_cast<const double2 *>(a.im);
    double2 *sre2 = reinterpret_cast<double2 *>(sre), *sim2 = reinterpret_cast<double2 *>(sim);
    for (unsigned c = 0; c < len / 2; c += 4u * blockDim.x) {
        double2 tr[4], ti[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned j2 = c + threadIdx.x + (unsigned)u * blockDim.x; // the pair of amplitudes 2 j2, 2 j2 + 1 of the tile
            if (j2 < len / 2) {
                const unsigned j = 2u * j2;
                const unsigned long long g = base + hoff[j >> a.L] + (j & lmask); // even: L >= 1 and the tile starts at a multiple of 2^L
                tr[u] = r2[g >> 1];
                ti[u] = m2[g >> 1];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned j2 = c + threadIdx.x + (unsigned)u * blockDim.x;
            if (j2 < len / 2) { sre2[j2] = tr[u]; sim2[j2] = ti[u]; }
        }
    }
    __syncthreads();
    // Pair sums from registers, four tile bits at a time: a (virtual) thread holds the 16 amplitudes that differ in the bits
    // lo .. lo+3 and takes the pair sums of those four bits from them -- one shared-memory read per pair sum instead of four.
    // For the lowest group the 16 amplitudes of a thread are contiguous, so neighbouring lanes would meet in the same banks:
    // lane v reads element k ^ (v & 15) into slot k.  Slots that differ in bit b still hold a pair of bit b; only which member
    // has the bit set is swapped when bit b of v is set -- irrelevant for X (symmetric), a sign for Y (antisymmetric).
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const int lo = 4 * g;
        if (lo >= tb || !((a.tmask >> lo) & 15u)) continue; // (uniform over the CTA)
        const int gb = tb - lo < 4 ? tb - lo : 4;
        const unsigned cnt = 1u << gb;
        for (unsigned v = threadIdx.x; v < (len >> gb); v += blockDim.x) {
            const unsigned base_idx = ((v >> lo) << (lo + gb)) | (v & ((1u << lo) - 1u));
            const unsigned m = g == 0 ? (v & (cnt - 1u)) : 0u;
            double xr[16], xi[16];
#pragma unroll
            for (unsigned k = 0; k < 16; ++k) {
                if (k < cnt) {
                    const unsigned j = base_idx | ((k ^ m) << lo);
                    xr[k] = sre[j];
                    xi[k] = sim[j];
                } else {
                    xr[k] = 0.0; xi[k] = 0.0;
                }
            }
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                if (bb < gb && ((a.tmask >> (lo + bb)) & 1u)) {
                    double s = 0.0;
#pragma unroll
                    for (unsigned k = 0; k < 16; ++k) {
                        if (!(k & (1u << bb)) && k < cnt) {
                            const unsigned k1 = k | (1u << bb);
                            if (a.obs == 0) { s = fma(xr[k], xr[k1], s); s = fma(xi[k], xi[k1], s); }
                            else { s = fma(xr[k], xi[k1], s); s = fma(-xi[k], xr[k1], s); }
                        }
                    }
                    if (a.obs != 0 && ((m >> bb) & 1u)) s = -s;
                    acc[lo + bb] += s;
                }
            }
        }
    }
    __syncthreads(); // the next tile overwrites the staging arrays
}

} // namespace spz
