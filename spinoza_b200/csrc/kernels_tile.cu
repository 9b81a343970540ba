// kernels_tile.cu -- fused execution: a run of gates applied tile-by-tile in shared memory so that one
// HBM pass (32 * 2^n bytes) serves the whole run.  This is the engine behind spz_execute(SPZ_EXEC_FUSE);
// the reference applies one gate per full pass (circuit.rs:553-599, "no fusion").
//
// A tile is the set of 2^T amplitudes that share all index bits OUTSIDE the tile's bit set
//   { 0 .. L-1 }  U  { high[0] < high[1] < ... }         (T = L + n_high)
// i.e. 2^n_high contiguous segments of 2^L amplitudes each.  Any gate whose target is a tile qubit can be
// applied inside the tile; controls may be anywhere (outer controls are constant per tile, so the CTA skips
// the gate or applies it unconditionally); diagonal gates (Z, P, RZ) may even target an outer qubit, where
// they reduce to a per-tile constant factor.  Gates are applied one after another with the same per-pair
// arithmetic as the unfused kernels (gate_math.cuh), so fused results are bit-identical to unfused ones.
//
// One CTA per tile: coalesced 128-bit loads of every segment into shared memory (SoA: re[2^T], im[2^T]),
// the op list applied with a __syncthreads between ops, coalesced stores back.  T = 12 -> 64 KB of shared
// memory per CTA, 3 CTAs per SM so loads of one tile overlap the arithmetic of another.
#include <algorithm>
#include <vector>

#include "gate_math.cuh"

namespace spz {

constexpr int kTileThreads = 256;
constexpr int kMaxTileBits = 12;
constexpr int kMaxHigh = 8;

int max_tile_bits() { return kMaxTileBits; }

struct TileArgs {
    double *re;
    double *im;
    const TileOp *ops;
    int n_ops;
    int T, L, n_high;
    int high[kMaxHigh];
};

// deposit the bits of x into the positions of the zero bits... insert a zero at every set bit of `mask`
// (ascending), leaving room for the tile bits.
__device__ __forceinline__ unsigned insert_zeros_mask(unsigned x, unsigned mask) {
    while (mask) {
        const int b = __ffs(mask) - 1;
        x = (unsigned)insert_zero(x, b);
        mask &= mask - 1;
    }
    return x;
}

__global__ void __launch_bounds__(kTileThreads) k_tile(const TileArgs a) {
    extern __shared__ double smem[];
    const int T = a.T, L = a.L;
    const unsigned tile_len = 1u << T;
    double *sre = smem;
    double *sim = smem + tile_len;
    __shared__ unsigned long long seg_off[1 << kMaxHigh];

    // absolute index of the tile's first amplitude: CTA id bits go to the non-tile positions
    unsigned long long base = (unsigned long long)blockIdx.x << L;
    for (int k = 0; k < a.n_high; ++k) base = insert_zero(base, a.high[k]);

    const int n_seg = 1 << a.n_high;
    for (int sgi = threadIdx.x; sgi < n_seg; sgi += kTileThreads) {
        unsigned long long off = 0;
        for (int k = 0; k < a.n_high; ++k)
            if ((sgi >> k) & 1) off |= 1ull << a.high[k];
        seg_off[sgi] = off;
    }
    __syncthreads();

    // ---- load: 2^(T-1) double2 vectors per array ----
    const unsigned n_vec = tile_len >> 1;
    const unsigned seg_mask = (1u << L) - 1u;
    for (unsigned v = threadIdx.x; v < n_vec; v += kTileThreads) {
        const unsigned j = v << 1;
        const unsigned long long g = base + seg_off[j >> L] + (j & seg_mask);
        const double2 r = *reinterpret_cast<const double2 *>(a.re + g);
        const double2 m = *reinterpret_cast<const double2 *>(a.im + g);
        *reinterpret_cast<double2 *>(sre + j) = r;
        *reinterpret_cast<double2 *>(sim + j) = m;
    }
    __syncthreads();

    // ---- apply the op list ----
    for (int oi = 0; oi < a.n_ops; ++oi) {
        const TileOp &op = a.ops[oi];
        const unsigned long long ocm = op.outer_cmask;
        if ((base & ocm) != ocm) continue; // an outer control is 0 for this whole tile
        const int kind = op.kind;
        const unsigned icm = op.inner_cmask;
        if (kind == SPZ_GATE_SWAP) {
            // swap_apply gates.rs:1376-1386 restricted to the tile: (lo=1,hi=0) <-> (lo=0,hi=1)
            const unsigned blo = 1u << min(op.tbit, op.tbit2), bhi = 1u << max(op.tbit, op.tbit2);
            const unsigned ins = blo | bhi;
            const unsigned cnt = tile_len >> 2;
            for (unsigned p = threadIdx.x; p < cnt; p += kTileThreads) {
                const unsigned x = insert_zeros_mask(p, ins);
                const unsigned ia = x | blo, ib = x | bhi;
                double t = sre[ia]; sre[ia] = sre[ib]; sre[ib] = t;
                t = sim[ia]; sim[ia] = sim[ib]; sim[ib] = t;
            }
        } else if (op.tbit >= 0) {
            const unsigned tb = 1u << op.tbit;
            const unsigned ins = tb | icm;
            const unsigned cnt = tile_len >> __popc(ins);
            const bool s0_too = !(kind == SPZ_GATE_Z || kind == SPZ_GATE_P);
            for (unsigned p = threadIdx.x; p < cnt; p += kTileThreads) {
                const unsigned i0 = insert_zeros_mask(p, ins) | icm;
                const unsigned i1 = i0 | tb;
                double x0 = 0.0, y0 = 0.0;
                if (s0_too) { x0 = sre[i0]; y0 = sim[i0]; }
                double x1 = sre[i1], y1 = sim[i1];
                pair_update_rt(kind, op.s, x0, y0, x1, y1);
                if (s0_too) { sre[i0] = x0; sim[i0] = y0; }
                sre[i1] = x1; sim[i1] = y1;
            }
        } else {
            // diagonal gate whose target is an outer qubit: constant factor for this tile
            const bool hi = op.const_hi ? (op.const_hi == 2) : (bool)((base >> op.outer_target) & 1ull);
            if (!hi && kind != SPZ_GATE_RZ) continue; // Z / P leave target-bit-0 amplitudes alone
            const unsigned cnt = tile_len >> __popc(icm);
            for (unsigned p = threadIdx.x; p < cnt; p += kTileThreads) {
                const unsigned i = insert_zeros_mask(p, icm) | icm;
                double x = sre[i], y = sim[i];
                diag_update_rt(kind, op.s, hi, x, y);
                sre[i] = x; sim[i] = y;
            }
        }
        __syncthreads();
    }

    // ---- store ----
    for (unsigned v = threadIdx.x; v < n_vec; v += kTileThreads) {
        const unsigned j = v << 1;
        const unsigned long long g = base + seg_off[j >> L] + (j & seg_mask);
        *reinterpret_cast<double2 *>(a.re + g) = *reinterpret_cast<const double2 *>(sre + j);
        *reinterpret_cast<double2 *>(a.im + g) = *reinterpret_cast<const double2 *>(sim + j);
    }
}

int launch_tile_group(spz_state *st, const TilePlan &plan, const TileOp *ops, int n_ops) {
    if (n_ops <= 0) return SPZ_OK;
    if (plan.tile_bits > kMaxTileBits || plan.n_high > kMaxHigh || plan.low_bits < 1 ||
        plan.tile_bits != plan.low_bits + plan.n_high || plan.tile_bits > st->n) {
        set_error("bad tile plan T=%d L=%d H=%d n=%d", plan.tile_bits, plan.low_bits, plan.n_high, st->n);
        return SPZ_ERR_INVALID_ARG;
    }
    // Op lists are staged in a device ring buffer: a group's list must stay intact until its kernel has
    // run, so the cursor only wraps after a stream synchronise.
    const size_t bytes = (sizeof(TileOp) * (size_t)n_ops + 255) & ~(size_t)255;
    if (st->d_ops_bytes < bytes || !st->d_ops) {
        if (st->d_ops) { SPZ_CUDA(cudaStreamSynchronize(st->stream)); SPZ_CUDA(cudaFree(st->d_ops)); st->d_ops = nullptr; }
        const size_t cap = std::max<size_t>(bytes * 2, (size_t)4 << 20);
        SPZ_CUDA(cudaMalloc(&st->d_ops, cap));
        st->d_ops_bytes = cap;
        st->d_ops_cursor = 0;
    }
    if (st->d_ops_cursor + bytes > st->d_ops_bytes) {
        SPZ_CUDA(cudaStreamSynchronize(st->stream));
        st->d_ops_cursor = 0;
    }
    char *slot = static_cast<char *>(st->d_ops) + st->d_ops_cursor;
    st->d_ops_cursor += bytes;
    SPZ_CUDA(cudaMemcpyAsync(slot, ops, sizeof(TileOp) * (size_t)n_ops, cudaMemcpyHostToDevice, st->stream));

    TileArgs a{};
    a.re = st->re; a.im = st->im;
    a.ops = reinterpret_cast<const TileOp *>(slot);
    a.n_ops = n_ops;
    a.T = plan.tile_bits; a.L = plan.low_bits; a.n_high = plan.n_high;
    for (int k = 0; k < plan.n_high; ++k) a.high[k] = plan.high[k];
    const size_t smem = sizeof(double) * 2u * ((size_t)1 << plan.tile_bits);
    static bool attr_set[64] = {false};
    if (!attr_set[st->device & 63]) {
        SPZ_CUDA(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(double) * 2u * ((size_t)1 << kMaxTileBits))));
        attr_set[st->device & 63] = true;
    }
    const unsigned grid = (unsigned)((uint64_t)st->len >> plan.tile_bits);
    k_tile<<<grid, kTileThreads, smem, st->stream>>>(a);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

} // namespace spz
