// kernels_tile3.cu -- the fused tile kernel of merged (default) execution, third generation: the Blackwell-native one.
//
// Same contract as k_tile (kernels_tile.cu): one CTA applies a whole run of gates to one 2^12-amplitude tile (2^n_high
// segments of 2^L contiguous amplitudes, 64 KB of re + im), so one HBM pass serves the run instead of one pass per gate
// (the reference's execute loop, circuit.rs:553-599).  What is different:
//
//   * The tile moves by TMA.  Every segment is one `cp.async.bulk.tensor.2d` box of a 2-D view [2^n / 16 rows x 16 doubles]
//     of the state array (SASS: UTMALDG / UTMASTG), landing in shared memory under the hardware 128-byte swizzle; the CTA
//     waits on one mbarrier, works between shared memory and registers, and writes the tile back with TMA stores.  There is
//     no global->register->shared staging code and no per-thread global addressing at all.  While the boxes are in flight
//     the CTA stages its program and reduces the per-tile phase constants.
//   * The program is lowered on the host (tile3_lower) from the scheduler's micro-program (TileInstr, engine.h) into a
//     16-byte instruction format the device decodes with two shared-memory loads, with everything that does not depend on
//     the data decided on the host: which phase accumulators are pending at each butterfly / layout change, which gate
//     variant runs, where its scalars live.
//   * Merged arithmetic is contracted and rescaled.  Uncontrolled H, RX, RY factor their common scalar out of the 2x2
//     matrix (H = 2^-1/2 [[1,1],[1,-1]], RX = cos [[1, i tan],[i tan, 1]], RY = cos [[1, -tan],[tan, 1]]); the product of
//     those scalars over the pass multiplies the F0 phase accumulator once.  A butterfly then costs 4 (H: add/sub) or
//     4 FMA (RX, RY) per amplitude pair instead of 8-12 uncontracted multiplies and adds.  This is the documented merged-mode
//     rounding difference: results agree with gate-by-gate application to ~1e-15 relative, tests hold 1e-12.
//   * Diagonal runs use tables.  A run of diagonal gates multiplies five per-thread accumulators (F0: all 16 register
//     amplitudes, F1..F4: those whose register bit is set).  Terms that depend on thread bits inside one nibble of the
//     thread id are folded by the host into two 16-entry tables per (run, accumulator); terms that depend on bits outside
//     the tile are reduced once per CTA to one constant per (run, accumulator).  An accumulator update is then two table
//     look-ups and at most three complex multiplies, whatever the number of gates in the run (a QFT pass has ~20 per run).
//
// Exact mode (SPZ_EXEC_EXACT, bit-identical to unfused), registers below 12 qubits and programs that do not fit the
// shared-memory budget stay on k_tile.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef SPZ_CPU_EMULATION
#include <cuda.h> // CUtensorMap and its enums (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#endif

#include "gate_math.cuh"

#ifdef SPZ_CPU_EMULATION
#define SPZ_DYN_SMEM3(name) unsigned char *name = spz_emu::dyn_smem
#else
#define SPZ_DYN_SMEM3(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

namespace spz {

namespace {

constexpr int kT3 = 12;                    // tile bits
constexpr int kThreads3 = 256;             // 2^(kT3 - 4): 16 amplitudes per thread
constexpr int kMaxHigh3 = 8;
constexpr unsigned kTileLen3 = 1u << kT3;
constexpr unsigned kArrayBytes3 = kTileLen3 * 8u;   // 32 KB per array
constexpr unsigned kAccBytes3 = 5u * kThreads3 * 16u; // the five phase accumulators of every thread (double2 each): 20 KB
constexpr unsigned kProgOff3 = 2u * kArrayBytes3 + kAccBytes3;
constexpr size_t kSmemBudget3 = 112 * 1024;         // two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)
constexpr int kMaxIns3 = 1024;                      // one skip flag per instruction, one byte each
constexpr int kMaxGroups3 = 1024;

// ---- device program ---------------------------------------------------------------------------------------------------
enum { T3_END = 0, T3_LAYOUT = 1, T3_GATE = 2, T3_ACC = 3, T3_ACCG = 4, T3_OTHER = 5 };
// gate variants of merged mode (see the header): *T = common scalar factored out (uncontrolled only)
enum { MK_H = 0, MK_RXT = 1, MK_RYT = 2, MK_X = 3, MK_Y = 4, MK_REAL = 5, MK_RXU = 6, MK_U = 7 };
// flag bits
enum {
    GF_ALL = 1,      // GATE: every pair of every thread (no in-tile control)
    GF_PRE = 2,      // GATE: accumulator F_{rpos+1} is pending: apply it to the amplitudes with the target bit set first
    GF_OUTER = 4,    // GATE: has controls outside the tile: consult the per-tile skip flag
    AF_LO = 1,       // ACC: table over the low nibble of the thread id at pool2[a .. a+16)
    AF_HI = 2,       // ACC: table over the high nibble at pool2[a+16 .. a+32)
    AF_TILE = 4,     // ACC / ACCG / OTHER: per-tile constant gfac[b]
    AF_SET = 8,      // ACC / ACCG: the accumulator holds no pending factor: assign instead of multiply
    AF_CONST = 16,   // ACCG / OTHER: constant factor at pool2[a]
};
struct Ins3 {
    uint8_t op, kind, rpos, flags; // rpos: GATE register bit; ACC / ACCG accumulator 0..4.  LAYOUT / END flags: pending accumulators
    uint16_t km;                   // GATE: pair mask over k0.  OTHER: register mask m
    uint16_t thr;                  // GATE / ACCG / OTHER: control bits in THREAD-ID space (8 bits)
    uint32_t a;                    // LAYOUT: the 4 register-resident tile bits, one byte each.  else: pool offset (see flags)
    uint32_t b;                    // per-tile constant index
};
static_assert(sizeof(Ins3) == 16, "Ins3 is decoded with one 128-bit shared-memory load");

// What the host uploads for one pass (one contiguous blob, 16-byte aligned sections):
//   Ins3 ins[n_ins] | double pool[n_pool] (gate scalars, tables, constants; double2 entries at even offsets)
//   | uint64 outer[n_ins] (controls outside the tile, GATE only) | TileGroup groups[n_groups] | TileTerm terms[n_terms]
// Shared memory of a CTA: tile re | tile im | accumulators F[5][256] | ins | pool | gfac[n_groups] | skip[n_ins] | mbarrier
struct Lowered3 {
    std::vector<Ins3> ins;
    std::vector<double> pool;
    std::vector<uint64_t> outer;
    std::vector<TileGroup> groups; // thr, m unused; first / count index `terms`
    std::vector<TileTerm> terms;   // outer, fr, fi used
    double scale = 1.0;            // product of the factored-out gate scalars: initial value of F0
    bool ctrl = false;             // some butterfly has an in-tile control
};

struct Tile3Args {
#ifndef SPZ_CPU_EMULATION
    CUtensorMap tm_re, tm_im;      // [len / 16 rows x 16 doubles], box 16 x 2^(L-4), 128-byte swizzle
#endif
    double *re, *im;               // (the CPU emulation moves the tile through these)
    const unsigned char *blob;     // device copy of the lowered program
    unsigned ins_bytes, pool_bytes;  // sizes of the two staged sections (multiples of 16)
    unsigned outer_off, groups_off, terms_off; // byte offsets of the other sections inside the blob
    int n_ins, n_groups;
    unsigned tile_offset;
    int L, n_high;
    double scale;
    int high[kMaxHigh3];
};

// position of tile index j inside a 32 KB array under the TMA 128-byte swizzle (16-byte chunk index ^= 128-byte row
// index mod 8); linear over GF(2), so swz3(a ^ b) == swz3(a) ^ swz3(b)
__host__ __device__ __forceinline__ unsigned swz3(unsigned j) { return j ^ (((j >> 4) & 7u) << 1); }

__device__ __forceinline__ void cmul3(double &xr, double &xi, double fr, double fi) {
    const double nr = xr * fr - xi * fi; // contracted by nvcc: 2 FMA-class instructions per component
    const double ni = xr * fi + xi * fr;
    xr = nr; xi = ni;
}

// ---- TMA / mbarrier primitives (sm_90+ PTX); the CPU emulation replaces them by synchronous copies ----------------------
#ifndef SPZ_CPU_EMULATION
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *tm, int row, void *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap *tm, int row, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// ---- gate variants ------------------------------------------------------------------------------------------------------
// (a, b) = (re, im) of the amplitude with target bit 0, (c, d) with target bit 1.  Plain operators: nvcc contracts them.
template <int MK>
__device__ __forceinline__ void pair3(const double (&s)[8], double &a, double &b, double &c, double &d) {
    if constexpr (MK == MK_H) {            // [[1, 1], [1, -1]]
        const double na = a + c, nb = b + d, nc = a - c, nd = b - d;
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (MK == MK_RXT) {   // [[1, i t], [i t, 1]]
        const double t = s[0];
        const double na = a - t * d, nb = b + t * c, nc = c - t * b, nd = d + t * a;
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (MK == MK_RYT) {   // [[1, -t], [t, 1]]
        const double t = s[0];
        const double na = a - t * c, nb = b - t * d, nc = c + t * a, nd = d + t * b;
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (MK == MK_X) {
        double t = a; a = c; c = t;
        t = b; b = d; d = t;
    } else if constexpr (MK == MK_Y) {     // [[0, -i], [i, 0]]
        const double na = d, nb = -c, nc = -b, nd = a;
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (MK == MK_REAL) {  // real [[s0, s1], [s2, s3]]
        const double na = s[0] * a + s[1] * c, nb = s[0] * b + s[1] * d;
        const double nc = s[2] * a + s[3] * c, nd = s[2] * b + s[3] * d;
        a = na; b = nb; c = nc; d = nd;
    } else if constexpr (MK == MK_RXU) {   // [[s0, i s1], [i s1, s0]]
        const double na = s[0] * a - s[1] * d, nb = s[0] * b + s[1] * c;
        const double nc = s[0] * c - s[1] * b, nd = s[0] * d + s[1] * a;
        a = na; b = nb; c = nc; d = nd;
    } else {                               // complex [[s0 + i s1, s2 + i s3], [s4 + i s5, s6 + i s7]]
        const double na = s[0] * a - s[1] * b + s[2] * c - s[3] * d;
        const double nb = s[0] * b + s[1] * a + s[2] * d + s[3] * c;
        const double nc = s[4] * a - s[5] * b + s[6] * c - s[7] * d;
        const double nd = s[4] * b + s[5] * a + s[6] * d + s[7] * c;
        a = na; b = nb; c = nc; d = nd;
    }
}
template <int MK>
constexpr int n_scalars3() { return MK == MK_RXT || MK == MK_RYT ? 1 : MK == MK_REAL ? 4 : MK == MK_RXU ? 2 : MK == MK_U ? 8 : 0; }

template <int MK, int R, bool ALL>
__device__ __forceinline__ void bfly3(double (&ar)[16], double (&ai)[16], const double *__restrict__ sp, unsigned km) {
    double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < n_scalars3<MK>(); ++i) s[i] = sp[i];
#pragma unroll
    for (int k0 = 0; k0 < 16; ++k0) {
        if (k0 & (1 << R)) continue;
        const int k1 = k0 | (1 << R);
        if (ALL || (km & (1u << k0))) pair3<MK>(s, ar[k0], ai[k0], ar[k1], ai[k1]); // km is the same for every thread
    }
}
template <int R>
__device__ __forceinline__ void bfly3_kind(int kind, bool all, double (&ar)[16], double (&ai)[16], const double *__restrict__ sp,
                                           unsigned km) {
    if (all) {
        switch (kind) {
        case MK_H: bfly3<MK_H, R, true>(ar, ai, sp, km); break;
        case MK_RXT: bfly3<MK_RXT, R, true>(ar, ai, sp, km); break;
        case MK_RYT: bfly3<MK_RYT, R, true>(ar, ai, sp, km); break;
        case MK_X: bfly3<MK_X, R, true>(ar, ai, sp, km); break;
        case MK_Y: bfly3<MK_Y, R, true>(ar, ai, sp, km); break;
        case MK_REAL: bfly3<MK_REAL, R, true>(ar, ai, sp, km); break;
        case MK_RXU: bfly3<MK_RXU, R, true>(ar, ai, sp, km); break;
        default: bfly3<MK_U, R, true>(ar, ai, sp, km); break;
        }
    } else {
        switch (kind) { // the factored variants are uncontrolled by construction
        case MK_X: bfly3<MK_X, R, false>(ar, ai, sp, km); break;
        case MK_Y: bfly3<MK_Y, R, false>(ar, ai, sp, km); break;
        case MK_REAL: bfly3<MK_REAL, R, false>(ar, ai, sp, km); break;
        case MK_RXU: bfly3<MK_RXU, R, false>(ar, ai, sp, km); break;
        default: bfly3<MK_U, R, false>(ar, ai, sp, km); break;
        }
    }
}
template <int R>
__device__ __forceinline__ void apply_bit3(double (&ar)[16], double (&ai)[16], const double2 f) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k & (1 << R)) cmul3(ar[k], ai[k], f.x, f.y);
}

__global__ void __launch_bounds__(kThreads3, 2) k_tile3(const
#ifndef SPZ_CPU_EMULATION
                                                        __grid_constant__
#endif
                                                        Tile3Args a) {
    SPZ_DYN_SMEM3(smem);
    double *sre = reinterpret_cast<double *>(smem);
    double *sim = reinterpret_cast<double *>(smem + kArrayBytes3);
    const Ins3 *sins = reinterpret_cast<const Ins3 *>(smem + kProgOff3);
    const double *pool = reinterpret_cast<const double *>(smem + kProgOff3 + a.ins_bytes);
    double2 *gfac = reinterpret_cast<double2 *>(smem + kProgOff3 + a.ins_bytes + a.pool_bytes);
    // F[c][tid]: accumulator c of this thread.  They live in shared memory (one 128-bit access each, conflict-free) rather
    // than in 20 registers: with 64 registers of amplitudes and up to 16 of gate scalars the budget of 128 is tight.
    double2 *facc = reinterpret_cast<double2 *>(smem + 2u * kArrayBytes3) + threadIdx.x;
    unsigned char *skip = reinterpret_cast<unsigned char *>(gfac + a.n_groups);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(
        smem + ((kProgOff3 + a.ins_bytes + a.pool_bytes + 16u * (unsigned)a.n_groups + (unsigned)a.n_ins + 15u) & ~15u));
    const unsigned tid = threadIdx.x;
    const int L = a.L;

    // absolute index of the tile's first amplitude: the CTA id fills the non-tile bit positions
    auto tile_base = [&]() -> unsigned long long {
        unsigned long long b = (unsigned long long)(blockIdx.x + a.tile_offset) << L;
#pragma unroll
        for (int k = 0; k < kMaxHigh3; ++k)
            if (k < a.n_high) b = insert_zero(b, a.high[k]);
        return b;
    };

    // ---- tile in: one TMA box per segment and array, all on one mbarrier ----
    const unsigned n_seg = 1u << a.n_high;
    const unsigned seg_bytes = 8u << L;
    auto seg_row = [&](unsigned long long base, unsigned s) -> int { // row (16 doubles) of segment s's first amplitude
        unsigned long long off = 0;
#pragma unroll
        for (int k = 0; k < kMaxHigh3; ++k)
            if (k < a.n_high && ((s >> k) & 1u)) off |= 1ull << a.high[k];
        return (int)((base + off) >> 4);
    };
    {
    const unsigned long long base = tile_base();
#ifndef SPZ_CPU_EMULATION
    if ((smem_u32(smem) & 1023u) != 0u) __trap(); // the swizzle pattern is anchored to 1 KB-aligned shared addresses
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, 2u * kArrayBytes3);
    }
    __syncthreads();
    if (tid < 32) {
        for (unsigned s = tid; s < n_seg; s += 32) {
            const int row = seg_row(base, s);
            tma_load_box(smem + s * seg_bytes, &a.tm_re, row, bar);
            tma_load_box(smem + kArrayBytes3 + s * seg_bytes, &a.tm_im, row, bar);
        }
    }
#else
    if (tid == 0) { // the emulation's "TMA": rows of 16 doubles, 16-byte chunks XORed with the row number mod 8
        for (unsigned s = 0; s < n_seg; ++s) {
            const unsigned long long g0 = (unsigned long long)seg_row(base, s) << 4;
            for (unsigned j = 0; j < (1u << L); ++j) {
                const unsigned tj = (s << L) + j;
                sre[swz3(tj)] = a.re[g0 + j];
                sim[swz3(tj)] = a.im[g0 + j];
            }
        }
    }
#endif

    // ---- while the tile is in flight: stage the program, skip flags, per-tile constants ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem + kProgOff3);
        const unsigned n16 = (a.ins_bytes + a.pool_bytes) >> 4;
        for (unsigned i = tid; i < n16; i += kThreads3) dst[i] = src[i];
    }
    {
        const unsigned long long *outer = reinterpret_cast<const unsigned long long *>(a.blob + a.outer_off);
        for (int i = tid; i < a.n_ins; i += kThreads3) {
            const unsigned long long ocm = outer[i];
            skip[i] = (base & ocm) != ocm ? 1 : 0;
        }
        const TileGroup *groups = reinterpret_cast<const TileGroup *>(a.blob + a.groups_off);
        const TileTerm *terms = reinterpret_cast<const TileTerm *>(a.blob + a.terms_off);
        for (int g = tid; g < a.n_groups; g += kThreads3) {
            const TileGroup gd = groups[g];
            double fr = 1.0, fi = 0.0;
            for (int i = 0; i < gd.count; ++i) {
                const TileTerm t = terms[gd.first + i];
                if ((base & t.outer) == t.outer) cmul3(fr, fi, t.fr, t.fi);
            }
            gfac[g] = make_double2(fr, fi);
        }
    }
    } // base
    __syncthreads();
#ifndef SPZ_CPU_EMULATION
    mbar_wait(bar, 0);
#endif

    // ---- registers ----
    double ar[16], ai[16];
    facc[0] = make_double2(a.scale, 0.0); // F0 starts as the pass scale (pending from the start when it is not 1: the host knows)
    // The register layout is carried as one word (4 tile bits, one byte each) and expanded where it is used: the swizzled
    // position of amplitude k of this thread is stj ^ (sw0 if k & 1) ^ (sw1 if k & 2) ^ ...  (swz3 is linear over GF(2)).
    unsigned lay = sins[0].a;
    auto move_regs = [&](auto &&xfer2, auto &&xfer1) {
        const int r0 = lay & 255u, r1 = (lay >> 8) & 255u, r2 = (lay >> 16) & 255u, r3 = lay >> 24;
        const unsigned stj = swz3((unsigned)insert_zero(insert_zero(insert_zero(insert_zero(tid, r0), r1), r2), r3));
        const unsigned sw0 = swz3(1u << r0), sw1 = swz3(1u << r1), sw2 = swz3(1u << r2), sw3 = swz3(1u << r3);
        if (r0 == 0) { // register bit 0 is tile bit 0: amplitudes k, k + 1 are adjacent in shared memory (128-bit accesses)
#pragma unroll
            for (int k = 0; k < 16; k += 2) xfer2(k, stj ^ ((k & 2) ? sw1 : 0u) ^ ((k & 4) ? sw2 : 0u) ^ ((k & 8) ? sw3 : 0u));
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) xfer1(k, stj ^ ((k & 1) ? sw0 : 0u) ^ ((k & 2) ? sw1 : 0u) ^ ((k & 4) ? sw2 : 0u) ^ ((k & 8) ? sw3 : 0u));
        }
    };
    auto load_regs = [&]() {
        move_regs(
            [&](int k, unsigned s) {
                const double2 r = *reinterpret_cast<const double2 *>(sre + s);
                const double2 m = *reinterpret_cast<const double2 *>(sim + s);
                ar[k] = r.x; ar[k + 1] = r.y; ai[k] = m.x; ai[k + 1] = m.y;
            },
            [&](int k, unsigned s) { ar[k] = sre[s]; ai[k] = sim[s]; });
    };
    auto store_regs = [&]() {
        move_regs(
            [&](int k, unsigned s) {
                *reinterpret_cast<double2 *>(sre + s) = make_double2(ar[k], ar[k + 1]);
                *reinterpret_cast<double2 *>(sim + s) = make_double2(ai[k], ai[k + 1]);
            },
            [&](int k, unsigned s) { sre[s] = ar[k]; sim[s] = ai[k]; });
    };
    // apply the pending accumulators named by `mask` (the host knows which are pending: it is a property of the program,
    // so nothing is ever reset: an accumulator that has been applied is assigned, not multiplied, by its next update)
    auto flush = [&](unsigned mask) {
        if (!mask) return;
        if (__popc(mask) <= 2) {
            if (mask & 1u) {
                const double2 f = facc[0];
#pragma unroll
                for (int k = 0; k < 16; ++k) cmul3(ar[k], ai[k], f.x, f.y);
            }
            if (mask & 2u) apply_bit3<0>(ar, ai, facc[1 * kThreads3]);
            if (mask & 4u) apply_bit3<1>(ar, ai, facc[2 * kThreads3]);
            if (mask & 8u) apply_bit3<2>(ar, ai, facc[3 * kThreads3]);
            if (mask & 16u) apply_bit3<3>(ar, ai, facc[4 * kThreads3]);
            return;
        }
        // three or more: expand the 16 per-amplitude factors (15 complex multiplies) and apply them once
        const double2 one = make_double2(1.0, 0.0);
        const double2 f0 = (mask & 1u) ? facc[0] : one, f1 = (mask & 2u) ? facc[1 * kThreads3] : one;
        const double2 f2 = (mask & 4u) ? facc[2 * kThreads3] : one, f3 = (mask & 8u) ? facc[3 * kThreads3] : one;
        const double2 f4 = (mask & 16u) ? facc[4 * kThreads3] : one;
#pragma unroll
        for (int b3 = 0; b3 < 2; ++b3) {
            double g3r = f0.x, g3i = f0.y;
            if (b3) cmul3(g3r, g3i, f4.x, f4.y);
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
                double g2r = g3r, g2i = g3i;
                if (b2) cmul3(g2r, g2i, f3.x, f3.y);
#pragma unroll
                for (int b1 = 0; b1 < 2; ++b1) {
                    double g1r = g2r, g1i = g2i;
                    if (b1) cmul3(g1r, g1i, f2.x, f2.y);
#pragma unroll
                    for (int b0 = 0; b0 < 2; ++b0) {
                        double gr = g1r, gi = g1i;
                        if (b0) cmul3(gr, gi, f1.x, f1.y);
                        cmul3(ar[b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)], ai[b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)], gr, gi);
                    }
                }
            }
        }
    };
    auto acc = [&](unsigned cls, bool set, double fr, double fi) {
        double2 *f = facc + cls * kThreads3;
        if (!set) { const double2 old = *f; cmul3(fr, fi, old.x, old.y); }
        *f = make_double2(fr, fi);
    };

    load_regs();

    for (int pc = 1;; ++pc) {
        const Ins3 ins = sins[pc];
        const int op = ins.op;
        if (op == T3_GATE) {
            // skipped when a control outside the tile is 0 for this whole tile, or a thread-bit control is 0 for this thread
            const bool thr_ok = (tid & ins.thr) == ins.thr && !((ins.flags & GF_OUTER) && skip[pc]);
            const double *sp = pool + ins.a;
            const bool all = ins.flags & GF_ALL;
            const bool pre = ins.flags & GF_PRE;
            // Only the accumulator of the target's own register bit separates the two members of a pair; the others
            // scale both by the same factor and stay pending.
            switch (ins.rpos) {
            case 0: if (pre) apply_bit3<0>(ar, ai, facc[1 * kThreads3]); if (thr_ok) bfly3_kind<0>(ins.kind, all, ar, ai, sp, ins.km); break;
            case 1: if (pre) apply_bit3<1>(ar, ai, facc[2 * kThreads3]); if (thr_ok) bfly3_kind<1>(ins.kind, all, ar, ai, sp, ins.km); break;
            case 2: if (pre) apply_bit3<2>(ar, ai, facc[3 * kThreads3]); if (thr_ok) bfly3_kind<2>(ins.kind, all, ar, ai, sp, ins.km); break;
            default: if (pre) apply_bit3<3>(ar, ai, facc[4 * kThreads3]); if (thr_ok) bfly3_kind<3>(ins.kind, all, ar, ai, sp, ins.km); break;
            }
            continue;
        }
        if (op == T3_ACC) {
            const double2 *tab = reinterpret_cast<const double2 *>(pool + ins.a);
            double fr = 1.0, fi = 0.0;
            if (ins.flags & AF_LO) { const double2 t = tab[tid & 15u]; fr = t.x; fi = t.y; }
            if (ins.flags & AF_HI) { const double2 t = tab[16u + (tid >> 4)]; cmul3(fr, fi, t.x, t.y); }
            if (ins.flags & AF_TILE) { const double2 t = gfac[ins.b]; cmul3(fr, fi, t.x, t.y); }
            acc(ins.rpos, ins.flags & AF_SET, fr, fi);
            continue;
        }
        if (op == T3_ACCG) { // a term that needs thread bits from both nibbles, or thread bits and bits outside the tile
            const bool hit = (tid & ins.thr) == ins.thr;
            double2 *f = facc + ins.rpos * kThreads3;
            if (ins.flags & AF_SET) {
                // the accumulator held no pending factor: every thread assigns (nothing is ever reset, see flush)
                *f = hit ? ((ins.flags & AF_TILE) ? gfac[ins.b] : *reinterpret_cast<const double2 *>(pool + ins.a)) : make_double2(1.0, 0.0);
            } else if (hit) {
                const double2 t = (ins.flags & AF_TILE) ? gfac[ins.b] : *reinterpret_cast<const double2 *>(pool + ins.a);
                acc(ins.rpos, false, t.x, t.y);
            }
            continue;
        }
        if (op == T3_OTHER) { // a diagonal term over two or more register bits: applied at once to the amplitudes it selects
            if ((tid & ins.thr) != ins.thr) continue;
            const double2 f = (ins.flags & AF_TILE) ? gfac[ins.b] : *reinterpret_cast<const double2 *>(pool + ins.a);
            const unsigned m = ins.km;
#define SPZ_M4(A, B, C, D) cmul3(ar[A], ai[A], f.x, f.y); cmul3(ar[B], ai[B], f.x, f.y); cmul3(ar[C], ai[C], f.x, f.y); cmul3(ar[D], ai[D], f.x, f.y)
            switch (m) {
            case 3: SPZ_M4(3, 7, 11, 15); break;
            case 5: SPZ_M4(5, 7, 13, 15); break;
            case 6: SPZ_M4(6, 7, 14, 15); break;
            case 9: SPZ_M4(9, 11, 13, 15); break;
            case 10: SPZ_M4(10, 11, 14, 15); break;
            case 12: SPZ_M4(12, 13, 14, 15); break;
            default:
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (((unsigned)k & m) == m) cmul3(ar[k], ai[k], f.x, f.y);
                break;
            }
#undef SPZ_M4
            continue;
        }
        // LAYOUT or END: apply what is pending, registers -> shared memory
        flush(ins.flags);
        store_regs();
        if (op == T3_END) break;
        __syncthreads();
        lay = ins.a;
        load_regs(); // no second barrier: this thread's next shared-memory access is store_regs() to the cells it has just read
    }

    // ---- tile out ----
    const unsigned long long base = tile_base();
#ifndef SPZ_CPU_EMULATION
    fence_proxy_async_smem(); // this thread's shared-memory writes become visible to the TMA engine
    __syncthreads();
    if (tid < 32) {
        for (unsigned s = tid; s < n_seg; s += 32) {
            const int row = seg_row(base, s);
            tma_store_box(&a.tm_re, row, smem + s * seg_bytes);
            tma_store_box(&a.tm_im, row, smem + kArrayBytes3 + s * seg_bytes);
        }
        tma_store_commit_and_wait_read(); // shared memory must stay alive until the engine has read it
    }
#else
    __syncthreads();
    if (tid == 0) {
        for (unsigned s = 0; s < n_seg; ++s) {
            const unsigned long long g0 = (unsigned long long)seg_row(base, s) << 4;
            for (unsigned j = 0; j < (1u << L); ++j) {
                const unsigned tj = (s << L) + j;
                a.re[g0 + j] = sre[swz3(tj)];
                a.im[g0 + j] = sim[swz3(tj)];
            }
        }
    }
#endif
}

// ---- host: lowering -----------------------------------------------------------------------------------------------------
inline void cmul_h(double &xr, double &xi, double fr, double fi) {
    const double nr = xr * fr - xi * fi, ni = xr * fi + xi * fr;
    xr = nr; xi = ni;
}

// Lower the scheduler's micro-program of one merged-mode pass.  Returns false when the program cannot run on k_tile3
// (exact-mode instructions, too many instructions / constants): the caller falls back to k_tile.
bool tile3_lower(const TilePlan &plan, const TileInstr *prog, int n_instr, const TileGroup *groups, int n_groups, const TileTerm *terms,
                 int n_terms, Lowered3 &out) {
    (void)n_groups; (void)n_terms; (void)plan;
    if (n_instr < 1 || prog[0].op != TI_LAYOUT) return false;
    out = Lowered3();
    int R[4] = {0, 1, 2, 3};
    unsigned dirty = 0;       // accumulators that are not the identity
    auto push = [&](const Ins3 &i, uint64_t outer) { out.ins.push_back(i); out.outer.push_back(outer); };
    auto pool2 = [&](double x, double y) -> uint32_t { // a double2 entry: even offset
        if (out.pool.size() & 1) out.pool.push_back(0.0);
        const uint32_t off = (uint32_t)out.pool.size();
        out.pool.push_back(x); out.pool.push_back(y);
        return off;
    };
    // tile-index mask -> thread-id mask under the current layout (register bits must not be set)
    auto to_tid = [&](uint32_t tile_mask, bool &ok) -> uint32_t {
        uint32_t r = 0;
        for (int b = 0; b < kT3; ++b) {
            if (!((tile_mask >> b) & 1u)) continue;
            int below = 0;
            for (int i = 0; i < 4; ++i) { if (R[i] == b) ok = false; if (R[i] < b) ++below; }
            r |= 1u << (b - below);
        }
        return r;
    };
    auto new_group = [&](const std::vector<TileTerm> &ts) -> uint32_t {
        TileGroup g{0, 0, (int)out.terms.size(), (int)ts.size()};
        out.terms.insert(out.terms.end(), ts.begin(), ts.end());
        out.groups.push_back(g);
        return (uint32_t)out.groups.size() - 1;
    };
    // which variant runs a non-diagonal gate, its scalars, and the real factor it leaves to the pass scale
    struct Variant { int kind = -1; double s[8]; int ns = 0; double factor = 1.0; };
    // `budget`: how much smaller the pass scale may still get (the amplitudes grow by the inverse until the scale is applied)
    auto variant_of = [](const TileInstr &t, double budget) -> Variant {
        Variant v;
        const bool controlled = t.reg_cmask || t.thr_cmask || t.outer_cmask;
        const double *s = t.s;
        auto set = [&](int kind, std::initializer_list<double> sc, double factor) {
            v.kind = kind; v.ns = 0; v.factor = factor;
            for (double x : sc) v.s[v.ns++] = x;
        };
        // The factored form is as accurate as the matrix for any cosine (same absolute rounding error once the scale is
        // applied); what bounds it is range: |cos| >= 2^-10 per gate, and the running product stays above 1e-100.
        constexpr double kMinScale = 1.0 / 1024.0;
        switch (t.kind) {
        case SPZ_GATE_H:
            if (!controlled && SPZ_SQRT_ONE_HALF >= budget) set(MK_H, {}, SPZ_SQRT_ONE_HALF);
            else set(MK_REAL, {SPZ_SQRT_ONE_HALF, SPZ_SQRT_ONE_HALF, SPZ_SQRT_ONE_HALF, -SPZ_SQRT_ONE_HALF}, 1.0);
            break;
        case SPZ_GATE_X: set(MK_X, {}, 1.0); break;
        case SPZ_GATE_Y: set(MK_Y, {}, 1.0); break;
        case SPZ_GATE_RX: // [[cs, i ns], [i ns, cs]], s = (cs, ns)   (gate_math.cuh)
            if (!controlled && std::fabs(s[0]) >= kMinScale && std::fabs(s[0]) >= budget) set(MK_RXT, {s[1] / s[0]}, s[0]);
            else set(MK_RXU, {s[0], s[1]}, 1.0);
            break;
        case SPZ_GATE_RY: // [[cs, -sn], [sn, cs]], s = (sn, cs)
            if (!controlled && std::fabs(s[1]) >= kMinScale && std::fabs(s[1]) >= budget) set(MK_RYT, {s[0] / s[1]}, s[1]);
            else set(MK_REAL, {s[1], -s[0], s[0], s[1]}, 1.0);
            break;
        case SPZ_GATE_U: // [[a, k + i l], [q + i r, ss + i t]], s = (a, k, l, q, r, ss, t)
            set(MK_U, {s[0], 0.0, s[1], s[2], s[3], s[4], s[5], s[6]}, 1.0);
            break;
        default: break;
        }
        return v;
    };
    std::vector<Variant> variants((size_t)n_instr);
    for (int pc = 0; pc < n_instr; ++pc) {
        if (prog[pc].op != TI_GATE) continue;
        variants[pc] = variant_of(prog[pc], 1e-100 / std::fabs(out.scale));
        out.scale *= variants[pc].factor;
    }
    // F0 starts as the pass scale in the kernel: when something was factored out it is pending from the first instruction on
    dirty = out.scale != 1.0 ? 1u : 0u;
    for (int pc = 0; pc < n_instr; ++pc) {
        const TileInstr &t = prog[pc];
        if (t.op == TI_LAYOUT) {
            Ins3 i{};
            i.op = T3_LAYOUT;
            i.flags = (uint8_t)dirty;
            for (int k = 0; k < 4; ++k) { R[k] = t.rbit[k]; if (R[k] < 0 || R[k] >= kT3 || (k && R[k] <= R[k - 1])) return false; }
            i.a = (uint32_t)R[0] | ((uint32_t)R[1] << 8) | ((uint32_t)R[2] << 16) | ((uint32_t)R[3] << 24);
            if (pc > 0) dirty = 0;
            push(i, 0);
            continue;
        }
        if (t.op == TI_GATE) {
            Ins3 i{};
            i.op = T3_GATE;
            i.rpos = (uint8_t)t.rpos;
            bool ok = true;
            i.thr = (uint16_t)to_tid(t.thr_cmask, ok);
            if (!ok || t.rpos < 0 || t.rpos > 3) return false;
            i.km = (uint16_t)t.t_mask;
            const bool in_tile_ctrl = t.reg_cmask || t.thr_cmask;
            if (in_tile_ctrl) out.ctrl = true; else i.flags |= GF_ALL;
            if (t.outer_cmask) i.flags |= GF_OUTER;
            if (dirty & (2u << t.rpos)) { i.flags |= GF_PRE; dirty &= ~(2u << t.rpos); }
            const Variant &v = variants[pc];
            if (v.kind < 0) return false;
            i.kind = (uint8_t)v.kind;
            if (v.ns) { i.a = (uint32_t)out.pool.size(); out.pool.insert(out.pool.end(), v.s, v.s + v.ns); }
            push(i, t.outer_cmask);
            continue;
        }
        if (t.op != TI_RUN) return false; // TI_DIAG: exact mode
        // ---- a merged run of diagonal gates: groups [rpos, rpos + sum of counts), classes m = 0, 1, 2, 4, 8, other ----
        const int counts[6] = {t.rbit[0], t.rbit[1], t.rbit[2], t.rbit[3], (int)t.reg_cmask, (int)t.thr_cmask};
        int g = t.rpos;
        for (int cls = 0; cls < 6; ++cls) {
            // gather this class's terms, masks converted to thread-id space
            struct Term { uint64_t outer; uint32_t thr, m; double fr, fi; };
            std::vector<Term> ts;
            for (int k = 0; k < counts[cls]; ++k, ++g) {
                const TileGroup &gd = groups[g];
                for (int j = 0; j < gd.count; ++j) {
                    const TileTerm &tm = terms[gd.first + j];
                    bool ok = true;
                    const uint32_t thr = to_tid(tm.thr, ok);
                    if (!ok) return false;
                    ts.push_back(Term{tm.outer, thr, tm.m, tm.fr, tm.fi});
                }
            }
            if (ts.empty()) continue;
            if (cls < 5) {
                double lo[16][2], hi[16][2];
                for (int e = 0; e < 16; ++e) { lo[e][0] = hi[e][0] = 1.0; lo[e][1] = hi[e][1] = 0.0; }
                bool use_lo = false, use_hi = false;
                std::vector<TileTerm> tile_terms;                                  // thr == 0, outer != 0: one constant per tile
                std::vector<std::pair<uint32_t, std::vector<TileTerm>>> general;   // everything else, by thread mask
                for (const Term &x : ts) {
                    if (x.outer == 0 && (x.thr & 0xF0u) == 0) {
                        for (unsigned e = 0; e < 16; ++e) if ((e & x.thr) == x.thr) cmul_h(lo[e][0], lo[e][1], x.fr, x.fi);
                        use_lo = true;
                    } else if (x.outer == 0 && (x.thr & 0x0Fu) == 0) {
                        for (unsigned e = 0; e < 16; ++e) if ((e & (x.thr >> 4)) == (x.thr >> 4)) cmul_h(hi[e][0], hi[e][1], x.fr, x.fi);
                        use_hi = true;
                    } else if (x.thr == 0) {
                        tile_terms.push_back(TileTerm{x.outer, 0, 0, x.fr, x.fi});
                    } else {
                        size_t k = 0;
                        while (k < general.size() && general[k].first != x.thr) ++k;
                        if (k == general.size()) general.emplace_back(x.thr, std::vector<TileTerm>());
                        general[k].second.push_back(TileTerm{x.outer, 0, 0, x.fr, x.fi});
                    }
                }
                if (use_lo || use_hi || !tile_terms.empty()) {
                    Ins3 i{};
                    i.op = T3_ACC;
                    i.rpos = (uint8_t)cls;
                    if (use_lo || use_hi) {
                        if (out.pool.size() & 1) out.pool.push_back(0.0);
                        i.a = (uint32_t)out.pool.size();
                        for (int e = 0; e < 16; ++e) { out.pool.push_back(lo[e][0]); out.pool.push_back(lo[e][1]); }
                        if (use_hi) for (int e = 0; e < 16; ++e) { out.pool.push_back(hi[e][0]); out.pool.push_back(hi[e][1]); }
                        if (use_lo) i.flags |= AF_LO;
                        if (use_hi) i.flags |= AF_HI;
                    }
                    if (!tile_terms.empty()) { i.flags |= AF_TILE; i.b = new_group(tile_terms); }
                    if (!(dirty & (1u << cls))) i.flags |= AF_SET;
                    dirty |= 1u << cls;
                    push(i, 0);
                }
                for (auto &gp : general) {
                    Ins3 i{};
                    i.op = T3_ACCG;
                    i.rpos = (uint8_t)cls;
                    i.thr = (uint16_t)gp.first;
                    bool any_outer = false;
                    for (const TileTerm &x : gp.second) any_outer |= x.outer != 0;
                    if (any_outer) { i.flags |= AF_TILE; i.b = new_group(gp.second); }
                    else {
                        double fr = 1.0, fi = 0.0;
                        for (const TileTerm &x : gp.second) cmul_h(fr, fi, x.fr, x.fi);
                        i.flags |= AF_CONST; i.a = pool2(fr, fi);
                    }
                    if (!(dirty & (1u << cls))) i.flags |= AF_SET;
                    dirty |= 1u << cls;
                    push(i, 0);
                }
            } else {
                // two or more register bits: by (thread mask, register mask)
                std::vector<std::pair<std::pair<uint32_t, uint32_t>, std::vector<TileTerm>>> byk;
                for (const Term &x : ts) {
                    size_t k = 0;
                    while (k < byk.size() && !(byk[k].first.first == x.thr && byk[k].first.second == x.m)) ++k;
                    if (k == byk.size()) byk.emplace_back(std::make_pair(x.thr, x.m), std::vector<TileTerm>());
                    byk[k].second.push_back(TileTerm{x.outer, 0, 0, x.fr, x.fi});
                }
                for (auto &gp : byk) {
                    Ins3 i{};
                    i.op = T3_OTHER;
                    i.thr = (uint16_t)gp.first.first;
                    i.km = (uint16_t)gp.first.second;
                    bool any_outer = false;
                    for (const TileTerm &x : gp.second) any_outer |= x.outer != 0;
                    if (any_outer) { i.flags |= AF_TILE; i.b = new_group(gp.second); }
                    else {
                        double fr = 1.0, fi = 0.0;
                        for (const TileTerm &x : gp.second) cmul_h(fr, fi, x.fr, x.fi);
                        i.flags |= AF_CONST; i.a = pool2(fr, fi);
                    }
                    push(i, 0);
                }
            }
        }
    }
    Ins3 e{};
    e.op = T3_END;
    e.flags = (uint8_t)dirty;
    push(e, 0);
    if (out.pool.size() & 1) out.pool.push_back(0.0);
    return out.ins.size() <= (size_t)kMaxIns3 && out.groups.size() <= (size_t)kMaxGroups3;
}

// Serialise for upload (layout: see Lowered3) and fill the size fields of the kernel arguments.
size_t tile3_pack(const Lowered3 &lw, std::vector<unsigned char> &blob, Tile3Args &a) {
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t ins_bytes = up16(lw.ins.size() * sizeof(Ins3));
    const size_t pool_bytes = up16(lw.pool.size() * sizeof(double));
    const size_t outer_off = ins_bytes + pool_bytes;
    const size_t groups_off = up16(outer_off + lw.outer.size() * sizeof(uint64_t));
    const size_t terms_off = up16(groups_off + lw.groups.size() * sizeof(TileGroup));
    const size_t total = up16(terms_off + lw.terms.size() * sizeof(TileTerm));
    blob.assign(total, 0);
    std::memcpy(blob.data(), lw.ins.data(), lw.ins.size() * sizeof(Ins3));
    if (!lw.pool.empty()) std::memcpy(blob.data() + ins_bytes, lw.pool.data(), lw.pool.size() * sizeof(double));
    std::memcpy(blob.data() + outer_off, lw.outer.data(), lw.outer.size() * sizeof(uint64_t));
    if (!lw.groups.empty()) std::memcpy(blob.data() + groups_off, lw.groups.data(), lw.groups.size() * sizeof(TileGroup));
    if (!lw.terms.empty()) std::memcpy(blob.data() + terms_off, lw.terms.data(), lw.terms.size() * sizeof(TileTerm));
    a.ins_bytes = (unsigned)ins_bytes; a.pool_bytes = (unsigned)pool_bytes;
    a.outer_off = (unsigned)outer_off; a.groups_off = (unsigned)groups_off; a.terms_off = (unsigned)terms_off;
    a.n_ins = (int)lw.ins.size(); a.n_groups = (int)lw.groups.size();
    a.scale = lw.scale;
    return total;
}

size_t tile3_smem_bytes(const Tile3Args &a) {
    return ((kProgOff3 + a.ins_bytes + a.pool_bytes + 16u * (size_t)a.n_groups + (size_t)a.n_ins + 15u) & ~(size_t)15u) + 16u;
}

bool tile3_shape_ok(int n_qubits, const TilePlan &plan) {
    return plan.tile_bits == kT3 && plan.n_high <= kMaxHigh3 && plan.low_bits >= 4 && n_qubits >= kT3 &&
           plan.tile_bits == plan.low_bits + plan.n_high;
}

} // namespace

#ifndef SPZ_CPU_EMULATION
// SPZ_TILE_V3=0 sends merged passes to k_tile as well (A/B runs); default on.
bool tile3_enabled() {
    const char *e = std::getenv("SPZ_TILE_V3");
    return !(e && e[0] == '0');
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// 2-D view of one state array: rows of 16 doubles (128 bytes, the widest inner box the 128-byte swizzle allows); a tile
// segment of 2^L amplitudes is a box of 2^(L-4) rows.
static int make_tensor_map(CUtensorMap *tm, double *base, int64_t len, int L) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SPZ_ERR_CUDA; }
    const cuuint64_t gdim[2] = {16, (cuuint64_t)(len >> 4)};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {16, 1u << (L - 4)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for len=%lld L=%d", (int)r, (long long)len, L); return SPZ_ERR_CUDA; }
    return SPZ_OK;
}

int tile3_prepare() {
    SPZ_CUDA(cudaFuncSetAttribute(k_tile3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget3));
    return SPZ_OK;
}

// One pass on k_tile3: lower, upload through the state's program ring buffer, encode the two tensor maps.  *handled = false
// (nothing staged) when the pass has to go to k_tile instead.  run_tile3 launches tiles [first, first + count).
int prepare_tile3(spz_state *st, const TilePlan &plan, const TileInstr *prog, int n_instr, const TileGroup *groups, int n_groups,
                  const TileTerm *terms, int n_terms, Tile3Launch *out, bool *handled) {
    *handled = false;
    if (!tile3_shape_ok(st->n, plan)) return SPZ_OK;
    Lowered3 lw;
    if (!tile3_lower(plan, prog, n_instr, groups, n_groups, terms, n_terms, lw)) return SPZ_OK;
    Tile3Args a{};
    std::vector<unsigned char> blob;
    const size_t bytes = tile3_pack(lw, blob, a);
    const size_t smem = tile3_smem_bytes(a);
    if (smem > kSmemBudget3) return SPZ_OK;
    static bool prepared[64] = {false};
    if (st->device >= 0 && st->device < 64 && !prepared[st->device]) {
        SPZ_TRY(tile3_prepare());
        prepared[st->device] = true;
    }
    SPZ_TRY(make_tensor_map(&a.tm_re, st->re, st->len, plan.low_bits));
    SPZ_TRY(make_tensor_map(&a.tm_im, st->im, st->len, plan.low_bits));
    char *slot = nullptr;
    SPZ_TRY(tile_ring_alloc(st, bytes, &slot));
    SPZ_CUDA(cudaMemcpyAsync(slot, blob.data(), bytes, cudaMemcpyHostToDevice, st->stream));
    a.re = st->re; a.im = st->im;
    a.blob = reinterpret_cast<const unsigned char *>(slot);
    a.L = plan.low_bits; a.n_high = plan.n_high;
    for (int k = 0; k < plan.n_high; ++k) a.high[k] = plan.high[k];
    static_assert(sizeof(Tile3Args) <= sizeof(out->args), "Tile3Launch::args too small");
    std::memcpy(out->args, &a, sizeof a);
    out->smem = smem;
    *handled = true;
    return SPZ_OK;
}

void run_tile3(spz_state *st, const Tile3Launch &l, unsigned first, unsigned count) {
    Tile3Args a;
    std::memcpy(&a, l.args, sizeof a);
    a.tile_offset = first;
    k_tile3<<<count, kThreads3, l.smem, st->stream>>>(a);
}
#endif // !SPZ_CPU_EMULATION

} // namespace spz
