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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
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
// op byte = the interpreter's arm: one flat switch (a jump table) instead of a chain of tests.  T3_PRE + r: register bit r.
// T3_GATE + 4 * variant + r: variant 0..2 = H, RX, RY on every pair (no in-tile control), 3..7 = HS, RX, RY, X, Y under a pair mask.
enum { T3_END = 0, T3_LAYOUT = 1, T3_ACC = 2, T3_ACCG = 3, T3_OTHER = 4, T3_PRE = 5, T3_GATE = 9, T3_N_ARMS = 41 };
// gate variants of merged mode (see the header and pair3): H = [[1,1],[1,-1]] with 2^-1/2 left to the pass scale, HS the same
// with the scale applied (controlled H), RX / RY = rotation by three shears, X / Y = exchanges
enum { MK_H = 0, MK_HS = 1, MK_RX = 2, MK_RY = 3, MK_X = 4, MK_Y = 5 };
// flag bits
enum {
    GF_ALL = 1,      // GATE: every pair of every thread (no in-tile control)
    GF_OUTER = 32,   // GATE: has controls outside the tile: consult the per-tile skip flag
    LF_LAST = 128,   // LAYOUT: the last one of the program: shared memory is free once its registers are loaded
    AF_LO = 1,       // ACC: table over the low nibble of the thread id at pool2[a .. a+16)
    AF_HI = 2,       // ACC: table over the high nibble at pool2[a+16 .. a+32)
    AF_TILE = 4,     // ACC / ACCG / OTHER: per-tile constant gfac[b]
    AF_SET = 8,      // ACC / ACCG: the accumulator holds no pending factor: assign instead of multiply
    AF_CONST = 16,   // ACCG / OTHER: constant factor at pool2[a]
};
struct Ins3 {
    uint8_t op, kind, rpos, flags; // op: see T3_*.  rpos: ACC / ACCG accumulator 0..4 (GATE / PRE: register bit, also in op).
                                   // LAYOUT / END flags: pending accumulators
    uint16_t km;                   // GATE: pair mask over k0.  OTHER: register mask m
    uint16_t thr;                  // GATE / ACCG / OTHER: control bits in THREAD-ID space (8 bits)
    uint32_t a;                    // LAYOUT: the 4 register-resident tile bits, one byte each.  ACC: pool offset of its tables
    uint32_t b;                    // per-tile constant index (LAYOUT / END: pool offset of the pending-constant table)
    double s[2];                   // the instruction's own scalars: a rotation's (tan, sin), the constant of a PRE / ACCG / OTHER
                                   // with AF_CONST -- in the instruction, so that an arm's operand is ONE load at a known address
                                   // (through the pool it was instruction word -> address -> scalars: two dependent loads)
};
static_assert(sizeof(Ins3) == 32, "Ins3 is decoded with 128-bit shared-memory loads: operand words at +0, scalars at +16");
constexpr unsigned kIns3 = 32;

// What the host uploads for one pass (one contiguous blob, 16-byte aligned sections):
//   Ins3 ins[n_ins] | double pool[n_pool] (gate scalars, tables, constants; double2 entries at even offsets)
//   | uint64 outer[n_ins] (controls outside the tile, GATE only) | TileGroup groups[n_groups] | TileTerm terms[n_terms]
// Shared memory of a CTA: tile re | tile im | accumulators F[5][256] | the blob | gfac[n_groups] | skip[n_ins] | 2 mbarriers
struct Lowered3 {
    std::vector<Ins3> ins;
    std::vector<double> pool;
    std::vector<uint64_t> outer;
    std::vector<TileGroup> groups; // thr, m unused; first / count index `terms`
    std::vector<TileTerm> terms;   // outer, fr, fi used
    double scale = 1.0;            // product of the factored-out gate scalars: initial value of F0
    bool ctrl = false;             // some butterfly has an in-tile control
    bool single_layout = true;     // no LAYOUT after the first
};

struct Tile3Args {
#ifndef SPZ_CPU_EMULATION
    CUtensorMap tm_re, tm_im;      // see make_tensor_maps: 128-byte rows x rows of a segment x up to three runs of high tile qubits
#endif
    // How the tile maps to TMA boxes: a box covers the low L bits and the first `rank - 2` runs of contiguous high tile qubits
    // (2^box_shift amplitudes, contiguous in tile-index order); the n_rest high tile qubits above them are enumerated, one box
    // per combination.  dim_lo[i] / dim_len[i]: the index bits tensor dimension i spans (its coordinate is that bit field of
    // the box's first amplitude).
    int rank, n_rest, box_shift;
    int rest_bit[kMaxHigh3];
    int dim_lo[5], dim_len[5];
    double *re, *im;               // (the CPU emulation moves the tile through these)
    const unsigned char *blob;     // device copy of the lowered program
    unsigned ins_bytes, blob_bytes;  // instruction section and whole blob (multiples of 16)
    unsigned outer_off, groups_off, terms_off; // byte offsets of the other sections inside the blob (the pool follows the instructions)
    int n_ins, n_groups, n_terms;
    unsigned tile_first, tile_end; // this launch's tiles (a pass may be launched in two halves, see dist.cu)
    int single_layout;             // the program never changes the register layout: the next tile can be fetched at once
    int L, n_high;
    double scale;
    int high[kMaxHigh3];
    unsigned long long *prof; // SPZ_TILE_PROF=1 (diagnostic): per-phase nanoseconds summed over CTAs, see prepare_tile3
};

// position of tile index j inside a 32 KB array under the TMA 128-byte swizzle (16-byte chunk index ^= 128-byte row
// index mod 8); linear over GF(2), so swz3(a ^ b) == swz3(a) ^ swz3(b)
__host__ __device__ __forceinline__ unsigned swz3(unsigned j) { return j ^ (((j >> 4) & 7u) << 1); }

__device__ __forceinline__ void cmul3(double &xr, double &xi, double fr, double fi) {
    // two products into temporaries, then each result lands in its own operand's register by one FMA (see pair3): no copy of
    // the old real part is needed (t = xr; ... cost two MOVs per multiply, a sixth of the kernel's integer instructions)
    const double p = xi * fi, q = xr * fi;
    xr = fma(xr, fr, -p);
    xi = fma(xi, fr, q);
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
// 1-D bulk copy global -> shared (addresses and size multiples of 16 bytes), completing on an mbarrier
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// One box of a rank-2..5 tensor map (coordinate 0 is always 0: the 16 doubles of a 128-byte row).
__device__ __forceinline__ void tma_load_box(int rank, void *dst, const CUtensorMap *tm, const int (&c)[5], void *bar) {
    const unsigned d = smem_u32(dst), b = smem_u32(bar);
    const uint64_t m = reinterpret_cast<uint64_t>(tm);
    switch (rank) {
    case 2:
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(d), "l"(m), "r"(0), "r"(c[1]), "r"(b) : "memory");
        break;
    case 3:
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(d), "l"(m), "r"(0), "r"(c[1]), "r"(c[2]), "r"(b) : "memory");
        break;
    case 4:
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(d), "l"(m), "r"(0), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(b) : "memory");
        break;
    default:
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(d), "l"(m), "r"(0), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(b) : "memory");
        break;
    }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// ---- gate variants ------------------------------------------------------------------------------------------------------
// (a, b) = (re, im) of the amplitude with target bit 0, (c, d) with target bit 1.
//
// EVERY update below is strictly in place: each new value overwrites one of its own operands at that operand's last use.
// This is not style.  The 32 amplitude values live in registers across the interpreter loop; when an arm computes a new value
// while the old one is still needed (na = a + c; nc = a - c), ptxas gives the new values a second bank of 64 registers and
// copies bank to bank around EVERY interpreted instruction (~130 MOVs each, measured: 12 G warp instructions for a 27-gate
// pass, of which 1.2 G were FP64), and spills what no longer fits.  Written in place, the same loop compiles to no copies.
//   H        c <- a - c; a <- 2a - c                 (2 ops per component, like the textbook form)
//   RX / RY  a rotation of two real couples by three shears: x <- x - t y; y <- y + s x; x <- x - t y, t = tan(phi/2),
//            s = sin(phi), |phi| <= pi/2 (the host reduces the angle; the sign of cos goes to the pass scale or to a phase term)
//   X / Y    exchanges by three XORs inside an asm block: written as a swap, the compiler sees a renaming and we are back to
//            bank copies
__device__ __forceinline__ void xor_swap(double &x, double &y) {
#ifdef SPZ_CPU_EMULATION
    const double t = x; x = y; y = t;
#else
    asm volatile("xor.b64 %0, %0, %1;\n\txor.b64 %1, %1, %0;\n\txor.b64 %0, %0, %1;" : "+d"(x), "+d"(y));
#endif
}
// Shared-memory loads by 32-bit shared-window address.  Through generic pointers the compiler re-derives the window base
// (S2R SR_CgaCtaId, LEA) at every use inside the interpreter loop; the emulation's "address" is an offset into its window.
#ifdef SPZ_CPU_EMULATION
__device__ __forceinline__ unsigned sm_addr(const void *p) { return (unsigned)(static_cast<const unsigned char *>(p) - spz_emu::dyn_smem); }
__device__ __forceinline__ uint4 lds128(unsigned a) { return *reinterpret_cast<const uint4 *>(spz_emu::dyn_smem + a); }
__device__ __forceinline__ double2 lds_d2(unsigned a) { return *reinterpret_cast<const double2 *>(spz_emu::dyn_smem + a); }
__device__ __forceinline__ double lds_d(unsigned a) { return *reinterpret_cast<const double *>(spz_emu::dyn_smem + a); }
__device__ __forceinline__ unsigned lds32(unsigned a) { return *reinterpret_cast<const unsigned *>(spz_emu::dyn_smem + a); }
__device__ __forceinline__ void sts_d2(unsigned a, double x, double y) { *reinterpret_cast<double2 *>(spz_emu::dyn_smem + a) = make_double2(x, y); }
__device__ __forceinline__ void sts_d(unsigned a, double x) { *reinterpret_cast<double *>(spz_emu::dyn_smem + a) = x; }
__device__ __forceinline__ unsigned lds8(unsigned a) { return spz_emu::dyn_smem[a]; }
#else
// (volatile: the conversion reads SR_CgaCtaId, a slow special register; left to itself the compiler re-derives the address
// inside the interpreter loop instead of keeping it)
__device__ __forceinline__ unsigned sm_addr(const void *p) {
    unsigned r;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 lds128(unsigned a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_d2(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_d(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds32(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_d2(unsigned a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts_d(unsigned a, double x) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }
__device__ __forceinline__ unsigned lds8(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
#endif
// 256-bit global store (one full 32-byte sector per lane); plain stores in the CPU emulation build
__device__ __forceinline__ void st_global4(double *p, double x0, double x1, double x2, double x3) {
#ifdef SPZ_CPU_EMULATION
    p[0] = x0; p[1] = x1; p[2] = x2; p[3] = x3;
#else
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
#endif
}
// The value, but opaque to the optimiser: expressions derived from it are computed where they are used.  Without it the
// compiler hoists every arm's address arithmetic (accumulator c of this thread, ...) to the head of the interpreter loop and
// recomputes all of it for every interpreted instruction.
__device__ __forceinline__ unsigned opaque(unsigned v) {
#ifndef SPZ_CPU_EMULATION
    asm volatile("" : "+r"(v));
#endif
    return v;
}
__device__ __forceinline__ void shear3(double &x, double &y, double t, double sn) {
    x = x - t * y;
    y = y + sn * x;
    x = x - t * y;
}
template <int MK>
__device__ __forceinline__ void pair3(const double (&s)[2], double &a, double &b, double &c, double &d) {
    if constexpr (MK == MK_H) {            // [[1, 1], [1, -1]]
        c = a - c; a = fma(2.0, a, -c);
        d = b - d; b = fma(2.0, b, -d);
    } else if constexpr (MK == MK_HS) {    // 2^-1/2 [[1, 1], [1, -1]]
        c = a - c; a = fma(2.0, a, -c); c = c * SPZ_SQRT_ONE_HALF; a = a * SPZ_SQRT_ONE_HALF;
        d = b - d; b = fma(2.0, b, -d); d = d * SPZ_SQRT_ONE_HALF; b = b * SPZ_SQRT_ONE_HALF;
    } else if constexpr (MK == MK_RX) {    // [[cos, i sin], [i sin, cos]]: rotates the couples (a, d) and (c, b)
        shear3(a, d, s[0], s[1]);
        shear3(c, b, s[0], s[1]);
    } else if constexpr (MK == MK_RY) {    // [[cos, -sin], [sin, cos]]: rotates the couples (a, c) and (b, d)
        shear3(a, c, s[0], s[1]);
        shear3(b, d, s[0], s[1]);
    } else if constexpr (MK == MK_X) {
        xor_swap(a, c);
        xor_swap(b, d);
    } else {                               // Y = [[0, -i], [i, 0]]: s0 <- (d, -c), s1 <- (-b, a)
        xor_swap(a, d);
        xor_swap(b, c);
        b = -b; c = -c;
    }
}
template <int MK>
constexpr int n_scalars3() { return MK == MK_RX || MK == MK_RY ? 2 : 0; }

template <int MK, int R, bool ALL>
__device__ __forceinline__ void bfly3(double (&ar)[16], double (&ai)[16], unsigned sp, unsigned km) {
    double s[2] = {0.0, 0.0};
    if (n_scalars3<MK>() == 2) { const double2 v = lds_d2(sp); s[0] = v.x; s[1] = v.y; } // (rotation scalars sit at even pool offsets)
#pragma unroll
    for (int k0 = 0; k0 < 16; ++k0) {
        if (k0 & (1 << R)) continue;
        const int k1 = k0 | (1 << R);
        if (ALL || (km & (1u << k0))) pair3<MK>(s, ar[k0], ai[k0], ar[k1], ai[k1]); // km is the same for every thread
    }
}
// X / Y under a pair mask.  The mask of a guarded exchange is one of a few shapes: every pair (the controls are thread bits or
// bits outside the tile), or the pairs whose register index has one other register bit set (a CX between two register-resident
// qubits: four pairs).  Testing the mask once and running a branch-free block costs 48-96 logic instructions; testing it
// pair by pair cost a compare, a branch and a reconvergence point per pair on top (~90-130 per exchange, and a CX is a fifth
// of a layered circuit).
template <int MK, int R>
__device__ __forceinline__ void perm3(double (&ar)[16], double (&ai)[16], unsigned km) {
    const double s[2] = {0.0, 0.0};
    auto with_bit = [&](auto c_tag) {
        constexpr int C = decltype(c_tag)::value;
#pragma unroll
        for (int k0 = 0; k0 < 16; ++k0)
            if (!(k0 & (1 << R)) && (k0 & (1 << C))) pair3<MK>(s, ar[k0], ai[k0], ar[k0 | (1 << R)], ai[k0 | (1 << R)]);
    };
    constexpr unsigned m0 = 0xAAAAu, m1 = 0xCCCCu, m2 = 0xF0F0u, m3 = 0xFF00u; // register indices with bit 0 / 1 / 2 / 3 set
    if (km == 0xffffu) {
#pragma unroll
        for (int k0 = 0; k0 < 16; ++k0)
            if (!(k0 & (1 << R))) pair3<MK>(s, ar[k0], ai[k0], ar[k0 | (1 << R)], ai[k0 | (1 << R)]);
    } else if (R != 0 && km == m0) { with_bit(std::integral_constant<int, R != 0 ? 0 : 1>());
    } else if (R != 1 && km == m1) { with_bit(std::integral_constant<int, R != 1 ? 1 : 0>());
    } else if (R != 2 && km == m2) { with_bit(std::integral_constant<int, R != 2 ? 2 : 0>());
    } else if (R != 3 && km == m3) { with_bit(std::integral_constant<int, R != 3 ? 3 : 0>());
    } else {
        bfly3<MK, R, false>(ar, ai, 0u, km);
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
    // the whole program blob is staged behind the accumulators (one bulk copy): ins | pool | outer | groups | terms
    const Ins3 *sins = reinterpret_cast<const Ins3 *>(smem + kProgOff3);
    const double *pool = reinterpret_cast<const double *>(smem + kProgOff3 + a.ins_bytes);
    double2 *gfac = reinterpret_cast<double2 *>(smem + kProgOff3 + a.blob_bytes);
    // F[c][tid]: accumulator c of this thread.  They live in shared memory (one 128-bit access each, conflict-free) rather
    // than in 20 registers: with 64 registers of amplitudes the budget of 128 is tight.
    unsigned char *skip = reinterpret_cast<unsigned char *>(gfac + a.n_groups);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(
        smem + ((kProgOff3 + a.blob_bytes + 16u * (unsigned)a.n_groups + (unsigned)a.n_ins + 15u) & ~15u)); // [0] tile, [1] program
    // hi_off[s]: offset of tile segment s (the high tile bits of a tile index, deposited at their qubits); built once per CTA
    unsigned long long *hi_off = bar + 2;
    const unsigned tid = threadIdx.x;
    const int L = a.L;
#ifndef SPZ_CPU_EMULATION
    unsigned long long t_prev = 0;
    auto stamp = [&](int k) { // thread 0: time since the previous stamp into prof[k]
        if (a.prof && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (k >= 0) atomicAdd(a.prof + k, t - t_prev);
            t_prev = t;
        }
    };
    stamp(-1);
#else
    auto stamp = [](int) {};
#endif

    // absolute index of tile t's first amplitude: the tile number fills the non-tile bit positions
    auto tile_base = [&](unsigned t) -> unsigned long long {
        unsigned long long b = (unsigned long long)t << L;
#pragma unroll
        for (int k = 0; k < kMaxHigh3; ++k)
            if (k < a.n_high) b = insert_zero(b, a.high[k]);
        return b;
    };
    // global offset (inside the tile's footprint) of tile index x: the low L bits stay, bit L + k goes to qubit high[k]
    auto tile_to_global = [&](unsigned x) -> unsigned long long { return hi_off[x >> L] + (x & ((1u << L) - 1u)); };

    // ---- tile in: TMA boxes (one per array for tiles with up to three runs of high qubits) on mbarrier 0.  Called by warp 0. ----
#ifndef SPZ_CPU_EMULATION
    auto issue_tile_load = [&](unsigned t) {
        const unsigned long long base = tile_base(t);
        const unsigned n_box = 1u << a.n_rest;
        const unsigned box_bytes = 8u << a.box_shift;
        if (tid == 0) mbar_expect_tx(bar, 2u * kArrayBytes3);
        __syncwarp();
        for (unsigned e = tid; e < n_box; e += 32) {
            unsigned long long addr = base;
#pragma unroll
            for (int k = 0; k < kMaxHigh3; ++k)
                if (k < a.n_rest && ((e >> k) & 1u)) addr |= 1ull << a.rest_bit[k];
            int c[5];
            c[0] = 0;
#pragma unroll
            for (int i = 1; i < 5; ++i) c[i] = i < a.rank ? (int)((addr >> a.dim_lo[i]) & ((1ull << a.dim_len[i]) - 1ull)) : 0;
            tma_load_box(a.rank, smem + e * box_bytes, &a.tm_re, c, bar);
            tma_load_box(a.rank, smem + kArrayBytes3 + e * box_bytes, &a.tm_im, c, bar);
        }
    };
#endif
    const unsigned first_tile = a.tile_first + blockIdx.x;
    if (first_tile >= a.tile_end) return;
    const unsigned sbase = sm_addr(smem);
    const unsigned ins_addr = sbase + kProgOff3, pool_addr = ins_addr + a.ins_bytes, gfac_addr = ins_addr + a.blob_bytes;
    const unsigned skip_addr = gfac_addr + 16u * (unsigned)a.n_groups;
    // this thread's accumulator 0; accumulator c is 16 * kThreads3 * c further.  (facc_base + 16 * opaque(tid) inside the interpreter:
    // two instructions in the arm that needs them instead of three at the head of the loop for every instruction.)
    const unsigned facc_base = sbase + 2u * kArrayBytes3, facc_addr = facc_base + 16u * tid;
    for (unsigned sg = tid; sg < (1u << a.n_high); sg += kThreads3) {
        unsigned long long o = 0;
#pragma unroll
        for (int k = 0; k < kMaxHigh3; ++k)
            if (k < a.n_high && ((sg >> k) & 1u)) o |= 1ull << a.high[k];
        hi_off[sg] = o;
    }

    // ---- once per CTA: barriers, the program, the first tile ----
#ifndef SPZ_CPU_EMULATION
    if ((smem_u32(smem) & 1023u) != 0u) __trap(); // the swizzle pattern is anchored to 1 KB-aligned shared addresses
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_expect_tx(bar + 1, a.blob_bytes);
        bulk_load(smem + kProgOff3, a.blob, a.blob_bytes, bar + 1);
    }
    __syncthreads();
    if (tid < 32) issue_tile_load(first_tile);
    mbar_wait(bar + 1, 0);
#else
    for (unsigned i = tid; i < (a.blob_bytes >> 4); i += kThreads3)
        reinterpret_cast<uint4 *>(smem + kProgOff3)[i] = reinterpret_cast<const uint4 *>(a.blob)[i];
    __syncthreads();
#endif
    stamp(0); // launch -> program staged

    // ---- registers ----
    double ar[16], ai[16];
    // The register layout is carried as one word (4 tile bits, one byte each) and expanded where it is used: the swizzled
    // position of amplitude k of this thread is stj ^ (sw0 if k & 1) ^ (sw1 if k & 2) ^ ...  (swz3 is linear over GF(2)).
    unsigned lay = sins[0].a;
    auto thread_index = [&](int r0, int r1, int r2, int r3) -> unsigned { // this thread's tile index with the register bits clear
        unsigned x = tid;
        x = ((x >> r0) << (r0 + 1)) | (x & ((1u << r0) - 1u));
        x = ((x >> r1) << (r1 + 1)) | (x & ((1u << r1) - 1u));
        x = ((x >> r2) << (r2 + 1)) | (x & ((1u << r2) - 1u));
        x = ((x >> r3) << (r3 + 1)) | (x & ((1u << r3) - 1u));
        return x;
    };
    // (byte offsets inside the re array; the im array is kArrayBytes3 further; shared addresses, not generic pointers)
    auto move_regs = [&](auto &&xfer2, auto &&xfer1) {
        const int r0 = lay & 255u, r1 = (lay >> 8) & 255u, r2 = (lay >> 16) & 255u, r3 = lay >> 24;
        const unsigned stj = 8u * swz3(thread_index(r0, r1, r2, r3));
        const unsigned sw0 = 8u * swz3(1u << r0), sw1 = 8u * swz3(1u << r1), sw2 = 8u * swz3(1u << r2), sw3 = 8u * swz3(1u << r3);
        if (r0 == 0) { // register bit 0 is tile bit 0: amplitudes k, k + 1 are adjacent in shared memory (128-bit accesses)
#pragma unroll
            for (int k = 0; k < 16; k += 2) xfer2(k, sbase + (stj ^ ((k & 2) ? sw1 : 0u) ^ ((k & 4) ? sw2 : 0u) ^ ((k & 8) ? sw3 : 0u)));
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) xfer1(k, sbase + (stj ^ ((k & 1) ? sw0 : 0u) ^ ((k & 2) ? sw1 : 0u) ^ ((k & 4) ? sw2 : 0u) ^ ((k & 8) ? sw3 : 0u)));
        }
    };
    auto load_regs = [&]() {
        move_regs(
            [&](int k, unsigned s) {
                const double2 r = lds_d2(s), m = lds_d2(s + kArrayBytes3);
                ar[k] = r.x; ar[k + 1] = r.y; ai[k] = m.x; ai[k + 1] = m.y;
            },
            [&](int k, unsigned s) { ar[k] = lds_d(s); ai[k] = lds_d(s + kArrayBytes3); });
    };
    auto store_regs = [&]() {
        move_regs(
            [&](int k, unsigned s) {
                sts_d2(s, ar[k], ar[k + 1]);
                sts_d2(s + kArrayBytes3, ai[k], ai[k + 1]);
            },
            [&](int k, unsigned s) { sts_d(s, ar[k]); sts_d(s + kArrayBytes3, ai[k]); });
    };
    // Tile out: straight from the registers under the final layout.  The tile's shared-memory buffer is not involved, which is
    // what lets the NEXT tile's boxes land in it while this tile is still being computed (see prefetch below).  Lanes run over
    // the lowest non-register tile bits, so a warp writes runs of consecutive amplitudes; when register bits 0 (and 1) are
    // tile bits 0 (and 1) a thread's own amplitudes are adjacent and go out as 128-bit (256-bit) stores.
    auto direct_store = [&](unsigned long long base) {
        const int r0 = lay & 255u, r1 = (lay >> 8) & 255u, r2 = (lay >> 16) & 255u, r3 = lay >> 24;
        const unsigned long long g0 = base + tile_to_global(thread_index(r0, r1, r2, r3));
        const unsigned long long o0 = tile_to_global(1u << r0), o1 = tile_to_global(1u << r1);
        const unsigned long long o2 = tile_to_global(1u << r2), o3 = tile_to_global(1u << r3);
        if (r0 == 0 && r1 == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned long long g = g0 + ((q & 1) ? o2 : 0ull) + ((q & 2) ? o3 : 0ull);
                st_global4(a.re + g, ar[4 * q], ar[4 * q + 1], ar[4 * q + 2], ar[4 * q + 3]);
                st_global4(a.im + g, ai[4 * q], ai[4 * q + 1], ai[4 * q + 2], ai[4 * q + 3]);
            }
        } else if (r0 == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned long long g = g0 + ((q & 1) ? o1 : 0ull) + ((q & 2) ? o2 : 0ull) + ((q & 4) ? o3 : 0ull);
                *reinterpret_cast<double2 *>(a.re + g) = make_double2(ar[2 * q], ar[2 * q + 1]);
                *reinterpret_cast<double2 *>(a.im + g) = make_double2(ai[2 * q], ai[2 * q + 1]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned long long g = g0 + ((k & 1) ? o0 : 0ull) + ((k & 2) ? o1 : 0ull) + ((k & 4) ? o2 : 0ull) + ((k & 8) ? o3 : 0ull);
                a.re[g] = ar[k];
                a.im[g] = ai[k];
            }
        }
    };
    // apply the pending accumulators named by `mask` (the host knows which are pending: it is a property of the program,
    // so nothing is ever reset: an accumulator that has been applied is assigned, not multiplied, by its next update)
    // `cmask`: the classes whose pending factor is a constant the host folded (entry c of the five-entry table at shared address
    // `tab`) instead of the thread's accumulator
    auto flush = [&](unsigned mask, unsigned cmask, unsigned tab) {
        if (!mask) return;
        const unsigned fa = (facc_base + 16u * opaque(tid));
        auto pending = [&](unsigned c) -> double2 { return lds_d2((cmask >> c) & 1u ? tab + 16u * c : fa + c * 16u * kThreads3); };
        if (__popc(mask) <= 2) {
            if (mask & 1u) {
                const double2 f = pending(0);
#pragma unroll
                for (int k = 0; k < 16; ++k) cmul3(ar[k], ai[k], f.x, f.y);
            }
            if (mask & 2u) apply_bit3<0>(ar, ai, pending(1));
            if (mask & 4u) apply_bit3<1>(ar, ai, pending(2));
            if (mask & 8u) apply_bit3<2>(ar, ai, pending(3));
            if (mask & 16u) apply_bit3<3>(ar, ai, pending(4));
            return;
        }
        // three or more: expand the 16 per-amplitude factors (15 complex multiplies) and apply them once
        const double2 one = make_double2(1.0, 0.0);
        const double2 f0 = (mask & 1u) ? pending(0) : one, f1 = (mask & 2u) ? pending(1) : one;
        const double2 f2 = (mask & 4u) ? pending(2) : one, f3 = (mask & 8u) ? pending(3) : one;
        const double2 f4 = (mask & 16u) ? pending(4) : one;
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
        const unsigned f = (facc_base + 16u * opaque(tid)) + cls * 16u * kThreads3;
        if (!set) { const double2 old = lds_d2(f); cmul3(fr, fi, old.x, old.y); }
        sts_d2(f, fr, fi);
    };

    // ---- the CTA's tiles (persistent: 2 CTAs per SM walk the pass) ----
    unsigned parity = 0;
    for (unsigned t = first_tile; t < a.tile_end; t += gridDim.x) {
        const unsigned long long base = tile_base(t);
        // The next tile is fetched as soon as this one has been read out of shared memory for the last time: after the load
        // under the final register layout.  Its HBM latency then hides behind the rest of this tile's program and its stores.
        auto prefetch = [&]() {
            __syncthreads(); // every thread has read its registers: the buffer is free
#ifndef SPZ_CPU_EMULATION
            if (tid < 32 && t + gridDim.x < a.tile_end) {
                fence_proxy_async_smem(); // the generic-proxy reads above are ordered before the engine's writes
                issue_tile_load(t + gridDim.x);
            }
#endif
        };
#ifdef SPZ_CPU_EMULATION
        if (tid == 0) { // the emulation's "TMA": rows of 16 doubles, 16-byte chunks XORed with the row number mod 8
            for (unsigned j = 0; j < kTileLen3; ++j) {
                const unsigned long long g = base + tile_to_global(j);
                sre[swz3(j)] = a.re[g];
                sim[swz3(j)] = a.im[g];
            }
        }
#endif
        // ---- while the tile is in flight: skip flags and per-tile constants, from the staged program ----
        {
            const unsigned long long *outer = reinterpret_cast<const unsigned long long *>(smem + kProgOff3 + a.outer_off);
            for (int i = tid; i < a.n_ins; i += kThreads3) {
                const unsigned long long ocm = outer[i];
                skip[i] = (base & ocm) != ocm ? 1 : 0;
            }
            // Per-tile constants: every group's product over those of its terms whose outer bits are set in this tile.
            const TileGroup *groups = reinterpret_cast<const TileGroup *>(smem + kProgOff3 + a.groups_off);
            const TileTerm *terms = reinterpret_cast<const TileTerm *>(smem + kProgOff3 + a.terms_off);
#ifndef SPZ_CPU_EMULATION
            // one warp per group: a lane per term, then a butterfly product over the lanes (5 shuffle rounds).  One thread per
            // group walking its ~18 terms took 2.5 us per tile of a QFT pass, more than the tile takes to arrive.
            for (int g = tid >> 5; g < a.n_groups; g += kThreads3 / 32) {
                const TileGroup gd = groups[g];
                double fr = 1.0, fi = 0.0;
                for (int i = tid & 31; i < gd.count; i += 32) {
                    const TileTerm &tm = terms[gd.first + i];
                    if ((base & tm.outer) == tm.outer) cmul3(fr, fi, tm.fr, tm.fi);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    const double qr = __shfl_xor_sync(0xffffffffu, fr, o), qi = __shfl_xor_sync(0xffffffffu, fi, o);
                    cmul3(fr, fi, qr, qi); // (every lane ends with the same product up to the order of the factors; lane 0 writes)
                }
                if ((tid & 31) == 0) gfac[g] = make_double2(fr, fi);
            }
#else
            for (int g = tid; g < a.n_groups; g += kThreads3) {
                const TileGroup gd = groups[g];
                double fr = 1.0, fi = 0.0;
                for (int i = 0; i < gd.count; ++i) {
                    const TileTerm &tm = terms[gd.first + i];
                    if ((base & tm.outer) == tm.outer) cmul3(fr, fi, tm.fr, tm.fi);
                }
                gfac[g] = make_double2(fr, fi);
            }
#endif
        }
        sts_d2(facc_addr, a.scale, 0.0); // F0 starts as the pass scale (pending from the start when it is not 1: the host knows)
        __syncthreads(); // skip flags and constants in place (emulation: the tile too)
        stamp(1); // constants
#ifndef SPZ_CPU_EMULATION
        // One warp waits for the boxes (a failed try_wait costs issue slots the other CTA of this SM could use); after the
        // barrier the phase is complete and every thread's own try_wait -- its acquire of the TMA writes -- succeeds at once.
        if (tid < 32) mbar_wait(bar, parity);
        __syncthreads();
        mbar_wait(bar, parity);
        parity ^= 1u;
#endif
        stamp(2); // tile arrived

        lay = sins[0].a;
        load_regs();
        if (a.single_layout) prefetch();

        // The interpreter.  Dispatch needs one byte per instruction -- the arm -- and takes it from a register: the first word of
        // the NEXT instruction is fetched before the current one runs (the blob is padded, so the fetch past END reads zeros).
        // An arm that has operands loads its own 16-byte instruction: x = arm | kind << 8 | class << 16 | flags << 24, y = pair
        // mask | thread mask << 16, z = pool offset (doubles) / layout, w = per-tile constant index; the hot H on every pair has
        // none.  Everything is addressed by 32-bit shared addresses computed once.  (The arms take nothing from the prefetched
        // word but the arm byte: a second use would cost a register copy per interpreted instruction.)
        unsigned ip = ins_addr + kIns3;
        unsigned xn = lds32(ip);
        for (;; ip += kIns3) {
            const unsigned arm = xn & 0xffu;
            xn = lds32(ip + kIns3);
#define SPZ_W const uint4 w = lds128(ip); const unsigned x = w.x
            // guarded GATE / ACCG / OTHER act on the threads whose control bits are set; a GATE also needs its controls outside
            // the tile.  The unguarded arms (no control of any kind: the host routes everything else to the guarded ones) test nothing.
#define SPZ_OK(w) (((tid & ((w).y >> 16)) == ((w).y >> 16)) && !(((x >> 24) & GF_OUTER) && lds8(skip_addr + ((ip - ins_addr) >> 5))))
#define SPZ_GATE_ALL(V, MK)                                                                                  \
    case T3_GATE + 4 * V + 0: { bfly3<MK, 0, true>(ar, ai, ip + 16u, 0xffffu); break; }             \
    case T3_GATE + 4 * V + 1: { bfly3<MK, 1, true>(ar, ai, ip + 16u, 0xffffu); break; }             \
    case T3_GATE + 4 * V + 2: { bfly3<MK, 2, true>(ar, ai, ip + 16u, 0xffffu); break; }             \
    case T3_GATE + 4 * V + 3: { bfly3<MK, 3, true>(ar, ai, ip + 16u, 0xffffu); break; }          
#define SPZ_GATE_GUARDED(V, MK)                                                                                                       \
    case T3_GATE + 4 * V + 0: { SPZ_W; if (SPZ_OK(w)) bfly3<MK, 0, false>(ar, ai, ip + 16u, w.y & 0xffffu); break; }      \
    case T3_GATE + 4 * V + 1: { SPZ_W; if (SPZ_OK(w)) bfly3<MK, 1, false>(ar, ai, ip + 16u, w.y & 0xffffu); break; }      \
    case T3_GATE + 4 * V + 2: { SPZ_W; if (SPZ_OK(w)) bfly3<MK, 2, false>(ar, ai, ip + 16u, w.y & 0xffffu); break; }      \
    case T3_GATE + 4 * V + 3: { SPZ_W; if (SPZ_OK(w)) bfly3<MK, 3, false>(ar, ai, ip + 16u, w.y & 0xffffu); break; }
#define SPZ_PERM_GUARDED(V, MK)                                                                              \
    case T3_GATE + 4 * V + 0: { SPZ_W; if (SPZ_OK(w)) perm3<MK, 0>(ar, ai, w.y & 0xffffu); break; }          \
    case T3_GATE + 4 * V + 1: { SPZ_W; if (SPZ_OK(w)) perm3<MK, 1>(ar, ai, w.y & 0xffffu); break; }          \
    case T3_GATE + 4 * V + 2: { SPZ_W; if (SPZ_OK(w)) perm3<MK, 2>(ar, ai, w.y & 0xffffu); break; }          \
    case T3_GATE + 4 * V + 3: { SPZ_W; if (SPZ_OK(w)) perm3<MK, 3>(ar, ai, w.y & 0xffffu); break; }
            // Accumulator F_{r+1} -- or a constant of the pool, when the host folded the pending factor (AF_CONST) -- is pending
            // and a butterfly on register bit r follows: apply it to the amplitudes with that bit set.  Only the accumulator of the
            // target's own bit separates the two members of a pair; the others scale both by the same factor and stay pending.
#define SPZ_PRE(R)                                                                                                                    \
    case T3_PRE + R: {                                                                                                                \
        SPZ_W;                                                                                                                        \
        const unsigned src = ((x >> 24) & AF_CONST) ? ip + 16u : (facc_base + 16u * opaque(tid)) + (R + 1) * 16u * kThreads3;                      \
        apply_bit3<R>(ar, ai, lds_d2(src));                                                                                           \
        break; }
            // nvcc lowers a switch to a tree of compares and branches, never to a jump table; the build rewrites the tree that
            // follows this marker into one indirect branch (brx.idx / BRX) in the PTX: spinoza_b200/ptx_jump_table.py
#ifndef SPZ_CPU_EMULATION
            asm volatile("// SPZ_JUMP_TABLE 41");
#endif
            static_assert(T3_N_ARMS == 41, "the jump-table marker above carries the number of arms");
            switch (arm) {
            // H on every pair has no operand at all
            case T3_GATE + 0: bfly3<MK_H, 0, true>(ar, ai, 0u, 0xffffu); break;
            case T3_GATE + 1: bfly3<MK_H, 1, true>(ar, ai, 0u, 0xffffu); break;
            case T3_GATE + 2: bfly3<MK_H, 2, true>(ar, ai, 0u, 0xffffu); break;
            case T3_GATE + 3: bfly3<MK_H, 3, true>(ar, ai, 0u, 0xffffu); break;
            SPZ_GATE_ALL(1, MK_RX)
            SPZ_GATE_ALL(2, MK_RY)
            SPZ_GATE_GUARDED(3, MK_HS)
            SPZ_GATE_GUARDED(4, MK_RX)
            SPZ_GATE_GUARDED(5, MK_RY)
            SPZ_PERM_GUARDED(6, MK_X)
            SPZ_PERM_GUARDED(7, MK_Y)
            SPZ_PRE(0)
            SPZ_PRE(1)
            SPZ_PRE(2)
            SPZ_PRE(3)
            case T3_ACC: {
                SPZ_W;
                const unsigned flags = x >> 24, tab = pool_addr + 8u * w.z;
                double fr = 1.0, fi = 0.0;
                if (flags & AF_LO) { const double2 v = lds_d2(tab + 16u * (tid & 15u)); fr = v.x; fi = v.y; }
                if (flags & AF_HI) { const double2 v = lds_d2(tab + 256u + 16u * (tid >> 4)); cmul3(fr, fi, v.x, v.y); }
                if (flags & AF_TILE) { const double2 v = lds_d2(gfac_addr + 16u * w.w); cmul3(fr, fi, v.x, v.y); }
                acc((x >> 16) & 0xffu, flags & AF_SET, fr, fi);
                break; }
            case T3_ACCG: { // a term that needs thread bits from both nibbles, or thread bits and bits outside the tile
                SPZ_W;
                const unsigned flags = x >> 24, cls = (x >> 16) & 0xffu;
                const bool hit = SPZ_OK(w);
                const unsigned src = (flags & AF_TILE) ? gfac_addr + 16u * w.w : ip + 16u;
                if (flags & AF_SET) {
                    // the accumulator held no pending factor: every thread assigns (nothing is ever reset, see flush)
                    const double2 v = hit ? lds_d2(src) : make_double2(1.0, 0.0);
                    sts_d2((facc_base + 16u * opaque(tid)) + cls * 16u * kThreads3, v.x, v.y);
                } else if (hit) {
                    const double2 v = lds_d2(src);
                    acc(cls, false, v.x, v.y);
                }
                break; }
            case T3_OTHER: { // a diagonal term over two or more register bits: applied at once to the amplitudes it selects
                SPZ_W;
                if (!SPZ_OK(w)) break;
                const double2 f = lds_d2(((x >> 24) & AF_TILE) ? gfac_addr + 16u * w.w : ip + 16u);
                const unsigned km = w.y & 0xffffu;
#define SPZ_M4(A, B, C, D) cmul3(ar[A], ai[A], f.x, f.y); cmul3(ar[B], ai[B], f.x, f.y); cmul3(ar[C], ai[C], f.x, f.y); cmul3(ar[D], ai[D], f.x, f.y)
                switch (km) {
                case 3: SPZ_M4(3, 7, 11, 15); break;
                case 5: SPZ_M4(5, 7, 13, 15); break;
                case 6: SPZ_M4(6, 7, 14, 15); break;
                case 9: SPZ_M4(9, 11, 13, 15); break;
                case 10: SPZ_M4(10, 11, 14, 15); break;
                case 12: SPZ_M4(12, 13, 14, 15); break;
                default:
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (((unsigned)k & km) == km) cmul3(ar[k], ai[k], f.x, f.y);
                    break;
                }
#undef SPZ_M4
                break; }
            case T3_LAYOUT: { // apply what is pending, then change the register-resident bits through shared memory
                SPZ_W;
                flush((x >> 24) & 31u, w.y & 31u, pool_addr + 8u * w.w);
                store_regs();
                __syncthreads();
                lay = w.z;
                load_regs(); // no second barrier: this thread's next shared-memory access is store_regs() to the cells it has just read
                if ((x >> 24) & LF_LAST) prefetch();
                break; }
            default: { // END
                SPZ_W;
                flush((x >> 24) & 31u, w.y & 31u, pool_addr + 8u * w.w);
                goto program_done; }
            }
#undef SPZ_GATE_ALL
#undef SPZ_GATE_GUARDED
#undef SPZ_PERM_GUARDED
#undef SPZ_PRE
#undef SPZ_OK
#undef SPZ_W
        }
    program_done:
        stamp(3); // program interpreted
        direct_store(base);
        stamp(4); // stores issued
        __syncthreads(); // the next tile's flags and constants overwrite what slower warps may still be reading
    }
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
    unsigned dirty = 0;       // accumulators that hold a pending factor
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

    // ---- diagonal terms: amplitudes whose tile base contains `outer`, whose THREAD ID contains `thr` and whose register index
    // contains `m` are multiplied by (fr, fi) ----
    struct Term { uint64_t outer; uint32_t thr, m; double fr, fi; };
    auto class_of = [](uint32_t m) { return m == 0 ? 0 : m == 1 ? 1 : m == 2 ? 2 : m == 4 ? 3 : m == 8 ? 4 : 5; };
    // Factors that depend on nothing but a register bit (no thread bit, no bit outside the tile: an RZ or P on a register-resident
    // qubit, the global phase of any RZ) are the same number for every thread of every tile, so the HOST accumulates them, one
    // pending constant per class, and the device sees them once, where the class is consumed: as the operand of the PRE in front of
    // a butterfly on that bit, or as one ACCG in front of a flush.  In a layered circuit that is a quarter of all instructions.
    double hostK[5][2] = {{1.0, 0.0}, {1.0, 0.0}, {1.0, 0.0}, {1.0, 0.0}, {1.0, 0.0}};
    bool kpend[5] = {false, false, false, false, false};
    auto fold_const = [&](int cls) { // the pending constant of a class joins its accumulator
        if (!kpend[cls]) return;
        Ins3 i{};
        i.op = T3_ACCG;
        i.rpos = (uint8_t)cls;
        i.flags = AF_CONST;
        i.s[0] = hostK[cls][0]; i.s[1] = hostK[cls][1];
        if (!(dirty & (1u << cls))) i.flags |= AF_SET;
        dirty |= 1u << cls;
        push(i, 0);
        hostK[cls][0] = 1.0; hostK[cls][1] = 0.0; kpend[cls] = false;
    };
    // LAYOUT / END apply everything that is pending: flags = the classes, km = those of them whose factor is a host constant
    // (a class with both: the constant joins the accumulator first), b = the five-entry constant table in the pool
    auto flush_operands = [&](Ins3 &i) {
        unsigned cmask = 0;
        for (int cls = 0; cls < 5; ++cls) {
            if (kpend[cls] && (dirty & (1u << cls))) fold_const(cls);
            if (kpend[cls]) cmask |= 1u << cls;
        }
        i.flags = (uint8_t)(dirty | cmask);
        i.km = (uint16_t)cmask;
        if (cmask) {
            i.b = pool2(hostK[0][0], hostK[0][1]);
            for (int cls = 1; cls < 5; ++cls) pool2(hostK[cls][0], hostK[cls][1]);
            for (int cls = 0; cls < 5; ++cls) { hostK[cls][0] = 1.0; hostK[cls][1] = 0.0; kpend[cls] = false; }
        }
    };
    auto emit_terms = [&](const std::vector<Term> &all) {
        for (int cls = 0; cls < 6; ++cls) {
            std::vector<Term> ts;
            for (const Term &x : all) {
                if (class_of(x.m) != cls) continue;
                if (cls < 5 && x.outer == 0 && x.thr == 0) { cmul_h(hostK[cls][0], hostK[cls][1], x.fr, x.fi); kpend[cls] = true; continue; }
                ts.push_back(x);
            }
            if (ts.empty()) continue;
            if (cls < 5) {
                double lo[16][2], hi[16][2];
                for (int e = 0; e < 16; ++e) { lo[e][0] = hi[e][0] = 1.0; lo[e][1] = hi[e][1] = 0.0; }
                bool use_lo = false, use_hi = false;
                std::vector<TileTerm> tile_terms;                                  // thr == 0, outer != 0: one constant per tile
                std::vector<std::pair<uint32_t, std::vector<TileTerm>>> general;   // everything else, by thread mask
                // a lone table-able term is cheaper as a constant (16 bytes) than as a table (256-512 bytes of shared memory)
                int n_tabled = 0;
                for (const Term &x : ts) n_tabled += x.outer == 0 && ((x.thr & 0xF0u) == 0 || (x.thr & 0x0Fu) == 0);
                for (const Term &x : ts) {
                    if (n_tabled == 1 && x.outer == 0 && ((x.thr & 0xF0u) == 0 || (x.thr & 0x0Fu) == 0)) {
                        general.emplace_back(x.thr, std::vector<TileTerm>{TileTerm{0, 0, 0, x.fr, x.fi}});
                    } else if (x.outer == 0 && (x.thr & 0xF0u) == 0) {
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
                        i.flags |= AF_CONST; i.s[0] = fr; i.s[1] = fi;
                    }
                    if (!(dirty & (1u << cls))) i.flags |= AF_SET;
                    dirty |= 1u << cls;
                    push(i, 0);
                }
            } else {
                // two or more register bits: by (thread mask, register mask), applied to the amplitudes at once
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
                        i.flags |= AF_CONST; i.s[0] = fr; i.s[1] = fi;
                    }
                    push(i, 0);
                }
            }
        }
    };

    // ---- non-diagonal gates: the butterfly variant, its two scalars, what it leaves to the pass scale, and the phase terms
    // around it ----
    struct Variant {
        int kind = -1;
        double s[2] = {0.0, 0.0};
        int ns = 0;
        double factor = 1.0;      // multiplies the pass scale (uncontrolled gates only)
        bool ctrl_sign = false;   // controlled rotation by more than a quarter turn: -1 on the control subspace
        bool pre = false, post = false;  // U = diag(1, post) . rotation . diag(1, pre)
        double pre_f[2] = {1.0, 0.0}, post_f[2] = {1.0, 0.0};
    };
    // `budget`: how much smaller the pass scale may still get (the amplitudes grow by the inverse until the scale is applied)
    auto variant_of = [](const TileInstr &t, double budget) -> Variant {
        Variant v;
        const bool controlled = t.reg_cmask || t.thr_cmask || t.outer_cmask;
        const double *s = t.s;
        // rotation [[cs, -sn], [sn, cs]] as three shears with |angle| <= pi/2: -R(angle -+ pi) when cs < 0
        auto rotation = [&](int kind, double cs, double sn) {
            v.kind = kind;
            if (cs < 0.0) {
                cs = -cs; sn = -sn;
                if (controlled) v.ctrl_sign = true; else v.factor = -1.0;
            }
            v.s[0] = sn / (1.0 + cs); v.s[1] = sn; v.ns = 2;
        };
        switch (t.kind) {
        case SPZ_GATE_H:
            if (!controlled && SPZ_SQRT_ONE_HALF >= budget) { v.kind = MK_H; v.factor = SPZ_SQRT_ONE_HALF; }
            else v.kind = MK_HS;
            break;
        case SPZ_GATE_X: v.kind = MK_X; break;
        case SPZ_GATE_Y: v.kind = MK_Y; break;
        case SPZ_GATE_RX: rotation(MK_RX, s[0], s[1]); break; // [[cs, i ns], [i ns, cs]], s = (cs, ns)   (gate_math.cuh)
        case SPZ_GATE_RY: rotation(MK_RY, s[1], s[0]); break; // [[cs, -sn], [sn, cs]], s = (sn, cs)
        case SPZ_GATE_U: { // [[ct, k + i l], [q + i r, ss + i tt]] = diag(1, e^{i phi}) [[ct, -st], [st, ct]] diag(1, e^{i lambda})
            const double ct = s[0], k = s[1], l = s[2], q = s[3], r = s[4], ss = s[5], tt = s[6];
            const double st = std::hypot(q, r); // |sin|: its sign is absorbed by the two phases
            if (st > 1e-150) {
                v.post_f[0] = q / st; v.post_f[1] = r / st;
                v.pre_f[0] = -k / st; v.pre_f[1] = -l / st;
            } else { // diagonal: diag(ct, ss + i tt), |ct| = 1
                v.post_f[0] = ss / ct; v.post_f[1] = tt / ct;
            }
            v.pre = !(v.pre_f[0] == 1.0 && v.pre_f[1] == 0.0);
            v.post = !(v.post_f[0] == 1.0 && v.post_f[1] == 0.0);
            rotation(MK_RY, ct, st);
            break; }
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
            if (pc > 0) flush_operands(i);
            else i.flags = (uint8_t)dirty;
            for (int k = 0; k < 4; ++k) { R[k] = t.rbit[k]; if (R[k] < 0 || R[k] >= kT3 || (k && R[k] <= R[k - 1])) return false; }
            i.a = (uint32_t)R[0] | ((uint32_t)R[1] << 8) | ((uint32_t)R[2] << 16) | ((uint32_t)R[3] << 24);
            if (pc > 0) dirty = 0;
            push(i, 0);
            continue;
        }
        if (t.op == TI_GATE) {
            const Variant &v = variants[pc];
            if (v.kind < 0 || t.rpos < 0 || t.rpos > 3) return false;
            bool ok = true;
            const uint32_t thr = to_tid(t.thr_cmask, ok);
            if (!ok) return false;
            const uint32_t tbit = 1u << t.rpos;
            std::vector<Term> before, after;
            if (v.pre) before.push_back(Term{t.outer_cmask, thr, t.reg_cmask | tbit, v.pre_f[0], v.pre_f[1]});
            if (v.ctrl_sign) before.push_back(Term{t.outer_cmask, thr, t.reg_cmask, -1.0, 0.0});
            if (v.post) after.push_back(Term{t.outer_cmask, thr, t.reg_cmask | tbit, v.post_f[0], v.post_f[1]});
            if (!before.empty()) emit_terms(before);
            // the factor pending on the target's own register bit is applied first: the class's accumulator, or the host's constant
            // (both: the constant joins the accumulator)
            if (kpend[t.rpos + 1] && (dirty & (2u << t.rpos))) fold_const(t.rpos + 1);
            if (kpend[t.rpos + 1] || (dirty & (2u << t.rpos))) {
                Ins3 p{};
                p.op = (uint8_t)(T3_PRE + t.rpos);
                p.rpos = (uint8_t)t.rpos;
                if (kpend[t.rpos + 1]) {
                    double (&K)[2] = hostK[t.rpos + 1];
                    p.flags = AF_CONST;
                    p.s[0] = K[0]; p.s[1] = K[1];
                    K[0] = 1.0; K[1] = 0.0; kpend[t.rpos + 1] = false;
                }
                push(p, 0);
                dirty &= ~(2u << t.rpos);
            }
            Ins3 i{};
            i.rpos = (uint8_t)t.rpos;
            i.thr = (uint16_t)thr;
            i.km = (uint16_t)t.t_mask;
            const bool in_tile_ctrl = t.reg_cmask || t.thr_cmask;
            if (in_tile_ctrl) out.ctrl = true;
            if (t.outer_cmask) i.flags |= GF_OUTER;
            i.kind = (uint8_t)v.kind;
            // the arm: unguarded for the three gates that fill a circuit when nothing in the tile controls them, else under the
            // pair mask (which is all ones for an uncontrolled X, Y or scaled H)
            int variant;
            if (!in_tile_ctrl && !t.outer_cmask && (v.kind == MK_H || v.kind == MK_RX || v.kind == MK_RY)) {
                variant = v.kind == MK_H ? 0 : v.kind == MK_RX ? 1 : 2;
                i.flags |= GF_ALL;
            } else {
                if (v.kind == MK_H) return false; // (cannot happen: an H under any control is lowered to HS)
                variant = v.kind == MK_HS ? 3 : v.kind == MK_RX ? 4 : v.kind == MK_RY ? 5 : v.kind == MK_X ? 6 : 7;
            }
            i.op = (uint8_t)(T3_GATE + 4 * variant + t.rpos);
            if (v.ns) { i.s[0] = v.s[0]; i.s[1] = v.s[1]; }
            push(i, t.outer_cmask);
            if (!after.empty()) emit_terms(after);
            continue;
        }
        if (t.op != TI_RUN) return false; // TI_DIAG: exact mode
        // ---- a merged run of diagonal gates: groups [rpos, rpos + sum of counts), classes m = 0, 1, 2, 4, 8, other ----
        const int n_run_groups = t.rbit[0] + t.rbit[1] + t.rbit[2] + t.rbit[3] + (int)t.reg_cmask + (int)t.thr_cmask;
        std::vector<Term> run;
        for (int g = t.rpos; g < t.rpos + n_run_groups; ++g) {
            const TileGroup &gd = groups[g];
            for (int j = 0; j < gd.count; ++j) {
                const TileTerm &tm = terms[gd.first + j];
                bool ok = true;
                const uint32_t thr = to_tid(tm.thr, ok);
                if (!ok) return false;
                run.push_back(Term{tm.outer, thr, tm.m, tm.fr, tm.fi});
            }
        }
        emit_terms(run);
    }
    Ins3 e{};
    e.op = T3_END;
    flush_operands(e);
    push(e, 0);
    for (size_t k = out.ins.size(); k-- > 1;)
        if (out.ins[k].op == T3_LAYOUT) { out.ins[k].flags |= LF_LAST; out.single_layout = false; break; }
    if (out.pool.size() & 1) out.pool.push_back(0.0);
    return out.ins.size() <= (size_t)kMaxIns3 && out.groups.size() <= (size_t)kMaxGroups3;
}

// Serialise for upload (layout: see Lowered3) and fill the size fields of the kernel arguments.
size_t tile3_pack(const Lowered3 &lw, const TilePlan &plan, std::vector<unsigned char> &blob, Tile3Args &a) {
    a.L = plan.low_bits; a.n_high = plan.n_high;
    for (int k = 0; k < plan.n_high; ++k) a.high[k] = plan.high[k];
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t ins_bytes = up16(lw.ins.size() * sizeof(Ins3));
    const size_t pool_bytes = up16(lw.pool.size() * sizeof(double));
    const size_t outer_off = ins_bytes + pool_bytes;
    const size_t groups_off = up16(outer_off + lw.outer.size() * sizeof(uint64_t));
    const size_t terms_off = up16(groups_off + lw.groups.size() * sizeof(TileGroup));
    const size_t total = up16(terms_off + lw.terms.size() * sizeof(TileTerm)) + kIns3; // (the interpreter prefetches one word past END)
    blob.assign(total, 0);
    std::memcpy(blob.data(), lw.ins.data(), lw.ins.size() * sizeof(Ins3));
    if (!lw.pool.empty()) std::memcpy(blob.data() + ins_bytes, lw.pool.data(), lw.pool.size() * sizeof(double));
    std::memcpy(blob.data() + outer_off, lw.outer.data(), lw.outer.size() * sizeof(uint64_t));
    if (!lw.groups.empty()) std::memcpy(blob.data() + groups_off, lw.groups.data(), lw.groups.size() * sizeof(TileGroup));
    if (!lw.terms.empty()) std::memcpy(blob.data() + terms_off, lw.terms.data(), lw.terms.size() * sizeof(TileTerm));
    a.ins_bytes = (unsigned)ins_bytes; a.blob_bytes = (unsigned)total;
    a.outer_off = (unsigned)outer_off; a.groups_off = (unsigned)groups_off; a.terms_off = (unsigned)terms_off;
    a.n_ins = (int)lw.ins.size(); a.n_groups = (int)lw.groups.size(); a.n_terms = (int)lw.terms.size();
    a.scale = lw.scale;
    a.single_layout = lw.single_layout ? 1 : 0;
    return total;
}

size_t tile3_smem_bytes(const Tile3Args &a) {
    return ((kProgOff3 + a.blob_bytes + 16u * (size_t)a.n_groups + (size_t)a.n_ins + 15u) & ~(size_t)15u) + 16u + 8u * ((size_t)1 << a.n_high);
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

// How a tile becomes TMA boxes.  The state array is viewed as a tensor whose dimensions are bit fields of the amplitude index:
// dim 0 = bits [0, 4) (16 doubles = 128 bytes, the widest inner box the 128-byte swizzle allows), dim 1 = bits [4, p1) up to the
// first high tile qubit (the box takes the 2^(L-4) rows of a segment), then one dimension per run of contiguous high tile
// qubits, spanning from the run's first bit to the next run (the box takes the run; the bits above it, which are not tile
// qubits, are the coordinate).  A tensor map has at most five dimensions, so a box holds the low bits and the three lowest
// runs; higher tile qubits are enumerated by the kernel, one box per combination.  Tiles whose high qubits form up to three
// runs -- {24..29}, {12, 13, 20..23}, ... -- move with ONE instruction per array and direction.
static void plan_boxes(int n_qubits, const TilePlan &plan, Tile3Args &a) {
    const int L = plan.low_bits;
    struct Run { int lo, w; };
    Run runs[kMaxHigh3];
    int n_runs = 0;
    for (int k = 0; k < plan.n_high; ++k) {
        if (n_runs && runs[n_runs - 1].lo + runs[n_runs - 1].w == plan.high[k]) ++runs[n_runs - 1].w;
        else runs[n_runs++] = Run{plan.high[k], 1};
    }
    const int in_box = std::min(n_runs, 3);
    a.rank = 2 + in_box;
    a.box_shift = L;
    a.dim_lo[0] = 0; a.dim_len[0] = 4;
    a.dim_lo[1] = 4; a.dim_len[1] = (in_box ? runs[0].lo : n_qubits) - 4;
    for (int r = 0; r < in_box; ++r) {
        a.dim_lo[2 + r] = runs[r].lo;
        a.dim_len[2 + r] = (r + 1 < in_box ? runs[r + 1].lo : n_qubits) - runs[r].lo;
        a.box_shift += runs[r].w;
    }
    for (int i = a.rank; i < 5; ++i) { a.dim_lo[i] = 0; a.dim_len[i] = 0; }
    a.n_rest = 0;
    for (int r = in_box; r < n_runs; ++r)
        for (int b = 0; b < runs[r].w; ++b) a.rest_bit[a.n_rest++] = runs[r].lo + b;
}

static int make_tensor_map(CUtensorMap *tm, double *base, const TilePlan &plan, const Tile3Args &a) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SPZ_ERR_CUDA; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5];
    for (int i = 0; i < a.rank; ++i) {
        gdim[i] = (cuuint64_t)1 << a.dim_len[i];
        estr[i] = 1;
        if (i > 0) gstride[i - 1] = (cuuint64_t)8 << a.dim_lo[i];
    }
    box[0] = 16;
    box[1] = 1u << (plan.low_bits - 4);
    int k = 0;
    for (int i = 2; i < a.rank; ++i) { // the run that starts at dim_lo[i]
        int w = 0;
        while (k < plan.n_high && plan.high[k] == a.dim_lo[i] + w) { ++w; ++k; }
        box[i] = 1u << w;
    }
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)a.rank, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for rank %d, L=%d", (int)r, a.rank, plan.low_bits); return SPZ_ERR_CUDA; }
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
    const size_t bytes = tile3_pack(lw, plan, blob, a);
    const size_t smem = tile3_smem_bytes(a);
    if (smem > kSmemBudget3) return SPZ_OK;
    static bool prepared[64] = {false};
    if (st->device >= 0 && st->device < 64 && !prepared[st->device]) {
        SPZ_TRY(tile3_prepare());
        prepared[st->device] = true;
    }
    plan_boxes(st->n, plan, a);
    SPZ_TRY(make_tensor_map(&a.tm_re, st->re, plan, a));
    SPZ_TRY(make_tensor_map(&a.tm_im, st->im, plan, a));
    char *slot = nullptr;
    SPZ_TRY(tile_ring_alloc(st, bytes, &slot));
    SPZ_CUDA(cudaMemcpyAsync(slot, blob.data(), bytes, cudaMemcpyHostToDevice, st->stream));
    a.re = st->re; a.im = st->im;
    a.prof = nullptr;
    if (const char *e = std::getenv("SPZ_TILE_PROF")) {
        if (e[0] == '1') { // diagnostic: the previous pass's phase times are printed when the next one is prepared
            static unsigned long long *d_prof = nullptr;
            static long long n_cta = 0;
            if (!d_prof) SPZ_CUDA(cudaMalloc(&d_prof, 8 * sizeof(unsigned long long)));
            else {
                unsigned long long h[8];
                SPZ_CUDA(cudaStreamSynchronize(st->stream));
                SPZ_CUDA(cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost));
                if (n_cta) std::fprintf(stderr, "k_tile3 phases, ns per tile: program (once per CTA) %.0f, constants %.0f, tile wait %.0f, interpret %.0f, stores %.0f\n",
                                        (double)h[0] / n_cta, (double)h[1] / n_cta, (double)h[2] / n_cta, (double)h[3] / n_cta, (double)h[4] / n_cta);
            }
            SPZ_CUDA(cudaMemsetAsync(d_prof, 0, 8 * sizeof(unsigned long long), st->stream));
            n_cta = st->len >> kT3;
            a.prof = d_prof;
        }
    }
    a.blob = reinterpret_cast<const unsigned char *>(slot);
    static_assert(sizeof(Tile3Args) <= sizeof(out->args), "Tile3Launch::args too small");
    std::memcpy(out->args, &a, sizeof a);
    out->smem = smem;
    *handled = true;
    return SPZ_OK;
}

void run_tile3(spz_state *st, const Tile3Launch &l, unsigned first, unsigned count) {
    Tile3Args a;
    std::memcpy(&a, l.args, sizeof a);
    a.tile_first = first;
    a.tile_end = first + count;
    // persistent: two CTAs per SM walk the tiles of the launch
    static int sms[64] = {0};
    int n_sm = 148;
    if (st->device >= 0 && st->device < 64) {
        if (!sms[st->device]) cudaDeviceGetAttribute(&sms[st->device], cudaDevAttrMultiProcessorCount, st->device);
        if (sms[st->device] > 0) n_sm = sms[st->device];
    }
    const unsigned grid = std::min<unsigned>(count, 2u * (unsigned)n_sm);
    k_tile3<<<grid, kThreads3, l.smem, st->stream>>>(a);
}
#endif // !SPZ_CPU_EMULATION

} // namespace spz
