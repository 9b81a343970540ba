// kernels_tile2.cu -- second generation of the fused tile kernel (see kernels_tile.cu for the execution model).
//
// STATUS: opt-in (environment SPZ_TILE_V2=1).  Written at the end of round 1 from the ncu / SASS analysis of k_tile
// (profiles/round1_summary.md sections 3 and 10) when no GPU time was left.  It compiles for sm_100a without spills, and its
// body runs correctly -- bit-identical to the oracle in exact mode, race-free under ThreadSanitizer -- on the CPU emulation
// of tests/emu/ (tests/test_tile_cpu_emulation.py); it has NOT run on a B200 yet, so the default stays k_tile until
// tools/round2_first_call.sh (tests/test_gpu_tile_v2.py + A/B timing) has been run on hardware.
//
// Same micro-program (TileInstr / TileGroup / TileTerm, engine.h), same results as k_tile up to the documented merged
// rounding.  What changes is the instruction count per tile (k_tile: ~5 200 warp instructions per warp for a 27-gate QFT
// pass, ~7 900 for the 276-gate one, against ~1 000 of irreducible work):
//   1. Lazy phase flush.  A butterfly on register bit r only needs accumulator F_{r+1} applied first: F0 and the other
//      F_i multiply both members of every pair (k, k | 1 << r) by the same factor, which commutes with any (controlled)
//      2x2 gate on the pair.  k_tile expands all five accumulators before every butterfly (up to 31 complex multiplies);
//      here a butterfly costs at most 8, and the full expansion happens once per register layout.
//   2. Direct global <-> register transfers.  When the first (last) register layout keeps consecutive lanes on consecutive
//      amplitudes (all register bits >= 4), the tile is loaded (stored) straight into (from) the registers: no staging
//      store, no barrier, no swizzled re-read.  Layouts holding tile bits 0,1 can use 256/128-bit accesses instead
//      (SPZ_TILE_V2_DIRECT, see tile2_make_args).
//   3. The tile's loads are issued first; staging the program and reducing the phase groups run under the HBM latency.
//   4. The program is always decoded from shared memory (typed LDS instead of generic loads; longer programs fall back to
//      k_tile), and per-instruction flags computed once per tile replace the 64-bit tile base in the interpreter loop.
//   5. CTRL = false instantiation for passes whose butterflies have no in-tile controls (every QFT pass): the pair mask is
//      then the same for every thread, the per-pair guards compile to uniform branches (no convergence barriers) and the
//      pair updates stay in place.
//   6. One barrier per layout change and none after a register load: a thread's next shared-memory access after
//      load_regs() is store_regs() under the same layout, i.e. to the cells it has just read.
//   7. The per-tile reduction of the phase groups is parallel over TERMS (factors parked in the still unused tile region),
//      one L2 round trip for all of them instead of one per term in series inside each group.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "gate_math.cuh"

// tests/emu/ compiles this file with g++ (SPZ_CPU_EMULATION, one OS thread per CUDA thread, a pthread barrier for
// __syncthreads) so that the kernel body itself -- not a restatement of it -- is checked on the CPU; only the way the
// dynamic shared-memory window is named differs between the two builds.
#ifdef SPZ_CPU_EMULATION
#define SPZ_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(spz_emu::dyn_smem)
#else
#define SPZ_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

namespace spz {

namespace {

constexpr int kT2 = 12;          // full tiles only (registers of >= 12 qubits)
constexpr int kThreads2 = 256;   // 2^(kT2 - 4)
constexpr int kMaxHigh2 = 8;
constexpr int kMaxInstr2 = 256;  // one flag per instruction is computed by thread pc; the real limit is kSmemBudget2
// Two CTAs per SM (228 KB of shared memory, 1 KB reserved per CTA, ~2.3 KB static here): a pass whose tile + group table +
// program need more than this falls back to k_tile.
constexpr size_t kSmemBudget2 = 110 * 1024;

struct Tile2Args {
    double *re;
    double *im;
    const TileInstr *prog;
    const TileGroup *groups;
    const TileTerm *terms;
    int n_instr;
    int n_groups;
    int n_terms;
    unsigned tile_offset;
    unsigned prog_off;   // byte offset of the staged program inside dynamic shared memory
    int L, n_high;
    // how the first layout is loaded / the last one stored: 0 = staged through shared memory; direct global <-> registers
    // with 1 = one amplitude, 2 = two (register bit 0 is tile bit 0), 4 = four (register bits 0,1 are tile bits 0,1) per
    // access.  The host decides (tile2_make_args).
    int first_direct;
    int last_direct;
    unsigned first_rbits; // the first layout's register bits, one byte each (= prog[0].rbit, known before the program is staged)
    unsigned last_rbits;  // the last layout's
    int high[kMaxHigh2];
};

__device__ __forceinline__ unsigned swz2(unsigned j) { return j ^ (((j >> 4) ^ (j >> 8)) & 15u); }

__device__ __forceinline__ void cmul2(double &xr, double &xi, double fr, double fi) {
    const double nr = xr * fr - xi * fi;
    const double ni = xr * fi + xi * fr;
    xr = nr; xi = ni;
}

// 256-bit global access (one full 32-byte sector per lane); plain loads in the CPU emulation build.
__device__ __forceinline__ void ld4(const double *p, double &x0, double &x1, double &x2, double &x3) {
#ifdef SPZ_CPU_EMULATION
    x0 = p[0]; x1 = p[1]; x2 = p[2]; x3 = p[3];
#else
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3) : "l"(p));
#endif
}
__device__ __forceinline__ void st4(double *p, double x0, double x1, double x2, double x3) {
#ifdef SPZ_CPU_EMULATION
    p[0] = x0; p[1] = x1; p[2] = x2; p[3] = x3;
#else
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
#endif
}

// Butterfly over register bit RPOS.  CTRL: bit k0 of km says whether pair (k0, k0 | 1 << RPOS) is updated.
template <int KIND, int RPOS, bool CTRL>
__device__ __forceinline__ void butterfly2(double (&ar)[16], double (&ai)[16], const double *__restrict__ s, unsigned km) {
#pragma unroll
    for (int k0 = 0; k0 < 16; ++k0) {
        if (k0 & (1 << RPOS)) continue;
        const int k1 = k0 | (1 << RPOS);
        if (km & (1u << k0)) pair_update<KIND>(s, ar[k0], ai[k0], ar[k1], ai[k1]);
    }
}

// Apply accumulator (fr, fi) to the amplitudes whose register bit RPOS is set, then reset it.
template <int RPOS>
__device__ __forceinline__ void apply_bit_factor(double (&ar)[16], double (&ai)[16], double &fr, double &fi) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k & (1 << RPOS)) cmul2(ar[k], ai[k], fr, fi);
    fr = 1.0; fi = 0.0;
}

__device__ __forceinline__ void run_class2(const double2 *__restrict__ gfac, const unsigned *__restrict__ gthr, int cnt,
                                           unsigned tj, double &Fr, double &Fi) {
    if (cnt == 0) return;
    double ar_ = 1.0, ai_ = 0.0, br_ = 1.0, bi_ = 0.0;
    int i = 0;
    for (; i + 2 <= cnt; i += 2) {
        const unsigned t0 = gthr[i] & 0xffffu, t1 = gthr[i + 1] & 0xffffu;
        const double2 f0 = gfac[i], f1 = gfac[i + 1];
        if ((tj & t0) == t0) cmul2(ar_, ai_, f0.x, f0.y);
        if ((tj & t1) == t1) cmul2(br_, bi_, f1.x, f1.y);
    }
    if (i < cnt) {
        const unsigned t0 = gthr[i] & 0xffffu;
        const double2 f0 = gfac[i];
        if ((tj & t0) == t0) cmul2(ar_, ai_, f0.x, f0.y);
    }
    cmul2(ar_, ai_, br_, bi_);
    cmul2(Fr, Fi, ar_, ai_);
}

template <bool EXACT, bool CTRL>
__global__ void __launch_bounds__(kThreads2, 2) k_tile2(const Tile2Args a) {
    SPZ_DYN_SMEM(double, smem);
    constexpr unsigned tile_len = 1u << kT2;
    constexpr unsigned nthr = kThreads2;
    const int L = a.L;
    double *sre = smem;
    double *sim = smem + tile_len;
    double2 *gfac = reinterpret_cast<double2 *>(smem + 2 * tile_len);
    unsigned *gthr = reinterpret_cast<unsigned *>(gfac + a.n_groups);
    TileInstr *sprog = reinterpret_cast<TileInstr *>(reinterpret_cast<char *>(smem) + a.prog_off);
    __shared__ unsigned long long seg_off[1 << kMaxHigh2];

    // absolute index of the tile's first amplitude: CTA id bits go to the non-tile positions.  Recomputed for the
    // write-back instead of being kept alive through the interpreter loop (the loop reads per-instruction flags instead).
    auto tile_base = [&]() -> unsigned long long {
        unsigned long long b = (unsigned long long)(blockIdx.x + a.tile_offset) << L;
#pragma unroll
        for (int k = 0; k < kMaxHigh2; ++k)
            if (k < a.n_high) b = insert_zero(b, a.high[k]);
        return b;
    };
    __shared__ unsigned char iflag[kMaxInstr2]; // bit 0: skip (an outer control is 0 for this tile); bit 1: outer target bit is 1
    const int n_seg = 1 << a.n_high;
    for (int sgi = threadIdx.x; sgi < n_seg; sgi += nthr) {
        unsigned long long off = 0;
#pragma unroll
        for (int k = 0; k < kMaxHigh2; ++k)
            if (k < a.n_high && ((sgi >> k) & 1)) off |= 1ull << a.high[k];
        seg_off[sgi] = off;
    }
    __syncthreads(); // seg_off is needed to address the tile

    const unsigned seg_mask = (1u << L) - 1u;
    constexpr unsigned n_vec = tile_len >> 1;

    double ar[16], ai[16];
    double f0r = 1.0, f0i = 0.0, f1r = 1.0, f1i = 0.0, f2r = 1.0, f2i = 0.0, f3r = 1.0, f3i = 0.0, f4r = 1.0, f4i = 0.0;
    unsigned dirty = 0;
    unsigned tj = 0;
    unsigned stj = 0, sw0 = 1, sw1 = 2, sw2 = 4, sw3 = 8;
    auto saddr = [&](int k) -> unsigned {
        return stj ^ ((k & 1) ? sw0 : 0u) ^ ((k & 2) ? sw1 : 0u) ^ ((k & 4) ? sw2 : 0u) ^ ((k & 8) ? sw3 : 0u);
    };
    auto set_layout = [&](int r0, int r1, int r2, int r3) {
        tj = (unsigned)insert_zero(insert_zero(insert_zero(insert_zero(threadIdx.x, r0), r1), r2), r3);
        stj = swz2(tj); sw0 = swz2(1u << r0); sw1 = swz2(1u << r1); sw2 = swz2(1u << r2); sw3 = swz2(1u << r3);
    };
    // absolute offset of tile bit b (a power of two: low bits map to themselves, the others to high[b - L])
    auto bit_off = [&](unsigned b) -> unsigned long long { return b < (unsigned)L ? (1ull << b) : seg_off[1u << (b - L)]; };

    // ---- issue the tile's global loads FIRST, under the first register layout (passed in the kernel arguments), so
    // that staging the program and reducing the phase groups below run while the amplitudes are in flight ----
    set_layout((int)(a.first_rbits & 255u), (int)((a.first_rbits >> 8) & 255u), (int)((a.first_rbits >> 16) & 255u), (int)(a.first_rbits >> 24));
    {
    const unsigned long long base = tile_base();
    if (a.first_direct) {
        const unsigned rpack = a.first_rbits;
        const unsigned long long g0 = base + seg_off[tj >> L] + (tj & seg_mask);
        const unsigned long long o0 = bit_off(rpack & 255u), o1 = bit_off((rpack >> 8) & 255u);
        const unsigned long long o2 = bit_off((rpack >> 16) & 255u), o3 = bit_off(rpack >> 24);
        if (a.first_direct == 4) { // amplitudes k = 4q .. 4q+3 are consecutive in memory
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned long long g = g0 + ((q & 1) ? o2 : 0ull) + ((q & 2) ? o3 : 0ull);
                ld4(a.re + g, ar[4 * q], ar[4 * q + 1], ar[4 * q + 2], ar[4 * q + 3]);
                ld4(a.im + g, ai[4 * q], ai[4 * q + 1], ai[4 * q + 2], ai[4 * q + 3]);
            }
        } else if (a.first_direct == 2) { // pairs k = 2p, 2p+1
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned long long g = g0 + ((q & 1) ? o1 : 0ull) + ((q & 2) ? o2 : 0ull) + ((q & 4) ? o3 : 0ull);
                const double2 r = *reinterpret_cast<const double2 *>(a.re + g);
                const double2 m = *reinterpret_cast<const double2 *>(a.im + g);
                ar[2 * q] = r.x; ar[2 * q + 1] = r.y;
                ai[2 * q] = m.x; ai[2 * q + 1] = m.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned long long g = g0 + ((k & 1) ? o0 : 0ull) + ((k & 2) ? o1 : 0ull) + ((k & 4) ? o2 : 0ull) + ((k & 8) ? o3 : 0ull);
                ar[k] = a.re[g];
                ai[k] = a.im[g];
            }
        }
    } else {
        // coalesced 128-bit loads, parked in the amplitude registers (pairs 2*it, 2*it+1) until the staging stores below
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const unsigned v = threadIdx.x + it * nthr;
            const unsigned j = v << 1;
            const unsigned long long g = base + seg_off[j >> L] + (j & seg_mask);
            const double2 r = *reinterpret_cast<const double2 *>(a.re + g);
            const double2 m = *reinterpret_cast<const double2 *>(a.im + g);
            ar[2 * it] = r.x; ar[2 * it + 1] = r.y;
            ai[2 * it] = m.x; ai[2 * it + 1] = m.y;
        }
    }

    // ---- stage the program; per-instruction tile flags; reduce every phase group over this tile's outer bits ----
    if ((int)threadIdx.x < a.n_instr) {
        const TileInstr *gi = a.prog + threadIdx.x; // straight from global memory: independent of the staging copy below
        const unsigned long long ocm = gi->outer_cmask;
        unsigned f = ((base & ocm) != ocm) ? 1u : 0u;
        if (gi->op == TI_DIAG && gi->t_where == 0) {
            const bool hi = gi->const_hi ? (gi->const_hi == 2u) : (bool)((base >> gi->outer_target) & 1ull);
            f |= hi ? 2u : 0u;
        }
        iflag[threadIdx.x] = (unsigned char)f;
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.prog);
        uint4 *dst = reinterpret_cast<uint4 *>(sprog);
        const int n16 = a.n_instr * (int)(sizeof(TileInstr) / sizeof(uint4));
        for (int i = threadIdx.x; i < n16; i += nthr) dst[i] = src[i];
    }
    // Phase factors: every group's product over the terms whose outer bits are set in this tile.  Term-parallel when the
    // terms fit the (still unused) tile region as scratch: one L2 round trip for all terms at once instead of one per term
    // in series per group -- a QFT pass has groups of ~20 terms, i.e. microseconds at the head of every tile otherwise.
    constexpr int kScratchTerms = 2048; // 32 KB of factors + 2 KB of hit flags inside the 64 KB tile region
    const bool term_parallel = a.n_terms <= kScratchTerms;
    double2 *scr = reinterpret_cast<double2 *>(sre);
    unsigned char *scr_hit = reinterpret_cast<unsigned char *>(scr + kScratchTerms);
    if (term_parallel) {
        for (int t = threadIdx.x; t < a.n_terms; t += nthr) {
            const TileTerm tm = a.terms[t];
            const bool hit = (base & tm.outer) == tm.outer;
            scr[t] = hit ? make_double2(tm.fr, tm.fi) : make_double2(1.0, 0.0);
            scr_hit[t] = hit ? 1 : 0;
        }
        __syncthreads();
    }
    for (int g = threadIdx.x; g < a.n_groups; g += nthr) {
        const TileGroup gd = a.groups[g];
        double fr = 1.0, fi = 0.0;
        bool any = false;
        if (term_parallel) {
            for (int i = 0; i < gd.count; ++i) {
                if (!scr_hit[gd.first + i]) continue;
                const double2 f = scr[gd.first + i];
                cmul2(fr, fi, f.x, f.y);
                any = true;
            }
        } else {
            for (int i = 0; i < gd.count; ++i) {
                const TileTerm t = a.terms[gd.first + i];
                if ((base & t.outer) == t.outer) { cmul2(fr, fi, t.fr, t.fi); any = true; }
            }
        }
        gfac[g] = make_double2(fr, fi);
        gthr[g] = (any ? gd.thr : 0xffffu) | (gd.m << 16);
    }
    if (term_parallel && !a.first_direct) __syncthreads(); // the staged tile is about to overwrite the scratch
    } // base

    // expand all five accumulators into the 16 per-amplitude factors (layout changes and the end of the program)
    auto flush_all = [&]() {
        if (!dirty) return;
        if (__popc(dirty) <= 2) {
            if (dirty & 1u) {
#pragma unroll
                for (int k = 0; k < 16; ++k) cmul2(ar[k], ai[k], f0r, f0i);
            }
            if (dirty & 2u) apply_bit_factor<0>(ar, ai, f1r, f1i);
            if (dirty & 4u) apply_bit_factor<1>(ar, ai, f2r, f2i);
            if (dirty & 8u) apply_bit_factor<2>(ar, ai, f3r, f3i);
            if (dirty & 16u) apply_bit_factor<3>(ar, ai, f4r, f4i);
        } else {
#pragma unroll
            for (int b3 = 0; b3 < 2; ++b3) {
                double g3r = f0r, g3i = f0i;
                if (b3) cmul2(g3r, g3i, f4r, f4i);
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    double g2r = g3r, g2i = g3i;
                    if (b2) cmul2(g2r, g2i, f3r, f3i);
#pragma unroll
                    for (int b1 = 0; b1 < 2; ++b1) {
                        double g1r = g2r, g1i = g2i;
                        if (b1) cmul2(g1r, g1i, f2r, f2i);
#pragma unroll
                        for (int b0 = 0; b0 < 2; ++b0) {
                            double gr = g1r, gi = g1i;
                            if (b0) cmul2(gr, gi, f1r, f1i);
                            cmul2(ar[b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)], ai[b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)], gr, gi);
                        }
                    }
                }
            }
        }
        f0r = f1r = f2r = f3r = f4r = 1.0;
        f0i = f1i = f2i = f3i = f4i = 0.0;
        dirty = 0;
    };
    auto store_regs = [&]() {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned s = saddr(k);
            sre[s] = ar[k]; sim[s] = ai[k];
        }
    };
    auto load_regs = [&]() {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned s = saddr(k);
            ar[k] = sre[s]; ai[k] = sim[s];
        }
    };

    if (!a.first_direct) {
        // swizzled placement in shared memory, then re-read under the register layout
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const unsigned v = threadIdx.x + it * nthr;
            const unsigned s0 = swz2(v << 1);
            const bool flip = s0 & 1u;
            *reinterpret_cast<double2 *>(sre + (s0 & ~1u)) = flip ? make_double2(ar[2 * it + 1], ar[2 * it]) : make_double2(ar[2 * it], ar[2 * it + 1]);
            *reinterpret_cast<double2 *>(sim + (s0 & ~1u)) = flip ? make_double2(ai[2 * it + 1], ai[2 * it]) : make_double2(ai[2 * it], ai[2 * it + 1]);
        }
    }
    __syncthreads(); // program and group tables staged (and, on the staged path, the tile)
    if (!a.first_direct) {
        load_regs();
        // No barrier here (k_tile has one): the next shared-memory access of this thread is store_regs() under the SAME
        // layout, i.e. to exactly the cells it has just read, which no other thread touches in between.  Checked with
        // ThreadSanitizer on the CPU emulation (tests/test_tile2_cpu_emulation.py).
    }

    for (int pc = 1; pc < a.n_instr; ++pc) {
        const TileInstr &ins = sprog[pc];
        const int op = ins.op;
        if (op == TI_LAYOUT) {
            flush_all();
            store_regs();
            __syncthreads();
            set_layout(ins.rbit[0], ins.rbit[1], ins.rbit[2], ins.rbit[3]);
            load_regs(); // one barrier per layout change is enough, see above
            continue;
        }
        if (op == TI_RUN) {
            int g = ins.rpos;
            const int c0 = ins.rbit[0], c1 = ins.rbit[1], c2 = ins.rbit[2], c3 = ins.rbit[3];
            const int c4 = (int)ins.reg_cmask, c5 = (int)ins.thr_cmask;
            run_class2(gfac + g, gthr + g, c0, tj, f0r, f0i); g += c0;
            run_class2(gfac + g, gthr + g, c1, tj, f1r, f1i); g += c1;
            run_class2(gfac + g, gthr + g, c2, tj, f2r, f2i); g += c2;
            run_class2(gfac + g, gthr + g, c3, tj, f3r, f3i); g += c3;
            run_class2(gfac + g, gthr + g, c4, tj, f4r, f4i); g += c4;
            dirty |= (c0 ? 1u : 0u) | (c1 ? 2u : 0u) | (c2 ? 4u : 0u) | (c3 ? 8u : 0u) | (c4 ? 16u : 0u);
            for (int i = 0; i < c5; ++i) {
                const unsigned packed = gthr[g + i], thr = packed & 0xffffu, m = packed >> 16;
                const double2 f = gfac[g + i];
                if ((tj & thr) != thr) continue;
#define SPZ_M4(A, B, C, D) cmul2(ar[A], ai[A], f.x, f.y); cmul2(ar[B], ai[B], f.x, f.y); cmul2(ar[C], ai[C], f.x, f.y); cmul2(ar[D], ai[D], f.x, f.y)
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
                        if (((unsigned)k & m) == m) cmul2(ar[k], ai[k], f.x, f.y);
                    break;
                }
#undef SPZ_M4
            }
            continue;
        }
        const unsigned fl = iflag[pc];
        if (fl & 1u) continue; // an outer control is 0 for this whole tile
        if (op == TI_GATE) {
            // CTRL = false: no in-tile controls anywhere in the program, so the pair mask is the same for every thread
            // (0xffff, read from shared memory with a uniform address): the per-pair guards become uniform branches,
            // which keeps the register pressure of the guarded form without its convergence barriers.
            unsigned km = ins.t_mask;
            if constexpr (CTRL) km = ((tj & ins.thr_cmask) == ins.thr_cmask) ? km : 0u;
            const int kind = ins.kind;
            // Only the accumulator of the target's own register bit separates the two members of a pair; the others
            // scale both by the same factor and stay pending (see the header of this file).
#define SPZ_BFLY(R, FR, FI, BIT)                                                                   \
    {                                                                                              \
        if (dirty & BIT) { apply_bit_factor<R>(ar, ai, FR, FI); dirty &= ~BIT; }                   \
        switch (kind) {                                                                            \
        case SPZ_GATE_H: butterfly2<SPZ_GATE_H, R, CTRL>(ar, ai, ins.s, km); break;                \
        case SPZ_GATE_X: butterfly2<SPZ_GATE_X, R, CTRL>(ar, ai, ins.s, km); break;                \
        case SPZ_GATE_Y: butterfly2<SPZ_GATE_Y, R, CTRL>(ar, ai, ins.s, km); break;                \
        case SPZ_GATE_RX: butterfly2<SPZ_GATE_RX, R, CTRL>(ar, ai, ins.s, km); break;              \
        case SPZ_GATE_RY: butterfly2<SPZ_GATE_RY, R, CTRL>(ar, ai, ins.s, km); break;              \
        case SPZ_GATE_U: butterfly2<SPZ_GATE_U, R, CTRL>(ar, ai, ins.s, km); break;                \
        default: break;                                                                            \
        }                                                                                          \
    }
            switch (ins.rpos) {
            case 0: SPZ_BFLY(0, f1r, f1i, 2u) break;
            case 1: SPZ_BFLY(1, f2r, f2i, 4u) break;
            case 2: SPZ_BFLY(2, f3r, f3i, 8u) break;
            default: SPZ_BFLY(3, f4r, f4i, 16u) break;
            }
#undef SPZ_BFLY
            continue;
        }
        // ---- TI_DIAG (exact mode only: merged mode folds diagonal gates into TI_RUN groups) ----
        if constexpr (EXACT) {
            const int kind = ins.kind;
            const int tw = ins.t_where;
            bool outer_hi = false;
            if (tw == 0) outer_hi = (fl & 2u) != 0u;
            if (tw == 0 && !outer_hi && kind != SPZ_GATE_RZ) continue;
            const bool ok = (tj & ins.thr_cmask) == ins.thr_cmask;
            const bool thr_hi = tw == 1 ? ((tj & ins.t_mask) != 0) : outer_hi;
            const unsigned creg = ins.reg_cmask;
            const unsigned tmask = ins.t_mask;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (ok && ((unsigned)k & creg) == creg) {
                    const bool hi = tw == 2 ? (((unsigned)k & tmask) != 0) : thr_hi;
                    diag_update_rt(kind, ins.s, hi, ar[k], ai[k]);
                }
            }
        }
    }

    flush_all();
    const unsigned long long base = tile_base();
    if (a.last_direct) {
        // registers -> global under the final layout (tj is current; its register bits come with the arguments)
        const unsigned rpack = a.last_rbits;
        const unsigned long long g0 = base + seg_off[tj >> L] + (tj & seg_mask);
        const unsigned long long o0 = bit_off(rpack & 255u), o1 = bit_off((rpack >> 8) & 255u);
        const unsigned long long o2 = bit_off((rpack >> 16) & 255u), o3 = bit_off(rpack >> 24);
        if (a.last_direct == 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned long long g = g0 + ((q & 1) ? o2 : 0ull) + ((q & 2) ? o3 : 0ull);
                st4(a.re + g, ar[4 * q], ar[4 * q + 1], ar[4 * q + 2], ar[4 * q + 3]);
                st4(a.im + g, ai[4 * q], ai[4 * q + 1], ai[4 * q + 2], ai[4 * q + 3]);
            }
        } else if (a.last_direct == 2) {
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
        return;
    }
    store_regs();
    __syncthreads();
    for (unsigned v = threadIdx.x; v < n_vec; v += nthr) {
        const unsigned j = v << 1;
        const unsigned long long g = base + seg_off[j >> L] + (j & seg_mask);
        const unsigned s0 = swz2(j);
        const bool flip = s0 & 1u;
        const double2 r = *reinterpret_cast<const double2 *>(sre + (s0 & ~1u));
        const double2 m = *reinterpret_cast<const double2 *>(sim + (s0 & ~1u));
        *reinterpret_cast<double2 *>(a.re + g) = flip ? make_double2(r.y, r.x) : r;
        *reinterpret_cast<double2 *>(a.im + g) = flip ? make_double2(m.y, m.x) : m;
    }
}

size_t tile2_smem_bytes(int n_instr, int n_groups, unsigned *prog_off) {
    size_t smem = sizeof(double) * 2u * ((size_t)1 << kT2) + (size_t)n_groups * (sizeof(double2) + sizeof(unsigned));
    smem = (smem + 15) & ~(size_t)15;
    if (prog_off) *prog_off = (unsigned)smem;
    return smem + sizeof(TileInstr) * (size_t)n_instr;
}

// Kernel arguments, dynamic shared-memory size and instantiation for tiles [first, ...) of one pass.  Pure host code,
// shared by the launcher below and by the CPU emulation harness.
Tile2Args tile2_make_args(double *re, double *im, const TilePlan &plan, const TileInstr *h_prog, int n_instr, const TileInstr *d_prog,
                          const TileGroup *d_groups, int n_groups, const TileTerm *d_terms, int n_terms, unsigned first,
                          int direct_level, size_t *smem_bytes, bool *ctrl) {
    Tile2Args a{};
    a.re = re; a.im = im;
    a.prog = d_prog; a.groups = d_groups; a.terms = d_terms;
    a.n_instr = n_instr; a.n_groups = n_groups; a.n_terms = n_terms;
    a.tile_offset = first;
    a.L = plan.low_bits; a.n_high = plan.n_high;
    for (int k = 0; k < plan.n_high; ++k) a.high[k] = plan.high[k];
    *smem_bytes = tile2_smem_bytes(n_instr, n_groups, &a.prog_off);
    *ctrl = false;
    int last_layout = 0;
    for (int i = 0; i < n_instr; ++i) {
        if (h_prog[i].op == TI_LAYOUT) last_layout = i;
        if (h_prog[i].op == TI_GATE && (h_prog[i].thr_cmask || h_prog[i].reg_cmask)) *ctrl = true;
    }
    // How is a layout transferred?  (rbit is ascending.)  The kernel is correct for every choice; only coalescing differs,
    // which is what SPZ_TILE_V2_DIRECT lets a GPU run explore:
    //   level 0: always staged through shared memory (as k_tile).
    //   level 1 (default): direct, one amplitude per access, when all register bits are >= 4 -- 16 consecutive lanes own one
    //            128-byte line, the best case; staged otherwise.
    //   level 2: additionally direct with 256-bit accesses when register bits 0,1 are tile bits 0,1 (each lane owns full
    //            32-byte sectors, lanes 128 bytes or more apart).
    //   level 3: always direct: 256-bit if possible, else 128-bit when register bit 0 is tile bit 0, else one amplitude.
    auto mode = [&](const TileInstr &l) -> int {
        const bool quad = l.rbit[0] == 0 && l.rbit[1] == 1, pair = l.rbit[0] == 0;
        if (direct_level <= 0) return 0;
        if (l.rbit[0] >= 4) return 1;
        if (direct_level == 1) return 0;
        if (quad) return 4;
        if (direct_level == 2) return 0;
        return pair ? 2 : 1;
    };
    a.first_direct = mode(h_prog[0]);
    a.last_direct = mode(h_prog[last_layout]);
    auto pack = [](const TileInstr &l) {
        return (unsigned)l.rbit[0] | ((unsigned)l.rbit[1] << 8) | ((unsigned)l.rbit[2] << 16) | ((unsigned)l.rbit[3] << 24);
    };
    a.first_rbits = pack(h_prog[0]);
    a.last_rbits = pack(h_prog[last_layout]);
    return a;
}

bool tile2_shape_ok(int n_qubits, const TilePlan &plan, const TileInstr *prog, int n_instr, int n_groups) {
    if (plan.tile_bits != kT2 || plan.n_high > kMaxHigh2 || plan.low_bits < 4 || n_qubits < kT2) return false;
    if (n_instr < 1 || n_instr > kMaxInstr2 || n_groups > kMaxTileGroups) return false;
    if (tile2_smem_bytes(n_instr, n_groups, nullptr) > kSmemBudget2) return false;
    return prog[0].op == TI_LAYOUT;
}

} // namespace

#ifndef SPZ_CPU_EMULATION
// 0 = off (default), 1 = on.  Read per call so that tests can toggle it inside one process.
bool tile2_enabled() {
    const char *e = std::getenv("SPZ_TILE_V2");
    return e && e[0] == '1';
}

static int tile2_prepare() {
    const int max_smem = (int)kSmemBudget2;
    SPZ_CUDA(cudaFuncSetAttribute(k_tile2<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SPZ_CUDA(cudaFuncSetAttribute(k_tile2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SPZ_CUDA(cudaFuncSetAttribute(k_tile2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    return SPZ_OK;
}

// Can this pass run on k_tile2?  (full 12-bit tile, program short enough for shared memory, starts with a LAYOUT)
bool tile2_eligible(const spz_state *st, const TilePlan &plan, const TileInstr *prog, int n_instr, int n_groups) {
    return tile2_shape_ok(st->n, plan, prog, n_instr, n_groups);
}

// d_prog / d_groups / d_terms: the device copies staged by launch_tile_program (kernels_tile.cu), which also owns the
// ring buffer and the split launch after an overlapped exchange; this function only picks the instantiation and launches
// tiles [first, first + count).
int launch_tile2(spz_state *st, const TilePlan &plan, const TileInstr *h_prog, int n_instr, const TileInstr *d_prog,
                 const TileGroup *d_groups, int n_groups, const TileTerm *d_terms, int n_terms, bool exact, unsigned first,
                 unsigned count) {
    static bool prepared[64] = {false};
    if (st->device >= 0 && st->device < 64 && !prepared[st->device]) {
        SPZ_TRY(tile2_prepare());
        prepared[st->device] = true;
    }
    size_t smem = 0;
    bool ctrl = false;
    int direct_level = 1; // SPZ_TILE_V2_DIRECT = 0..3, see tile2_make_args
    if (const char *e = std::getenv("SPZ_TILE_V2_DIRECT")) if (e[0] >= '0' && e[0] <= '3') direct_level = e[0] - '0';
    const Tile2Args a = tile2_make_args(st->re, st->im, plan, h_prog, n_instr, d_prog, d_groups, n_groups, d_terms, n_terms, first,
                                        direct_level, &smem, &ctrl);
    if (exact) k_tile2<true, true><<<count, kThreads2, smem, st->stream>>>(a);
    else if (ctrl) k_tile2<false, true><<<count, kThreads2, smem, st->stream>>>(a);
    else k_tile2<false, false><<<count, kThreads2, smem, st->stream>>>(a);
    return SPZ_OK;
}
#endif // !SPZ_CPU_EMULATION

} // namespace spz
