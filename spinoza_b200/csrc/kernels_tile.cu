// kernels_tile.cu -- fused execution: a run of gates applied tile-by-tile on chip so that one HBM pass
// (32 * 2^n bytes) serves the whole run.  This is the engine behind spz_execute(SPZ_EXEC_FUSE); the reference
// applies one gate per full pass (circuit.rs:553-599, "no fusion").
//
// Tile.  The 2^T amplitudes (T = 12: 64 KB of re+im) that share all index bits OUTSIDE the tile bit set
//   { 0 .. L-1 }  U  { high[0] < high[1] < ... }            (T = L + n_high)
// i.e. 2^n_high contiguous segments of 2^L amplitudes.  A gate whose target is a tile qubit can be applied inside
// the tile; controls may be anywhere (outer controls are constant per tile: the CTA skips the gate or applies it
// unconditionally); diagonal gates (Z, P, RZ) may even target an outer qubit (constant factor per tile).
//
// Execution model.  One CTA per tile, 2^(T-4) threads.  Shared memory is the tile's home (XOR-swizzled so every
// register layout below is bank-conflict free); each thread keeps 16 amplitudes in REGISTERS: the 4 tile bits of
// the current "register layout" R index the thread's amplitudes k = 0..15, the other T-4 tile bits come from the
// thread id.  The host compiles the op run into a micro-program (TileInstr, engine.h):
//   LAYOUT R      write registers back, barrier, re-read with 4 other bits register-resident
//   GATE          non-diagonal gate on a register bit: a butterfly among the thread's own 16 amplitudes, with the
//                 same per-pair arithmetic as the unfused kernels (gate_math.cuh) -> bit-identical results
//   DIAG          diagonal gate.  Exact mode: applied per amplitude with the reference arithmetic (bit-identical).
//                 Merged mode (default): a run of diagonal gates is folded into per-thread phase accumulators
//                 F0 (all 16 amplitudes) and F1..F4 (amplitudes whose register bit i is 1); only gates whose
//                 support has >= 2 register bits touch amplitudes directly.  At the end of the run the 16
//                 per-amplitude factors are expanded from the 5 accumulators (15 complex multiplies) and
//                 applied once.  Terms of a run that share their thread/register masks differ only in which OUTER
//                 bits they need, so their product is a per-tile constant: it is computed once per CTA
//                 (cooperatively, before the first barrier) and read from shared memory.  A QFT pass with ~280
//                 controlled-phase gates then costs ~80 complex multiplies per THREAD instead of 16 per gate.  Rounding differs from gate-by-gate application by a few
//                 ulp per run (documented in DESIGN.md; tests hold it to 1e-12 absolute).
// No barrier is needed between gates that act on register bits; barriers only surround LAYOUT changes.
#include <algorithm>
#include <vector>

#include "gate_math.cuh"

// tests/emu/ compiles the kernel body with g++ (SPZ_CPU_EMULATION: one OS thread per CUDA thread) to check it on the CPU;
// only the name of the dynamic shared-memory window differs between the two builds.
#ifdef SPZ_CPU_EMULATION
#define SPZ_TILE_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(spz_emu::dyn_smem)
#else
#define SPZ_TILE_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

namespace spz {

constexpr int kMaxTileBits = 12;
constexpr int kMaxHigh = 8;
constexpr int kRegBits = 4; // 16 amplitudes per thread
constexpr int kMaxSmemInstr = 96; // programs up to this many instructions (12 KB) are staged in shared memory

int max_tile_bits() { return kMaxTileBits; }
int min_tile_bits() { return kRegBits; }

struct TileArgs {
    double *re;
    double *im;
    const TileInstr *prog;
    const TileGroup *groups;
    const TileTerm *terms;
    int n_instr;
    int n_groups;
    unsigned tile_offset; // first tile of this launch (a pass may be launched in two halves, see dist.cu)
    int prog_in_smem; // 1: the program is staged in shared memory (it fits the budget)
    unsigned prog_off; // byte offset of the staged program inside dynamic shared memory
    int T, L, n_high;
    int high[kMaxHigh];
};

__device__ __forceinline__ unsigned swz(unsigned j) { return j ^ (((j >> 4) ^ (j >> 8)) & 15u); }

__device__ __forceinline__ void cmul(double &xr, double &xi, double fr, double fi) {
    const double nr = xr * fr - xi * fi;
    const double ni = xr * fi + xi * fr;
    xr = nr; xi = ni;
}

// Butterfly of a non-diagonal gate over register bit RPOS: pairs (k, k | 1 << RPOS).
template <int KIND, int RPOS>
__device__ __forceinline__ void butterfly(double (&ar)[16], double (&ai)[16], const double *__restrict__ s, unsigned km) {
    // km: bit k0 set <=> this thread applies the gate to pair (k0, k0 | 1 << RPOS)  (controls already folded in)
#pragma unroll
    for (int k0 = 0; k0 < 16; ++k0) {
        if (k0 & (1 << RPOS)) continue;
        const int k1 = k0 | (1 << RPOS);
        if (km & (1u << k0)) pair_update<KIND>(s, ar[k0], ai[k0], ar[k1], ai[k1]);
    }
}

template <int KIND>
__device__ __forceinline__ void butterfly_pos(int rpos, double (&ar)[16], double (&ai)[16], const double *__restrict__ s,
                                              unsigned km) {
    switch (rpos) {
    case 0: butterfly<KIND, 0>(ar, ai, s, km); break;
    case 1: butterfly<KIND, 1>(ar, ai, s, km); break;
    case 2: butterfly<KIND, 2>(ar, ai, s, km); break;
    default: butterfly<KIND, 3>(ar, ai, s, km); break;
    }
}

// One class of a merged diagonal run: every group multiplies the same accumulator.  Group factors (already
// reduced over the tile's outer bits) and thread masks come from shared memory with warp-uniform addresses; two
// groups are fetched per iteration and folded into two independent partial products.
__device__ __forceinline__ void run_class(const double2 *__restrict__ gfac, const unsigned *__restrict__ gthr, int cnt,
                                          unsigned tj, double &Fr, double &Fi) {
    if (cnt == 0) return;
    double ar_ = 1.0, ai_ = 0.0, br_ = 1.0, bi_ = 0.0;
    int i = 0;
    for (; i + 2 <= cnt; i += 2) {
        const unsigned t0 = gthr[i] & 0xffffu, t1 = gthr[i + 1] & 0xffffu;
        const double2 f0 = gfac[i], f1 = gfac[i + 1];
        if ((tj & t0) == t0) cmul(ar_, ai_, f0.x, f0.y);
        if ((tj & t1) == t1) cmul(br_, bi_, f1.x, f1.y);
    }
    if (i < cnt) {
        const unsigned t0 = gthr[i] & 0xffffu;
        const double2 f0 = gfac[i];
        if ((tj & t0) == t0) cmul(ar_, ai_, f0.x, f0.y);
    }
    cmul(ar_, ai_, br_, bi_);
    cmul(Fr, Fi, ar_, ai_);
}

template <bool EXACT>
__global__ void __launch_bounds__(256, 2) k_tile(const TileArgs a) {
    SPZ_TILE_DYN_SMEM(double, smem);
    const int T = a.T, L = a.L;
    const unsigned tile_len = 1u << T;
    const unsigned nthr = blockDim.x;
    double *sre = smem;
    double *sim = smem + tile_len;
    double2 *gfac = reinterpret_cast<double2 *>(smem + 2 * tile_len); // per-group factor for THIS tile
    unsigned *gthr = reinterpret_cast<unsigned *>(gfac + a.n_groups);  // per-group thread mask (all ones = never)
    // the micro-program itself, staged once per CTA so that decoding reads shared memory with uniform addresses
    TileInstr *sprog = reinterpret_cast<TileInstr *>(reinterpret_cast<char *>(smem) + a.prog_off);
    __shared__ unsigned long long seg_off[1 << kMaxHigh];

    // absolute index of the tile's first amplitude: CTA id bits go to the non-tile positions
    unsigned long long base = (unsigned long long)(blockIdx.x + a.tile_offset) << L;
#pragma unroll
    for (int k = 0; k < kMaxHigh; ++k) // compile-time indices keep the kernel parameters in the constant bank
        if (k < a.n_high) base = insert_zero(base, a.high[k]);
    const int n_seg = 1 << a.n_high;
    for (int sgi = threadIdx.x; sgi < n_seg; sgi += nthr) {
        unsigned long long off = 0;
#pragma unroll
        for (int k = 0; k < kMaxHigh; ++k)
            if (k < a.n_high && ((sgi >> k) & 1)) off |= 1ull << a.high[k];
        seg_off[sgi] = off;
    }
    if (a.prog_in_smem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.prog);
        uint4 *dst = reinterpret_cast<uint4 *>(sprog);
        const int n16 = a.n_instr * (int)(sizeof(TileInstr) / sizeof(uint4));
        for (int i = threadIdx.x; i < n16; i += nthr) dst[i] = src[i];
    }
    // per-tile reduction of every group of every diagonal run: product of its terms whose outer bits are set
    for (int g = threadIdx.x; g < a.n_groups; g += nthr) {
        const TileGroup gd = a.groups[g];
        double fr = 1.0, fi = 0.0;
        bool any = false;
        for (int i = 0; i < gd.count; ++i) {
            const TileTerm t = a.terms[gd.first + i];
            if ((base & t.outer) == t.outer) { cmul(fr, fi, t.fr, t.fi); any = true; }
        }
        gfac[g] = make_double2(fr, fi);
        // low 16 bits: thread mask (all ones = nothing applies to this tile, no thread matches); high bits: register mask m
        gthr[g] = (any ? gd.thr : 0xffffu) | (gd.m << 16);
    }
    __syncthreads();

    // ---- global -> shared (coalesced 128-bit loads; swizzled placement) ----
    // All of a thread's loads (8 per array at T = 12) are issued before the first shared store, staged in the
    // registers (free at this point), so the HBM latency is paid once per tile, not once per vector.
    const unsigned n_vec = tile_len >> 1;
    const unsigned seg_mask = (1u << L) - 1u;
    {
        double2 tr[8], tm[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const unsigned v = threadIdx.x + it * nthr;
            tr[it] = make_double2(0.0, 0.0); tm[it] = make_double2(0.0, 0.0);
            if (v < n_vec) {
                const unsigned j = v << 1;
                const unsigned long long g = base + seg_off[j >> L] + (j & seg_mask);
                tr[it] = *reinterpret_cast<const double2 *>(a.re + g);
                tm[it] = *reinterpret_cast<const double2 *>(a.im + g);
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const unsigned v = threadIdx.x + it * nthr;
            if (v < n_vec) {
                const unsigned s0 = swz(v << 1); // swz(j + 1) == s0 ^ 1: the pair stays an aligned pair, possibly swapped
                const bool flip = s0 & 1u;
                *reinterpret_cast<double2 *>(sre + (s0 & ~1u)) = flip ? make_double2(tr[it].y, tr[it].x) : tr[it];
                *reinterpret_cast<double2 *>(sim + (s0 & ~1u)) = flip ? make_double2(tm[it].y, tm[it].x) : tm[it];
            }
        }
    }
    __syncthreads();

    double ar[16], ai[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { ar[k] = 0.0; ai[k] = 0.0; }
    // merged-mode phase accumulators: F0 multiplies all 16 amplitudes, Fi those whose register bit i-1 is set
    double f0r = 1.0, f0i = 0.0, f1r = 1.0, f1i = 0.0, f2r = 1.0, f2i = 0.0, f3r = 1.0, f3i = 0.0, f4r = 1.0, f4i = 0.0;
    unsigned dirty = 0; // bit c: accumulator of class c is not the identity
    bool have_regs = false;
    unsigned tj = 0;                       // this thread's tile index with the register bits cleared

    // swz() is linear over XOR, so the swizzled address of amplitude k is swz(tj) ^ (XOR of the swizzled register
    // bits selected by k): one XOR per access with warp-uniform operands.
    unsigned stj = 0, sw0 = 1, sw1 = 2, sw2 = 4, sw3 = 8;
    auto saddr = [&](int k) -> unsigned {
        return stj ^ ((k & 1) ? sw0 : 0u) ^ ((k & 2) ? sw1 : 0u) ^ ((k & 4) ? sw2 : 0u) ^ ((k & 8) ? sw3 : 0u);
    };

    auto flush_diag = [&]() {
        if (!dirty) return;
        if (__popc(dirty) <= 2) {
            // few accumulators: apply each to its amplitudes directly (16 or 8 complex multiplies each)
            if (dirty & 1u) {
#pragma unroll
                for (int k = 0; k < 16; ++k) cmul(ar[k], ai[k], f0r, f0i);
            }
            if (dirty & 2u) {
#pragma unroll
                for (int k = 0; k < 16; ++k) if (k & 1) cmul(ar[k], ai[k], f1r, f1i);
            }
            if (dirty & 4u) {
#pragma unroll
                for (int k = 0; k < 16; ++k) if (k & 2) cmul(ar[k], ai[k], f2r, f2i);
            }
            if (dirty & 8u) {
#pragma unroll
                for (int k = 0; k < 16; ++k) if (k & 4) cmul(ar[k], ai[k], f3r, f3i);
            }
            if (dirty & 16u) {
#pragma unroll
                for (int k = 0; k < 16; ++k) if (k & 8) cmul(ar[k], ai[k], f4r, f4i);
            }
        } else {
            // expand the 5 accumulators into the 16 per-amplitude factors, depth first (15 + 16 complex multiplies)
#pragma unroll
            for (int b3 = 0; b3 < 2; ++b3) {
                double g3r = f0r, g3i = f0i;
                if (b3) cmul(g3r, g3i, f4r, f4i);
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    double g2r = g3r, g2i = g3i;
                    if (b2) cmul(g2r, g2i, f3r, f3i);
#pragma unroll
                    for (int b1 = 0; b1 < 2; ++b1) {
                        double g1r = g2r, g1i = g2i;
                        if (b1) cmul(g1r, g1i, f2r, f2i);
#pragma unroll
                        for (int b0 = 0; b0 < 2; ++b0) {
                            double gr = g1r, gi = g1i;
                            if (b0) cmul(gr, gi, f1r, f1i);
                            cmul(ar[b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)], ai[b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)], gr, gi);
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

    for (int pc = 0; pc < a.n_instr; ++pc) {
        const TileInstr &ins = a.prog_in_smem ? sprog[pc] : a.prog[pc];
        const int op = ins.op;
        if (op == TI_LAYOUT) {
            if (have_regs) {
                flush_diag();
                store_regs();
                __syncthreads();
            }
            const int r0 = ins.rbit[0], r1 = ins.rbit[1], r2 = ins.rbit[2], r3 = ins.rbit[3];
            tj = (unsigned)insert_zero(insert_zero(insert_zero(insert_zero(threadIdx.x, r0), r1), r2), r3);
            stj = swz(tj); sw0 = swz(1u << r0); sw1 = swz(1u << r1); sw2 = swz(1u << r2); sw3 = swz(1u << r3);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned s = saddr(k);
                ar[k] = sre[s]; ai[k] = sim[s];
            }
            have_regs = true;
            __syncthreads(); // everyone has read before anyone's next store_regs
            continue;
        }
        if (op == TI_RUN) {
            int g = ins.rpos;
            const int c0 = ins.rbit[0], c1 = ins.rbit[1], c2 = ins.rbit[2], c3 = ins.rbit[3];
            const int c4 = (int)ins.reg_cmask, c5 = (int)ins.thr_cmask;
            run_class(gfac + g, gthr + g, c0, tj, f0r, f0i); g += c0;
            run_class(gfac + g, gthr + g, c1, tj, f1r, f1i); g += c1;
            run_class(gfac + g, gthr + g, c2, tj, f2r, f2i); g += c2;
            run_class(gfac + g, gthr + g, c3, tj, f3r, f3i); g += c3;
            run_class(gfac + g, gthr + g, c4, tj, f4r, f4i); g += c4;
            dirty |= (c0 ? 1u : 0u) | (c1 ? 2u : 0u) | (c2 ? 4u : 0u) | (c3 ? 8u : 0u) | (c4 ? 16u : 0u);
            for (int i = 0; i < c5; ++i) { // support with >= 2 register bits: touch the amplitudes directly
                const unsigned packed = gthr[g + i], thr = packed & 0xffffu, m = packed >> 16;
                const double2 f = gfac[g + i];
                if ((tj & thr) != thr) continue;
                // two register bits (a CP between two register-resident qubits) is the common case: 4 amplitudes
#define SPZ_M4(A, B, C, D) cmul(ar[A], ai[A], f.x, f.y); cmul(ar[B], ai[B], f.x, f.y); cmul(ar[C], ai[C], f.x, f.y); cmul(ar[D], ai[D], f.x, f.y)
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
                        if (((unsigned)k & m) == m) cmul(ar[k], ai[k], f.x, f.y);
                    break;
                }
#undef SPZ_M4
            }
            continue;
        }
        const unsigned long long ocm = ins.outer_cmask;
        if ((base & ocm) != ocm) continue; // an outer control is 0 for this whole tile
        if (op == TI_GATE) {
            flush_diag();
            // t_mask holds the host-computed set of register indices k whose register-bit controls are all set
            const unsigned km = ((tj & ins.thr_cmask) == ins.thr_cmask) ? ins.t_mask : 0u;
            switch (ins.kind) {
            case SPZ_GATE_H: butterfly_pos<SPZ_GATE_H>(ins.rpos, ar, ai, ins.s, km); break;
            case SPZ_GATE_X: butterfly_pos<SPZ_GATE_X>(ins.rpos, ar, ai, ins.s, km); break;
            case SPZ_GATE_Y: butterfly_pos<SPZ_GATE_Y>(ins.rpos, ar, ai, ins.s, km); break;
            case SPZ_GATE_RX: butterfly_pos<SPZ_GATE_RX>(ins.rpos, ar, ai, ins.s, km); break;
            case SPZ_GATE_RY: butterfly_pos<SPZ_GATE_RY>(ins.rpos, ar, ai, ins.s, km); break;
            case SPZ_GATE_U: butterfly_pos<SPZ_GATE_U>(ins.rpos, ar, ai, ins.s, km); break;
            default: break;
            }
            continue;
        }
        // ---- TI_DIAG ----
        const int kind = ins.kind;
        const int tw = ins.t_where;
        bool outer_hi = false;
        if (tw == 0) outer_hi = ins.const_hi ? (ins.const_hi == 2) : (bool)((base >> ins.outer_target) & 1ull);
        if constexpr (EXACT) {
            if (tw == 0 && !outer_hi && kind != SPZ_GATE_RZ) continue; // Z / P leave target-bit-0 amplitudes alone
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
        // (merged mode never sees TI_DIAG: the host folds diagonal gates into TI_RUN groups)
    }
    if (have_regs) {
        flush_diag();
        store_regs();
        __syncthreads();
    }

    // ---- shared -> global ----
    for (unsigned v = threadIdx.x; v < n_vec; v += nthr) {
        const unsigned j = v << 1;
        const unsigned long long g = base + seg_off[j >> L] + (j & seg_mask);
        const unsigned s0 = swz(j);
        const bool flip = s0 & 1u;
        const double2 r = *reinterpret_cast<const double2 *>(sre + (s0 & ~1u));
        const double2 m = *reinterpret_cast<const double2 *>(sim + (s0 & ~1u));
        *reinterpret_cast<double2 *>(a.re + g) = flip ? make_double2(r.y, r.x) : r;
        *reinterpret_cast<double2 *>(a.im + g) = flip ? make_double2(m.y, m.x) : m;
    }
}

#ifndef SPZ_CPU_EMULATION
int tile_prepare(spz_state *st) {
    if (!st->d_ops) {
        const size_t cap = (size_t)4 << 20;
        SPZ_CUDA(cudaMallocAsync(&st->d_ops, cap, st->stream)); // (stream-ordered, so that tile_ring_alloc may regrow it with cudaFreeAsync)
        st->d_ops_bytes = cap;
        st->d_ops_cursor = 0;
    }
    const int max_smem = (int)(sizeof(double) * 2u * ((size_t)1 << kMaxTileBits) + kMaxTileGroups * (sizeof(double2) + sizeof(unsigned)) + 16 +
                               kMaxSmemInstr * sizeof(TileInstr));
    SPZ_CUDA(cudaFuncSetAttribute(k_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SPZ_CUDA(cudaFuncSetAttribute(k_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    return SPZ_OK;
}

// Programs are staged in a device ring buffer: a pass's program must stay intact until its kernel has run, so the cursor
// only wraps after a stream synchronise.
int tile_ring_alloc(spz_state *st, size_t bytes, char **slot) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (st->d_ops_bytes < bytes || !st->d_ops) {
        // (rare: the buffer is sized in tile_prepare for programs within the scheduler's per-pass op limit.)  Stream-ordered
        // allocation: cudaMalloc / cudaFree synchronise the whole device, which must not happen while another shard's handshake
        // kernel spins on the same GPU; the old buffer is released behind the passes that still read it.
        if (st->d_ops) { SPZ_CUDA(cudaFreeAsync(st->d_ops, st->stream)); st->d_ops = nullptr; st->d_ops_bytes = 0; }
        const size_t cap = std::max<size_t>(bytes * 2, (size_t)4 << 20);
        SPZ_CUDA(cudaMallocAsync(&st->d_ops, cap, st->stream));
        st->d_ops_bytes = cap;
        st->d_ops_cursor = 0;
    }
    if (st->d_ops_cursor + bytes > st->d_ops_bytes) {
        SPZ_CUDA(cudaStreamSynchronize(st->stream));
        st->d_ops_cursor = 0;
    }
    *slot = static_cast<char *>(st->d_ops) + st->d_ops_cursor;
    st->d_ops_cursor += bytes;
    return SPZ_OK;
}

int launch_tile_program(spz_state *st, const TilePlan &plan, const TileInstr *prog, int n_instr, const TileGroup *groups,
                        int n_groups, const TileTerm *terms, int n_terms, bool exact) {
    if (n_instr <= 0) return SPZ_OK;
    if (plan.tile_bits > kMaxTileBits || plan.tile_bits < kRegBits || plan.n_high > kMaxHigh || plan.low_bits < 1 ||
        plan.tile_bits != plan.low_bits + plan.n_high || plan.tile_bits > st->n) {
        set_error("bad tile plan T=%d L=%d H=%d n=%d", plan.tile_bits, plan.low_bits, plan.n_high, st->n);
        return SPZ_ERR_INVALID_ARG;
    }
    if (n_groups > kMaxTileGroups) { set_error("internal: %d diagonal groups exceed the shared-memory table", n_groups); return SPZ_ERR_INVALID_ARG; }
    // the TMA kernel takes every merged-mode pass it can (full 12-bit tiles, program within its shared-memory budget)
    Tile3Launch t3{};
    bool v3 = false;
    if (!exact && tile3_enabled()) SPZ_TRY(prepare_tile3(st, plan, prog, n_instr, groups, n_groups, terms, n_terms, &t3, &v3));
    const size_t prog_bytes = (sizeof(TileInstr) * (size_t)n_instr + 255) & ~(size_t)255;
    const size_t group_bytes = (sizeof(TileGroup) * (size_t)n_groups + 255) & ~(size_t)255;
    const size_t term_bytes = (sizeof(TileTerm) * (size_t)n_terms + 255) & ~(size_t)255;
    TileArgs a{};
    size_t smem = 0;
    if (!v3) {
    char *slot = nullptr;
    SPZ_TRY(tile_ring_alloc(st, prog_bytes + group_bytes + term_bytes, &slot));
    SPZ_CUDA(cudaMemcpyAsync(slot, prog, sizeof(TileInstr) * (size_t)n_instr, cudaMemcpyHostToDevice, st->stream));
    if (n_groups > 0)
        SPZ_CUDA(cudaMemcpyAsync(slot + prog_bytes, groups, sizeof(TileGroup) * (size_t)n_groups, cudaMemcpyHostToDevice, st->stream));
    if (n_terms > 0)
        SPZ_CUDA(cudaMemcpyAsync(slot + prog_bytes + group_bytes, terms, sizeof(TileTerm) * (size_t)n_terms, cudaMemcpyHostToDevice, st->stream));

    a.re = st->re; a.im = st->im;
    a.prog = reinterpret_cast<const TileInstr *>(slot);
    a.groups = reinterpret_cast<const TileGroup *>(slot + prog_bytes);
    a.terms = reinterpret_cast<const TileTerm *>(slot + prog_bytes + group_bytes);
    a.n_groups = n_groups;
    a.n_instr = n_instr;
    a.T = plan.tile_bits; a.L = plan.low_bits; a.n_high = plan.n_high;
    for (int k = 0; k < plan.n_high; ++k) a.high[k] = plan.high[k];
    smem = sizeof(double) * 2u * ((size_t)1 << plan.tile_bits) + (size_t)n_groups * (sizeof(double2) + sizeof(unsigned));
    smem = (smem + 15) & ~(size_t)15;
    a.prog_off = (unsigned)smem;
    a.prog_in_smem = n_instr <= kMaxSmemInstr ? 1 : 0;
    if (a.prog_in_smem) smem += sizeof(TileInstr) * (size_t)n_instr;
    } // !v3
    const unsigned grid = (unsigned)((uint64_t)st->len >> plan.tile_bits);
    const unsigned threads = 1u << (plan.tile_bits - kRegBits);
    auto launch = [&](unsigned first, unsigned count) {
        if (v3) { run_tile3(st, t3, first, count); return; }
        a.tile_offset = first;
        if (exact) k_tile<true><<<count, threads, smem, st->stream>>>(a);
        else k_tile<false><<<count, threads, smem, st->stream>>>(a);
    };
    // Directly after an overlapped exchange (dist.cu) the shard arrives in K contiguous chunks.  Tile numbers count through the
    // bits outside the tile, so as long as the top log2(parts) local bits are outside the tile, the chunks are ranges of the
    // grid: start on the first while the others are still on the wire.
    int K = 0;
    cudaEvent_t ev[8] = {};
    if (take_chunks(st, &K, ev)) {
        int parts = 1;
        for (int bits = 1; (1 << bits) <= K && (grid >> bits) >= 1; ++bits) {
            const int q = st->n - bits; // the next lower local bit must not be a tile bit
            bool is_tile = q < plan.low_bits;
            for (int k = 0; k < plan.n_high; ++k) if (plan.high[k] == q) is_tile = true;
            if (is_tile) break;
            parts = 1 << bits;
        }
        for (int j = 0; j < parts; ++j) {
            SPZ_CUDA(cudaStreamWaitEvent(st->stream, ev[(j + 1) * (K / parts) - 1], 0));
            const unsigned first = (unsigned)((uint64_t)grid * j / parts), end = (unsigned)((uint64_t)grid * (j + 1) / parts);
            launch(first, end - first);
        }
        if (parts > 1) count_launch(parts - 1);
    } else {
        launch(0, grid);
    }
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

#endif // !SPZ_CPU_EMULATION

} // namespace spz
