// kernels_xgate.cuh -- the exchange of a sharded register fused with the gate that asked for it.
//
// STATUS: opt-in (SPZ_DIST_FUSE_GATE=1).  Written after round 1's GPU budget was spent: the protocol below is checked on
// the CPU emulation (tests/test_xgate_cpu_emulation.py: two ranks, concurrent thread blocks, ThreadSanitizer); it has not
// run on NVLink yet.
//
// A non-diagonal gate G on a global qubit is lowered to "exchange the rank bit with local bit l, then apply G on l"
// (dist_plan.h).  Done as two kernels that is an NVLink-bound pass followed by a full HBM pass for one gate -- and the
// dry run of the sharded QFT shows those lone gates are exactly what follows most exchanges.
//
// Notation: a_gl = the amplitudes with rank bit g and local bit l (each a quarter... each HALF a shard).  Rank A (g = 0) owns
// a00, a01; rank B (g = 1) owns a10, a11.  After "exchange, then G on l":
//     A holds  (lo', hi') = G(a00, a10)  in its l = 0 / l = 1 halves,      B holds  (lo', hi') = G(a01, a11).
// So A needs a10 and B needs a01: each rank READS half a shard from its partner over NVLink (the same volume as the plain
// exchange) and keeps all its writes local.  One thread doing both updates of its pair -- the way the plain in-place
// exchange splits the work -- would need a10 AND a11 plus two remote writes: twice the NVLink volume.
//
// The hazard of working in place: A overwrites its l = 1 half (old a01) with hi' while B still has to read a01 from it, and B
// overwrites its l = 0 half (old a10) which A reads.  Each CTA therefore works in steps: load the remote operands of one
// block of pairs, compute, store the output that goes to the slot nobody else reads, then tell the partner CTA (same
// blockIdx on the other rank, same blocks in the same order) "I have read step i" and wait for the same message before
// storing the other output.  Flags are per-CTA counters in the partner's control block, written with st.release.sys after
// a system fence and polled with ld.acquire.sys, like the handshakes of dist.cu.
#pragma once

#include "kernels_direct.cuh" // Vec / ldv / stv / LogW, gate_math.cuh

namespace spz {

constexpr int kMaxXgCtas = 128; // flag slots per control block

struct XGArgs {
    double *mine_re, *mine_im;
    const double *peer_re, *peer_im;
    long long nvec;                   // W-vectors of pairs: 2^(n_local - 1) / W
    int lq;                           // local physical bit that receives the global qubit (>= log2 W)
    int my_bit;                       // this rank's value of the rank bit being exchanged
    unsigned long long *peer_flag;    // partner's control block: xg_flag[0..gridDim.x)
    const unsigned long long *my_flag; // this rank's control block: xg_flag[0..gridDim.x), written by the partner
    unsigned long long flag_base;     // flags of this launch count flag_base + 1, flag_base + 2, ...
    unsigned long long *err;          // set when the partner does not answer within the spin timeout
    unsigned long long timeout_ns;
    double s[7];                      // gate scalars (gate_math.cuh)
};

#ifdef SPZ_CPU_EMULATION
// provided by the emulation harness (std::atomic / clock)
#else
__device__ __forceinline__ unsigned long long xg_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void xg_release(unsigned long long *p, unsigned long long v) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long xg_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void xg_pause() { __nanosleep(100); }
#endif

template <int KIND, int W, int U, int THREADS>
__global__ void __launch_bounds__(THREADS) k_exchange_gate(const XGArgs a) {
    const unsigned long long lbit = 1ull << a.lq;
    const long long per = (long long)THREADS * U;
    unsigned long long step = 0;
    for (long long blk = (long long)blockIdx.x * per; blk < a.nvec; blk += (long long)gridDim.x * per) {
        unsigned long long idx[U];
        Vec<W> lo_r[U], lo_i[U], hi_r[U], hi_i[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long v = blk + threadIdx.x + (long long)u * THREADS;
            if (v < a.nvec) {
                const unsigned long long base = insert_zero((unsigned long long)v << LogW<W>::v, a.lq);
                idx[u] = a.my_bit ? (base | lbit) : base; // the half this rank keeps AND the half it reads from the partner
                if (a.my_bit) { // B: lo = a01 (partner's l = 1 half), hi = a11 (mine)
                    lo_r[u] = ldv<W, 0>(a.peer_re + idx[u]); lo_i[u] = ldv<W, 0>(a.peer_im + idx[u]);
                    hi_r[u] = ldv<W, 0>(a.mine_re + idx[u]); hi_i[u] = ldv<W, 0>(a.mine_im + idx[u]);
                } else {        // A: lo = a00 (mine), hi = a10 (partner's l = 0 half)
                    hi_r[u] = ldv<W, 0>(a.peer_re + idx[u]); hi_i[u] = ldv<W, 0>(a.peer_im + idx[u]);
                    lo_r[u] = ldv<W, 0>(a.mine_re + idx[u]); lo_i[u] = ldv<W, 0>(a.mine_im + idx[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long v = blk + threadIdx.x + (long long)u * THREADS;
            if (v < a.nvec) {
#pragma unroll
                for (int l = 0; l < W; ++l) pair_update<KIND>(a.s, lo_r[u].v[l], lo_i[u].v[l], hi_r[u].v[l], hi_i[u].v[l]);
                // the output whose slot only this rank touches
                if (a.my_bit) { stv<W, 0>(a.mine_re + idx[u], hi_r[u]); stv<W, 0>(a.mine_im + idx[u], hi_i[u]); }
                else          { stv<W, 0>(a.mine_re + idx[u], lo_r[u]); stv<W, 0>(a.mine_im + idx[u], lo_i[u]); }
            }
        }
        // every thread of the CTA has consumed its remote loads: tell the partner CTA, and wait until it has consumed its own
        ++step;
        int abort = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            xg_release(a.peer_flag + blockIdx.x, a.flag_base + step);
            const unsigned long long t0 = xg_timer_ns();
            while (xg_acquire(a.my_flag + blockIdx.x) < a.flag_base + step) {
                if (xg_timer_ns() - t0 > a.timeout_ns) { *a.err = a.flag_base + step; abort = 1; break; }
                xg_pause();
            }
        }
        if (__syncthreads_or(abort)) return; // the partner is gone: the error word is set, do not touch memory it may still be reading
        // the other output goes where the partner has just finished reading
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long v = blk + threadIdx.x + (long long)u * THREADS;
            if (v < a.nvec) {
                const unsigned long long other = idx[u] ^ lbit;
                if (a.my_bit) { stv<W, 0>(a.mine_re + other, lo_r[u]); stv<W, 0>(a.mine_im + other, lo_i[u]); }
                else          { stv<W, 0>(a.mine_re + other, hi_r[u]); stv<W, 0>(a.mine_im + other, hi_i[u]); }
            }
        }
    }
}

} // namespace spz
