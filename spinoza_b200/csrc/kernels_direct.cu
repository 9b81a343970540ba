// kernels_direct.cu -- host-side dispatch of the one-gate-per-pass kernels (see kernels_direct.cuh).
#include <algorithm>

#include "kernels_direct.cuh"

namespace spz {

// Tuned on B200 with tools/bw_probe.cu (profiles/round1_bw_probe.md).
#ifndef SPZ_W
#define SPZ_W 4
#endif
#ifndef SPZ_U
#define SPZ_U 4
#endif
#ifndef SPZ_THREADS
#define SPZ_THREADS 256
#endif
#ifndef SPZ_POL
#define SPZ_POL 0
#endif

constexpr int W = SPZ_W, U = SPZ_U, THREADS = SPZ_THREADS, POL = SPZ_POL;
constexpr int LOGW = LogW<W>::v;

// Every launch of this file goes through SPZ_LAUNCH so that tests/emu/ can run the dispatch code below -- not a copy of
// it -- on the CPU (these kernels have no barriers: the emulation simply loops over blocks and threads).
#ifdef SPZ_CPU_EMULATION
#define SPZ_LAUNCH(kern, grid, threads, stream, args) spz_emu::run_flat((grid), (threads), [&]() { kern(args); })
#else
#define SPZ_LAUNCH(kern, grid, threads, stream, args) kern<<<(grid), (threads), 0, (stream)>>>(args)
#endif

template <int KIND, int NINS>
static void launch_vec(const PairArgs &a, cudaStream_t s) {
    const long long per_block = (long long)THREADS * U;
    const unsigned grid = (unsigned)((a.nvec + per_block - 1) / per_block);
    auto kern = k_pair_vec<KIND, NINS, W, U, THREADS, POL>;
    SPZ_LAUNCH(kern, grid, THREADS, s, a);
}
template <int KIND, int NINS>
static void launch_low(const PairArgs &a, cudaStream_t s) {
    const long long per_block = (long long)THREADS * U;
    const unsigned grid = (unsigned)((a.nvec + per_block - 1) / per_block);
    auto kern = k_pair_low<KIND, NINS, W, U, THREADS, POL>;
    SPZ_LAUNCH(kern, grid, THREADS, s, a);
}

template <int KIND>
static void dispatch_kind(bool low, const PairArgs &a, cudaStream_t s) {
    if (low) {
        switch (a.nins) {
        case 0: launch_low<KIND, 0>(a, s); break;
        case 1: launch_low<KIND, 1>(a, s); break;
        default: launch_low<KIND, -1>(a, s); break;
        }
    } else {
        switch (a.nins) {
        case 1: launch_vec<KIND, 1>(a, s); break;
        case 2: launch_vec<KIND, 2>(a, s); break;
        case 3: launch_vec<KIND, 3>(a, s); break;
        default: launch_vec<KIND, -1>(a, s); break;
        }
    }
}

static int launch_gate_on(spz_state *st, double *re, double *im, int n, const GateK &g, uint64_t ctrl_mask, int target, uint64_t neg_mask = 0);

#ifndef SPZ_CPU_EMULATION
// Streaming (see launch_gate): the number of pieces if an op that acts inside the pieces may go to the per-piece lanes now, else 0.
int lanes_possible(spz_state *st, int *bits_out) {
    spz_state::Arrival &a = st->arrival;
    if (!((a.pending || a.lanes_active || (a.streaming && st->n >= 24)) && a.chunks > 1)) return 0;
    int bits = 0;
    while ((1 << bits) < a.chunks) ++bits;
    if (st->n - bits < 12) return 0;
    *bits_out = bits;
    return a.chunks;
}
// The lanes wait for what they follow: the pieces of an upload in flight, or (restart after a join) what the main stream has been
// given so far -- for a shard, including an exchange still in flight.
int lanes_start(spz_state *st) {
    spz_state::Arrival &a = st->arrival;
    if (a.pending) {
        for (int k = 0; k < a.chunks; ++k) SPZ_CUDA(cudaStreamWaitEvent(a.lane[k], a.ev[k], 0));
        a.pending = false;
    } else if (!a.lanes_active) {
        if (st->dist) SPZ_TRY(dist_join(st));
        SPZ_CUDA(cudaEventRecord(a.ready, st->stream));
        for (int k = 0; k < a.chunks; ++k) SPZ_CUDA(cudaStreamWaitEvent(a.lane[k], a.ready, 0));
    }
    a.lanes_active = true;
    return SPZ_OK;
}
// A constant diagonal factor on every amplitude whose (local) controls are set -- a diagonal gate whose target is a global qubit of
// a sharded register -- piece by piece on the lanes when the state is being streamed.  False: the caller runs it on the whole shard.
bool lanes_diag_const(spz_state *st, const GateK &g, uint64_t local_cmask, int hi, int *rc) {
    int bits = 0;
    const int chunks = lanes_possible(st, &bits);
    if (!chunks) return false;
    const int nl = st->n - bits;
    *rc = lanes_start(st);
    const unsigned c_piece = (unsigned)(local_cmask >> nl);
    const uint64_t c_local = local_cmask & ((1ull << nl) - 1ull);
    cudaStream_t main_stream = st->stream;
    for (int k = 0; k < chunks && *rc == SPZ_OK; ++k) {
        if (((unsigned)k & c_piece) != c_piece) continue;
        st->stream = st->arrival.lane[k];
        *rc = diag_const_on(st, st->re + ((size_t)k << nl), st->im + ((size_t)k << nl), 1ll << nl, g, c_local, hi);
    }
    st->stream = main_stream;
    return true;
}
#endif

int launch_gate(spz_state *st, const GateK &g, uint64_t ctrl_mask, int target) {
#ifndef SPZ_CPU_EMULATION
    // Directly after an overlapped exchange (dist.cu) the shard lands in K contiguous chunks.  A gate whose target and controls
    // lie below the chunk bits acts on every chunk separately -- a chunk is a register of n - log2(K) qubits at an offset -- so it
    // runs chunk by chunk behind the wire: the one-gate pass that asked for the exchange costs 1/K of its time on top of it.
    // Behind an asynchronous upload (spz_upload_async), and until the download that ends the round trip, a RUN of gates that act
    // inside the pieces goes to one stream per piece: every gate of the run works on piece 0 while piece 1 is still on the bus
    // (and, at the other end, piece 0 leaves while piece 3 is still being computed).  A gate acts inside the pieces when its
    // target lies below the piece bits, or when it is diagonal (a piece bit is a constant of the piece: a constant factor, as
    // for a global qubit of a sharded register); a control on a piece bit selects pieces.  The first op that needs the whole
    // state joins the lanes.
    {
        int bits = 0;
        const int chunks = lanes_possible(st, &bits);
        const int nl = st->n - bits;
        const bool diag = g.kind == SPZ_GATE_Z || g.kind == SPZ_GATE_P || g.kind == SPZ_GATE_RZ;
        const bool t_piece = target >= nl;
        if (chunks && target >= 0 && target < st->n && !((ctrl_mask >> target) & 1ull) && (st->n >= 64 || !(ctrl_mask >> st->n)) &&
            (!t_piece || diag)) {
            SPZ_TRY(lanes_start(st));
            const unsigned c_piece = (unsigned)(ctrl_mask >> nl);
            const uint64_t c_local = ctrl_mask & ((1ull << nl) - 1ull);
            cudaStream_t main_stream = st->stream;
            int rc = SPZ_OK;
            for (int k = 0; k < chunks && rc == SPZ_OK; ++k) {
                if (((unsigned)k & c_piece) != c_piece) continue; // a control that is 0 throughout this piece
                const size_t off = (size_t)k << nl;
                st->stream = st->arrival.lane[k]; // (the launch helpers take the stream from the state)
                if (t_piece) rc = diag_const_on(st, st->re + off, st->im + off, 1ll << nl, g, c_local, (k >> (target - nl)) & 1);
                else rc = launch_gate_on(st, st->re + off, st->im + off, nl, g, c_local, target);
            }
            st->stream = main_stream;
            return rc;
        }
    }
    int K = 0;
    cudaEvent_t ev[8] = {};
    if (take_chunks(st, &K, ev)) {
        int parts = 1;
        for (int bits = 1; (1 << bits) <= K && st->n - bits >= 12; ++bits) {
            const int q = st->n - bits;
            if (q == target || ((ctrl_mask >> q) & 1ull)) break;
            parts = 1 << bits;
        }
        int bits = 0;
        while ((1 << bits) < parts) ++bits;
        if (target < 0 || target >= st->n - bits) parts = 1, bits = 0; // (out of range: reported below)
        for (int j = 0; j < parts; ++j) {
            SPZ_CUDA(cudaStreamWaitEvent(st->stream, ev[(j + 1) * (K / parts) - 1], 0));
            const size_t off = (size_t)j << (st->n - bits);
            SPZ_TRY(launch_gate_on(st, st->re + off, st->im + off, st->n - bits, g, ctrl_mask, target));
        }
        return SPZ_OK;
    }
#else
    SPZ_TRY(join_pending(st));
#endif
    return launch_gate_on(st, st->re, st->im, st->n, g, ctrl_mask, target);
}

// A gate under signed controls (the spz_mc_apply_signed extension): every qubit of ctrl_mask is a control, those also in
// neg_mask fire on 0 instead of 1.  The pair set is the same shape -- zero bits inserted at every control position -- only the
// value OR-ed back in differs, so the kernels and their traffic are those of the all-ones form.  Whole-state op: joins anything
// in flight.
int launch_gate_signed(spz_state *st, const GateK &g, uint64_t ctrl_mask, uint64_t neg_mask, int target) {
    if (neg_mask & ~ctrl_mask) { set_error("negative controls 0x%llx are not a subset of the controls 0x%llx", (unsigned long long)neg_mask, (unsigned long long)ctrl_mask); return SPZ_ERR_INVALID_ARG; }
    if (!neg_mask) return launch_gate(st, g, ctrl_mask, target);
    SPZ_TRY(join_pending(st));
    return launch_gate_on(st, st->re, st->im, st->n, g, ctrl_mask, target, neg_mask);
}

// the gate on the register of n qubits at (re, im): the state's own arrays, or one contiguous chunk of them
static int launch_gate_on(spz_state *st, double *re, double *im, int n, const GateK &g, uint64_t ctrl_mask, int target, uint64_t neg_mask) {
    if (target < 0 || target >= n) { set_error("target %d out of range for %d qubits", target, n); return SPZ_ERR_INVALID_ARG; }
    if ((ctrl_mask >> target) & 1ull) { set_error("target %d is also a control", target); return SPZ_ERR_INVALID_ARG; }
    if (n < 64 && (ctrl_mask >> n)) { set_error("control mask 0x%llx exceeds %d qubits", (unsigned long long)ctrl_mask, n); return SPZ_ERR_INVALID_ARG; }

    const uint64_t low_bits_mask = (1ull << LOGW) - 1ull;
    const int n_ctrl = __builtin_popcountll(ctrl_mask);
    const int n_ctrl_hi = __builtin_popcountll(ctrl_mask & ~low_bits_mask);
    const bool low = target < LOGW;
    const int nins_vec = n_ctrl_hi + (low ? 0 : 1);

    if (n - LOGW - nins_vec >= 0 && nins_vec <= kMaxIns) {
        PairArgs a{};
        a.re = re; a.im = im;
        a.nvec = 1ll << (n - LOGW - nins_vec);
        a.setmask = ctrl_mask & ~neg_mask & ~low_bits_mask;
        a.tbit = low ? 0ull : (1ull << target);
        a.lane_cmask = (int)(ctrl_mask & low_bits_mask);
        a.lane_cval = (int)(ctrl_mask & ~neg_mask & low_bits_mask);
        a.tlow = low ? target : 0;
        int k = 0;
        for (int q = LOGW; q < n; ++q)
            if (((ctrl_mask >> q) & 1ull) || (!low && q == target)) a.pos[k++] = (unsigned char)q;
        a.nins = k;
        for (int i = 0; i < 7; ++i) a.s[i] = g.s[i];
        switch (g.kind) {
        case SPZ_GATE_H: dispatch_kind<SPZ_GATE_H>(low, a, st->stream); break;
        case SPZ_GATE_X: dispatch_kind<SPZ_GATE_X>(low, a, st->stream); break;
        case SPZ_GATE_Y: dispatch_kind<SPZ_GATE_Y>(low, a, st->stream); break;
        case SPZ_GATE_Z: dispatch_kind<SPZ_GATE_Z>(low, a, st->stream); break;
        case SPZ_GATE_P: dispatch_kind<SPZ_GATE_P>(low, a, st->stream); break;
        case SPZ_GATE_RX: dispatch_kind<SPZ_GATE_RX>(low, a, st->stream); break;
        case SPZ_GATE_RY: dispatch_kind<SPZ_GATE_RY>(low, a, st->stream); break;
        case SPZ_GATE_RZ: dispatch_kind<SPZ_GATE_RZ>(low, a, st->stream); break;
        case SPZ_GATE_U: dispatch_kind<SPZ_GATE_U>(low, a, st->stream); break;
        default: set_error("gate kind %d has no pair kernel", g.kind); return SPZ_ERR_UNSUPPORTED;
        }
    } else {
        if (n_ctrl + 1 > kMaxIns) { set_error("too many controls"); return SPZ_ERR_INVALID_ARG; }
        ScalarArgs a{};
        a.re = re; a.im = im;
        a.npairs = 1ll << (n - 1 - n_ctrl);
        a.setmask = ctrl_mask & ~neg_mask;
        a.tbit = 1ull << target;
        a.kind = g.kind;
        int k = 0;
        for (int q = 0; q < n; ++q)
            if (((ctrl_mask >> q) & 1ull) || q == target) a.pos[k++] = (unsigned char)q;
        a.nins = k;
        for (int i = 0; i < 7; ++i) a.s[i] = g.s[i];
        const unsigned grid = (unsigned)std::min<long long>((a.npairs + 255) / 256, 148 * 8);
        SPZ_LAUNCH(k_pair_scalar, grid, 256, st->stream, a);
    }
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

int launch_swap(spz_state *st, int t0, int t1) {
    SPZ_TRY(join_pending(st));
    const int n = st->n;
    // assert!(state.n > t0 && state.n > t1) gates.rs:1377
    if (t0 < 0 || t1 < 0 || t0 >= n || t1 >= n) { set_error("swap operands (%d,%d) out of range for %d qubits", t0, t1, n); return SPZ_ERR_INVALID_ARG; }
    if (t0 == t1) return SPZ_OK; // the reference's scan finds no index with bit t0 = 0 and bit t0 = 1
    SwapArgs a{};
    a.re = st->re; a.im = st->im;
    a.lo = std::min(t0, t1); a.hi = std::max(t0, t1);
    if (a.lo >= LOGW && n - 2 - LOGW >= 0) {
        a.nvec = 1ll << (n - 2 - LOGW);
        const unsigned grid = (unsigned)((a.nvec + THREADS - 1) / THREADS);
        auto kern = k_swap_vec<W, THREADS>;
        SPZ_LAUNCH(kern, grid, THREADS, st->stream, a);
    } else {
        a.nvec = 1ll << (n - 2);
        const unsigned grid = (unsigned)std::min<long long>((a.nvec + 255) / 256, 148 * 16);
        SPZ_LAUNCH(k_swap_scalar, grid, 256, st->stream, a);
    }
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

} // namespace spz
