// dist.cu -- multi-GPU layer: one process per GPU, shards exchanged over NVLink through CUDA IPC peer pointers.
//
// See dist_plan.h for the lowering rules.  This file holds the device side:
//   * k_exchange_*   : the pairwise half-shard exchange.  Rank pair (r, r ^ 2^k) swaps r's [local bit lq = 1]
//                      half with the partner's [lq = 0] half IN PLACE: every thread loads one vector from its own
//                      HBM and one from the partner's HBM (ld.global on the IPC-mapped peer pointer, i.e. over
//                      NVLink) and stores them crosswise (one local store, one peer store).  The pair list is split
//                      between the two ranks, so each direction of the link carries exactly half a shard
//                      (16 * 2^(n_local-1) bytes) and no staging buffer is needed next to a 137 GB shard.
//   * k_handshake    : stream-ordered cross-process barrier between the two partners (epoch flags in IPC-shared
//                      device memory, st.release.sys / ld.acquire.sys), before and after each exchange.
//   * k_allreduce    : scalar sum across all ranks through the same IPC control blocks (prob0 / norm / <O>).
// No NCCL: the data path is peer loads/stores issued by our own kernels.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dist_plan.h"
#include "kernels_direct.cuh"

#include "kernels_xgate.cuh"

namespace spz {

constexpr int kMaxRanks = 16;
constexpr int kMaxChunks = 8;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct CtrlBlock { // lives in device memory of each rank, mapped by every peer
    unsigned long long ready[kMaxRanks];
    unsigned long long done[kMaxRanks];
    unsigned long long done_k[kMaxChunks][kMaxRanks]; // chunk k of an overlapped exchange has landed
    unsigned long long red_epoch[2][kMaxRanks];
    double red_slot[2][kMaxRanks];
    double red_result;
    unsigned long long error;
    unsigned long long xg_flag[kMaxXgCtas]; // fused exchange + gate: "partner CTA b has read step i" (kernels_xgate.cuh)
};

struct IpcBlob {
    cudaIpcMemHandle_t re, im, ctrl;
    int rank, device;
    long long len;
    char pad[SPZ_IPC_BLOB_BYTES - 3 * sizeof(cudaIpcMemHandle_t) - 2 * sizeof(int) - sizeof(long long)];
};
static_assert(sizeof(IpcBlob) == SPZ_IPC_BLOB_BYTES, "blob size");

struct DistCtx {
    DistPlan plan;
    int rank = 0, world = 1;
    CtrlBlock *ctrl = nullptr;
    CtrlBlock *peer_ctrl[kMaxRanks] = {};
    double *peer_re[kMaxRanks] = {}, *peer_im[kMaxRanks] = {};
    CtrlBlock **d_table = nullptr; // device copy of peer_ctrl[] for k_allreduce
    bool connected = false, ipc = false;
    unsigned long long epoch = 0, red_epoch = 0;
    // overlapped exchange (second stream): the shard is exchanged in n_chunks chunks along the top bits of the pair list (the
    // top local bits that are not the traded one); the consumer that follows -- a fused pass or a one-gate pass -- starts on
    // chunk 0 while the others are still on the wire
    cudaStream_t xstream = nullptr;
    cudaEvent_t ev_main = nullptr, ev_chunk[kMaxChunks] = {};
    bool split_pending = false;
    int split_chunks = 0;      // chunks of the exchange in flight
    int split_lq = -1;         // the local bit it trades: the chunks are contiguous ranges of the shard iff it is below the chunk bits
    int n_chunks = 4;          // SPZ_XCHG_CHUNKS (1, 2, 4 or 8)
    bool overlap = true;
    int xchg_ctas = 40; // CTAs of the persistent exchange kernel in overlapped mode (SPZ_XCHG_CTAS)
    bool xchg_tma = false;   // SPZ_XCHG_TMA=1: bulk-copy variant (k_exchange_tma), SPZ_XCHG_TMA_CTAS one-warp CTAs
    int xchg_tma_ctas = 32;
    // fused exchange + gate (opt-in, SPZ_DIST_FUSE_GATE=1): flag values already used, and the size of its persistent grid
    unsigned long long xg_base = 0;
    int xg_ctas = 128;  // SPZ_XG_CTAS; every CTA must be resident at once, so <= the number of SMs
    // The register is (very likely) still the basis state |basis_index> it was set to (spz_dist_create, spz_reset_zero,
    // spz_set_basis): a hint only -- dist_place_basis checks the amplitudes before it relies on it
    bool basis_hint = false;
    uint64_t basis_index = 0;
    double n_placements = 0;
    // stats
    double n_exchanges = 0, bytes_sent = 0, ms_accum = 0, n_overlapped = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending, free_events;
};

static DistCtx *ctx_of(const spz_state *st) { return static_cast<DistCtx *>(st->dist); }

// device-resident table of peer control blocks (allocated at connect time, never lazily: see spz_create)
static int upload_peer_table(DistCtx *c) {
    if (!c->d_table) SPZ_CUDA(cudaMalloc(&c->d_table, sizeof(CtrlBlock *) * kMaxRanks));
    SPZ_CUDA(cudaMemcpy(c->d_table, c->peer_ctrl, sizeof(CtrlBlock *) * kMaxRanks, cudaMemcpyHostToDevice));
    return SPZ_OK;
}

int dist_total_qubits(const spz_state *st) { return ctx_of(st)->plan.n; }

// Dry-run support (spz_debug_compile_sharded, pure host code): a context that only knows the plan and the rank.
void dist_debug_attach(spz_state *st, int n_total, int world, int rank) {
    DistCtx *c = new DistCtx();
    c->plan.init(n_total, world);
    c->rank = rank;
    c->world = world;
    st->dist = c;
}
void dist_note_modified(spz_state *st) { if (st->dist) ctx_of(st)->basis_hint = false; } // (uploads: abi.cu)
void dist_debug_detach(spz_state *st) {
    delete ctx_of(st);
    st->dist = nullptr;
}

// ---- device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Tell the partner "I reached epoch e" and wait until it has too.  One thread.
__global__ void k_handshake(unsigned long long *partner_slot, const unsigned long long *my_slot, unsigned long long epoch,
                            unsigned long long *err) {
    __threadfence_system();
    st_release_sys(partner_slot, epoch);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(my_slot) < epoch) {
        if (globaltimer_ns() - t0 > kSpinTimeoutNs) { *err = epoch; break; }
        __nanosleep(200);
    }
}

// Scalar all-reduce (sum) over all ranks.  One thread per peer writes, thread 0 sums in rank order, so every
// rank obtains the bitwise-identical total.
__global__ void k_allreduce(CtrlBlock *mine, CtrlBlock *const *peers, int rank, int world, unsigned long long epoch,
                            const double *value_ptr, double value_imm, double *out) {
    __shared__ CtrlBlock *sp[kMaxRanks];
    const int par = (int)(epoch & 1ull);
    const double v = value_ptr ? *value_ptr : value_imm;
    if (threadIdx.x < world) sp[threadIdx.x] = peers[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < world) {
        CtrlBlock *p = sp[threadIdx.x];
        p->red_slot[par][rank] = v;
        __threadfence_system();
        st_release_sys(&p->red_epoch[par][rank], epoch);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t0 = globaltimer_ns();
        double acc = 0.0;
        for (int r = 0; r < world; ++r) {
            while (ld_acquire_sys(&mine->red_epoch[par][r]) < epoch) {
                if (globaltimer_ns() - t0 > kSpinTimeoutNs) { mine->error = epoch; break; }
                __nanosleep(200);
            }
            acc += *(volatile double *)&mine->red_slot[par][r];
        }
        *out = acc;
        mine->red_result = acc;
    }
}

// ---- the exchange ------------------------------------------------------------------------------------------
struct XArgs {
    double *mine_re, *mine_im, *peer_re, *peer_im;
    long long nvec_begin, nvec_end; // range of pair-list vectors this rank handles
    int lq;                         // local physical bit traded with the rank bit
    int my_bit;                     // this rank's value of the rank bit
};

template <int W, int U, int THREADS>
__global__ void __launch_bounds__(THREADS) k_exchange_vec(const XArgs a) {
    const long long v0 = a.nvec_begin + (long long)blockIdx.x * (THREADS * U) + threadIdx.x;
    const unsigned long long lbit = 1ull << a.lq;
    unsigned long long im_[U], ip_[U];
    Vec<W> mr[U], mi[U], pr[U], pi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long v = v0 + (long long)u * THREADS;
        if (v < a.nvec_end) {
            const unsigned long long base = insert_zero((unsigned long long)v << LogW<W>::v, a.lq);
            // low rank (bit 0) gives its lq=1 half and takes the partner's lq=0 half
            im_[u] = a.my_bit ? base : (base | lbit);
            ip_[u] = a.my_bit ? (base | lbit) : base;
            pr[u] = ldv<W, 0>(a.peer_re + ip_[u]); // NVLink loads first: longest latency
            pi[u] = ldv<W, 0>(a.peer_im + ip_[u]);
            mr[u] = ldv<W, 0>(a.mine_re + im_[u]);
            mi[u] = ldv<W, 0>(a.mine_im + im_[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long v = v0 + (long long)u * THREADS;
        if (v < a.nvec_end) {
            stv<W, 0>(a.peer_re + ip_[u], mr[u]);
            stv<W, 0>(a.peer_im + ip_[u], mi[u]);
            stv<W, 0>(a.mine_re + im_[u], pr[u]);
            stv<W, 0>(a.mine_im + im_[u], pi[u]);
        }
    }
}

// Persistent variant for the overlapped mode: a small fixed grid walks the range in blocks of THREADS * U vectors, so the
// exchange keeps only a few SMs busy (enough bytes in flight to fill the link) and leaves the rest to the fused pass.
template <int W, int U, int THREADS>
__global__ void __launch_bounds__(THREADS) k_exchange_vec_persistent(const XArgs a) {
    const unsigned long long lbit = 1ull << a.lq;
    const long long per = (long long)THREADS * U;
    for (long long blk = a.nvec_begin + (long long)blockIdx.x * per; blk < a.nvec_end; blk += (long long)gridDim.x * per) {
        unsigned long long im_[U], ip_[U];
        Vec<W> mr[U], mi[U], pr[U], pi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long v = blk + threadIdx.x + (long long)u * THREADS;
            if (v < a.nvec_end) {
                const unsigned long long base = insert_zero((unsigned long long)v << LogW<W>::v, a.lq);
                im_[u] = a.my_bit ? base : (base | lbit);
                ip_[u] = a.my_bit ? (base | lbit) : base;
                pr[u] = ldv<W, 0>(a.peer_re + ip_[u]);
                pi[u] = ldv<W, 0>(a.peer_im + ip_[u]);
                mr[u] = ldv<W, 0>(a.mine_re + im_[u]);
                mi[u] = ldv<W, 0>(a.mine_im + im_[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long v = blk + threadIdx.x + (long long)u * THREADS;
            if (v < a.nvec_end) {
                stv<W, 0>(a.peer_re + ip_[u], mr[u]);
                stv<W, 0>(a.peer_im + ip_[u], mi[u]);
                stv<W, 0>(a.mine_re + im_[u], pr[u]);
                stv<W, 0>(a.mine_im + im_[u], pi[u]);
            }
        }
    }
}

// TMA variant of the persistent exchange (opt-in, SPZ_XCHG_TMA=1): the same in-place swap, moved by the copy engine of the SM
// instead of by loads and stores of its threads.  A CTA is one warp; lane 0 issues, per step, four 1-D bulk loads of one
// contiguous run (mine re / im, partner's re / im over NVLink) into a shared-memory stage on an mbarrier, and when they have
// landed four bulk stores that write each side's run to the other side.  kXtStages stages are in flight per CTA.  The traded
// bit lq >= 8, so a run (the 2^lq amplitudes below the traded bit) is >= 2 KB; runs longer than kXtRun doubles are cut.
constexpr int kXtStages = 3;
constexpr unsigned kXtRun = 2048;          // doubles per bulk copy: 16 KB
__global__ void __launch_bounds__(32) k_exchange_tma(const XArgs a) {
#ifndef SPZ_CPU_EMULATION
    extern __shared__ __align__(128) unsigned char xsm[];
    const unsigned run = min((unsigned long long)kXtRun, 1ull << a.lq);      // doubles per piece
    const unsigned bytes = run * 8u;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(xsm + (size_t)kXtStages * 4u * kXtRun * 8u);
    const unsigned long long lbit = 1ull << a.lq;
    // pieces of the pair list: [nvec_begin, nvec_end) vectors of 4 doubles -> pieces of `run` doubles
    const unsigned long long p_begin = ((unsigned long long)a.nvec_begin * 4ull) / run, p_end = ((unsigned long long)a.nvec_end * 4ull) / run;
    if (threadIdx.x != 0) return;
    auto sh = [&](const void *p) { return (unsigned)__cvta_generic_to_shared(p); };
    for (int s = 0; s < kXtStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sh(bar + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto addr_of = [&](unsigned long long piece, unsigned long long *mine, unsigned long long *peer) {
        const unsigned long long base = insert_zero(piece * run, a.lq);
        *mine = a.my_bit ? base : (base | lbit);
        *peer = a.my_bit ? (base | lbit) : base;
    };
    auto issue_loads = [&](unsigned long long piece, int s) {
        unsigned long long im_, ip_;
        addr_of(piece, &im_, &ip_);
        unsigned char *st = xsm + (size_t)s * 4u * kXtRun * 8u;
        const unsigned b = sh(bar + s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4u * bytes) : "memory");
        const double *src[4] = {a.peer_re + ip_, a.peer_im + ip_, a.mine_re + im_, a.mine_im + im_}; // NVLink loads first
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sh(st + (size_t)k * kXtRun * 8u)), "l"(src[k]), "r"(bytes), "r"(b) : "memory");
    };
    unsigned phase[kXtStages] = {0, 0, 0};
    unsigned long long next = p_begin + blockIdx.x;
    // prologue: fill the stages
    int filled = 0;
    for (; filled < kXtStages && next < p_end; ++filled, next += gridDim.x) issue_loads(next, filled);
    unsigned long long cur = p_begin + blockIdx.x;
    for (int s = 0; cur < p_end; cur += gridDim.x, s = (s + 1) % kXtStages) {
        const unsigned b = sh(bar + s);
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(b), "r"(phase[s]) : "memory");
        phase[s] ^= 1u;
        unsigned long long im_, ip_;
        addr_of(cur, &im_, &ip_);
        unsigned char *st = xsm + (size_t)s * 4u * kXtRun * 8u;
        double *dst[4] = {a.mine_re + im_, a.mine_im + im_, a.peer_re + ip_, a.peer_im + ip_}; // what came from the partner stays here, mine goes there
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst[k]), "r"(sh(st + (size_t)k * kXtRun * 8u)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (next < p_end) {
            // the stage is refilled once its stores have read it
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue_loads(next, s);
            next += gridDim.x;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // every store complete before the done handshake that follows in the stream
#endif
}

__global__ void k_exchange_scalar(const XArgs a) {
    const unsigned long long lbit = 1ull << a.lq;
    for (long long v = a.nvec_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < a.nvec_end;
         v += (long long)gridDim.x * blockDim.x) {
        const unsigned long long base = insert_zero((unsigned long long)v, a.lq);
        const unsigned long long im_ = a.my_bit ? base : (base | lbit), ip_ = a.my_bit ? (base | lbit) : base;
        const double pr = a.peer_re[ip_], pi = a.peer_im[ip_], mr = a.mine_re[im_], mi = a.mine_im[im_];
        a.peer_re[ip_] = mr; a.peer_im[ip_] = mi; a.mine_re[im_] = pr; a.mine_im[im_] = pi;
    }
}

// Constant diagonal factor on every amplitude of the shard whose local controls are set (a diagonal gate whose
// target is a global qubit).  Same per-amplitude arithmetic as the pair kernels (gate_math.cuh).
__global__ void __launch_bounds__(256) k_diag_const(double *__restrict__ re, double *__restrict__ im, long long count,
                                                    unsigned long long cmask, int kind, int hi, GateK g) {
    const int nc = __popcll(cmask);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < count; p += (long long)gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)p, m = cmask;
        for (int k = 0; k < nc; ++k) { const int b = __ffsll((long long)m) - 1; x = insert_zero(x, b); m &= m - 1; }
        x |= cmask;
        double a = re[x], b = im[x];
        diag_update_rt(kind, g.s, hi != 0, a, b);
        re[x] = a; im[x] = b;
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
static int check_comm_error(spz_state *st) {
    DistCtx *c = ctx_of(st);
    unsigned long long err = 0;
    SPZ_CUDA(cudaMemcpyAsync(&err, &c->ctrl->error, sizeof err, cudaMemcpyDeviceToHost, st->stream));
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    if (err) { set_error("rank %d: peer did not arrive at epoch %llu within the spin timeout", c->rank, err); return SPZ_ERR_COMM; }
    return SPZ_OK;
}

int dist_join(spz_state *st);

// A pair of timing events for one exchange; the pairs of exchanges that have finished are recycled (keeps the pool bounded on
// long runs -- used by both exchange paths).
static int take_timing_events(DistCtx *c, std::pair<cudaEvent_t, cudaEvent_t> *ev) {
    if (c->pending.size() >= 32) {
        size_t keep = 0;
        for (auto &e : c->pending) {
            float ms = 0.f;
            if (cudaEventQuery(e.second) == cudaSuccess && cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) {
                c->ms_accum += ms;
                c->free_events.push_back(e);
            } else {
                c->pending[keep++] = e;
            }
        }
        c->pending.resize(keep);
        cudaGetLastError(); // cudaErrorNotReady from the query is not an error
    }
    if (!c->free_events.empty()) { *ev = c->free_events.back(); c->free_events.pop_back(); }
    else { SPZ_CUDA(cudaEventCreate(&ev->first)); SPZ_CUDA(cudaEventCreate(&ev->second)); }
    return SPZ_OK;
}

int dist_exchange(spz_state *st, int gbit, int lq) {
    DistCtx *c = ctx_of(st);
    c->basis_hint = false;
    if (!c->connected) { set_error("dist state used before spz_dist_connect"); return SPZ_ERR_COMM; }
    const int partner = c->rank ^ (1 << gbit);
    const int my_bit = (c->rank >> gbit) & 1;
    SPZ_TRY(join_pending(st)); // a previous overlapped exchange must have landed completely
    const unsigned long long e = ++c->epoch;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    SPZ_TRY(take_timing_events(c, &ev));
    XArgs a{};
    a.mine_re = st->re; a.mine_im = st->im; a.peer_re = c->peer_re[partner]; a.peer_im = c->peer_im[partner];
    a.lq = lq; a.my_bit = my_bit;
    const int n_local = st->n;
    constexpr int W = 4, U = 4, THREADS = 256;
    const bool vec = lq >= LogW<W>::v && n_local - 1 - LogW<W>::v >= 1;
    const long long nvec = vec ? (1ll << (n_local - 1 - LogW<W>::v)) : (1ll << (n_local - 1));
    auto launch_range = [&](long long b, long long e2, cudaStream_t stream) {
        XArgs r = a;
        r.nvec_begin = b; r.nvec_end = e2;
        const long long cnt = e2 - b, per = (long long)THREADS * U;
        if (vec) k_exchange_vec<W, U, THREADS><<<(unsigned)((cnt + per - 1) / per), THREADS, 0, stream>>>(r);
        else k_exchange_scalar<<<(unsigned)std::max<long long>(1, std::min<long long>((cnt + 255) / 256, 148 * 8)), 256, 0, stream>>>(r);
    };
    // The top bits of the pair-list index are the top local bits that are not lq: chunks of the pair list are chunks of the
    // shard along those bits.  Overlapped mode needs a vector path and enough vectors per chunk.
    int K = c->n_chunks;
    while (K > 1 && nvec < 64ll * K) K >>= 1;
    if (c->overlap && vec && nvec >= 64 && n_local >= 2) {
        // stream X: [ready handshake] { [chunk k: my half of it] [done_k handshake] ev_chunk[k] } for k = 0 .. K-1
        SPZ_CUDA(cudaEventRecord(c->ev_main, st->stream));
        SPZ_CUDA(cudaStreamWaitEvent(c->xstream, c->ev_main, 0));
        k_handshake<<<1, 1, 0, c->xstream>>>(&c->peer_ctrl[partner]->ready[c->rank], &c->ctrl->ready[partner], e, &c->ctrl->error);
        SPZ_CUDA(cudaEventRecord(ev.first, c->xstream));
        for (int h = 0; h < K; ++h) {
            const long long hb = h * (nvec / K) + my_bit * (nvec / (2 * K));
            XArgs r = a;
            r.nvec_begin = hb; r.nvec_end = hb + nvec / (2 * K);
            if (c->xchg_tma && lq >= 8 && ((r.nvec_end - r.nvec_begin) * 4) % (long long)std::min<unsigned long long>(kXtRun, 1ull << lq) == 0)
                k_exchange_tma<<<(unsigned)c->xchg_tma_ctas, 32, (size_t)kXtStages * 4u * kXtRun * 8u + 64, c->xstream>>>(r);
            else
                k_exchange_vec_persistent<W, U, THREADS><<<(unsigned)c->xchg_ctas, THREADS, 0, c->xstream>>>(r);
            k_handshake<<<1, 1, 0, c->xstream>>>(&c->peer_ctrl[partner]->done_k[h][c->rank], &c->ctrl->done_k[h][partner], e, &c->ctrl->error);
            if (h == K - 1) SPZ_CUDA(cudaEventRecord(ev.second, c->xstream));
            SPZ_CUDA(cudaEventRecord(c->ev_chunk[h], c->xstream));
        }
        c->pending.push_back(ev);
        c->split_pending = true;
        c->split_chunks = K;
        c->split_lq = lq;
        c->n_overlapped += 1;
        count_launch(1 + 2 * K);
    } else {
        k_handshake<<<1, 1, 0, st->stream>>>(&c->peer_ctrl[partner]->ready[c->rank], &c->ctrl->ready[partner], e, &c->ctrl->error);
        SPZ_CUDA(cudaEventRecord(ev.first, st->stream));
        launch_range(my_bit ? nvec / 2 : 0, my_bit ? nvec : nvec / 2, st->stream);
        SPZ_CUDA(cudaEventRecord(ev.second, st->stream));
        c->pending.push_back(ev);
        k_handshake<<<1, 1, 0, st->stream>>>(&c->peer_ctrl[partner]->done[c->rank], &c->ctrl->done[partner], e, &c->ctrl->error);
        count_launch(3);
    }
    SPZ_CUDA(cudaGetLastError());
    c->n_exchanges += 1;
    c->bytes_sent += 16.0 * (double)(1ll << (n_local - 1)); // half a shard leaves this GPU (re + im)
    return SPZ_OK;
}

// Exchange fused with the uncontrolled non-diagonal gate that asked for it (kernels_xgate.cuh).  Same bracket as the
// sequential exchange: ready handshake, one kernel, done handshake, all on the main stream.
bool dist_fuse_gate_enabled() {
    const char *e = std::getenv("SPZ_DIST_FUSE_GATE");
    return e && e[0] == '1';
}

int dist_exchange_gate(spz_state *st, int gbit, int lq, const GateK &g) {
    DistCtx *c = ctx_of(st);
    c->basis_hint = false;
    if (!c->connected) { set_error("dist state used before spz_dist_connect"); return SPZ_ERR_COMM; }
    // Bytes in flight decide the NVLink rate: a step of one CTA reads THREADS * U * W * 16 B = 64 KB from the partner, and a
    // step lasts a remote-load latency plus a flag round trip (several microseconds), so ~128 CTAs x 64 KB per step are
    // needed to cover ~700 GB/s.  Tuning on hardware: SPZ_XG_CTAS.
    constexpr int W = 4, U = 4, THREADS = 256;
    const int n_local = st->n;
    if (lq < LogW<W>::v || n_local - 1 - LogW<W>::v < 0) { set_error("internal: fused exchange needs a vector path"); return SPZ_ERR_INVALID_ARG; }
    const int partner = c->rank ^ (1 << gbit);
    // (checked before anything is launched or counted: an error here must leave this rank in step with its partner)
    const void *kern = nullptr;
    switch (g.kind) {
    case SPZ_GATE_H: kern = (const void *)k_exchange_gate<SPZ_GATE_H, W, U, THREADS>; break;
    case SPZ_GATE_X: kern = (const void *)k_exchange_gate<SPZ_GATE_X, W, U, THREADS>; break;
    case SPZ_GATE_Y: kern = (const void *)k_exchange_gate<SPZ_GATE_Y, W, U, THREADS>; break;
    case SPZ_GATE_RX: kern = (const void *)k_exchange_gate<SPZ_GATE_RX, W, U, THREADS>; break;
    case SPZ_GATE_RY: kern = (const void *)k_exchange_gate<SPZ_GATE_RY, W, U, THREADS>; break;
    case SPZ_GATE_U: kern = (const void *)k_exchange_gate<SPZ_GATE_U, W, U, THREADS>; break;
    default: set_error("internal: gate kind %d cannot be fused into an exchange", g.kind); return SPZ_ERR_INVALID_ARG;
    }
    // CTA b of one rank talks to CTA b of the other through spin flags, so every CTA of both kernels must be resident at once:
    // no more CTAs than the device can hold, half of that when the two shards share the device (a local group on one GPU).
    int per_sm = 0, n_sm = 0;
    SPZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0));
    SPZ_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, st->device));
    int resident = std::max(1, per_sm * n_sm);
    if (!c->ipc) resident = std::max(1, resident / 2);
    SPZ_TRY(join_pending(st));
    const unsigned long long e = ++c->epoch;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    SPZ_TRY(take_timing_events(c, &ev));
    XGArgs a{};
    a.mine_re = st->re; a.mine_im = st->im; a.peer_re = c->peer_re[partner]; a.peer_im = c->peer_im[partner];
    a.nvec = 1ll << (n_local - 1 - LogW<W>::v);
    a.lq = lq; a.my_bit = (c->rank >> gbit) & 1;
    a.peer_flag = c->peer_ctrl[partner]->xg_flag; a.my_flag = c->ctrl->xg_flag;
    a.flag_base = c->xg_base;
    a.err = &c->ctrl->error;
    a.timeout_ns = kSpinTimeoutNs;
    for (int k = 0; k < 7; ++k) a.s[k] = g.s[k];
    const long long per = (long long)THREADS * U;
    const int ctas = (int)std::max<long long>(1, std::min<long long>(std::min(std::min(c->xg_ctas, kMaxXgCtas), resident), (a.nvec + per - 1) / per));
    c->xg_base += (unsigned long long)((a.nvec + per * ctas - 1) / (per * ctas)); // steps of the busiest CTA, the same on both ranks
    k_handshake<<<1, 1, 0, st->stream>>>(&c->peer_ctrl[partner]->ready[c->rank], &c->ctrl->ready[partner], e, &c->ctrl->error);
    SPZ_CUDA(cudaEventRecord(ev.first, st->stream));
    {
        void *params[] = {&a};
        SPZ_CUDA(cudaLaunchKernel(kern, dim3((unsigned)ctas), dim3(THREADS), params, 0, st->stream));
    }
    SPZ_CUDA(cudaEventRecord(ev.second, st->stream));
    c->pending.push_back(ev);
    k_handshake<<<1, 1, 0, st->stream>>>(&c->peer_ctrl[partner]->done[c->rank], &c->ctrl->done[partner], e, &c->ctrl->error);
    count_launch(3);
    SPZ_CUDA(cudaGetLastError());
    c->n_exchanges += 1;
    c->bytes_sent += 16.0 * (double)(1ll << (n_local - 1)); // half a shard crosses the link in each direction, here as reads
    return SPZ_OK;
}

// May the gate be fused into the exchange that fetches its target?  Opted in, non-diagonal, vector path, and NO control at
// all in the logical op: a control that happened to be the evicted qubit would sit on the exchanged rank bit afterwards, one
// partner would skip the gate and the other apply it, and the two would run different kernels against each other.
bool dist_can_fuse_gate(const spz_state *st, int kind, uint64_t logical_cmask, int target, int lq) {
    if (!dist_fuse_gate_enabled() || logical_cmask != 0 || target != lq || lq < 2 || st->n < 3) return false;
    return kind == SPZ_GATE_H || kind == SPZ_GATE_X || kind == SPZ_GATE_Y || kind == SPZ_GATE_RX || kind == SPZ_GATE_RY || kind == SPZ_GATE_U;
}

// Make the main stream wait for an overlapped exchange still in flight (everything but the split tile pass needs
// the whole shard).
int dist_join(spz_state *st) {
    DistCtx *c = ctx_of(st);
    if (!c || !c->split_pending) return SPZ_OK;
    SPZ_CUDA(cudaStreamWaitEvent(st->stream, c->ev_chunk[c->split_chunks - 1], 0));
    c->split_pending = false;
    return SPZ_OK;
}

// For the consumer that directly follows an overlapped exchange.  The exchange lands in *n_chunks chunks; when they are
// contiguous ranges of the shard (the traded bit lies below the chunk bits) chunk k is amplitudes [k, k + 1) * len / n_chunks and
// ev[k] fires when it is complete on this rank, so the caller may work on it while the later ones are on the wire.  Returns false
// when nothing is in flight.  When the chunks are not contiguous ranges *n_chunks is 1 and ev[0] is the last event (a plain join).
bool dist_take_chunks(spz_state *st, int *n_chunks, cudaEvent_t *ev) {
    DistCtx *c = ctx_of(st);
    if (!c || !c->split_pending) return false;
    int K = c->split_chunks, bits = 0;
    while ((1 << bits) < K) ++bits;
    if (c->split_lq >= st->n - bits) { // the traded bit is one of the top bits: report the whole exchange as one chunk
        *n_chunks = 1;
        ev[0] = c->ev_chunk[K - 1];
    } else {
        *n_chunks = K;
        for (int k = 0; k < K; ++k) ev[k] = c->ev_chunk[k];
    }
    c->split_pending = false;
    return true;
}

int dist_allreduce(spz_state *st, const double *dev_value, double host_value, double *host_out, double **dev_out) {
    DistCtx *c = ctx_of(st);
    SPZ_TRY(join_pending(st));
    if (!c->connected) { set_error("dist state used before spz_dist_connect"); return SPZ_ERR_COMM; }
    SPZ_TRY(ensure_scratch(st));
    const unsigned long long e = ++c->red_epoch;
    if (!c->d_table) { set_error("internal: peer table missing"); return SPZ_ERR_COMM; }
    k_allreduce<<<1, 32, 0, st->stream>>>(c->ctrl, c->d_table, c->rank, c->world, e, dev_value, host_value, st->scratch.partials + 0);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    if (dev_out) *dev_out = st->scratch.partials + 0;
    if (host_out) {
        SPZ_CUDA(cudaMemcpyAsync(st->scratch.h_result, st->scratch.partials + 0, sizeof(double), cudaMemcpyDeviceToHost, st->stream));
        SPZ_CUDA(cudaStreamSynchronize(st->stream));
        *host_out = st->scratch.h_result[0];
        SPZ_TRY(check_comm_error(st));
    }
    return SPZ_OK;
}

int diag_const_on(spz_state *st, double *re, double *im, long long len, const GateK &g, uint64_t cmask, int hi) {
    const long long count = len >> __builtin_popcountll(cmask);
    const int grid = (int)std::max<long long>(1, std::min<long long>((count + 255) / 256, 148 * 16));
    k_diag_const<<<grid, 256, 0, st->stream>>>(re, im, count, cmask, g.kind, hi, g);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

int dist_diag_const(spz_state *st, const GateK &g, uint64_t local_cmask, int hi) {
    int rc = SPZ_OK;
    if (lanes_diag_const(st, g, local_cmask, hi, &rc)) return rc; // a streamed round trip: piece by piece
    SPZ_TRY(join_pending(st));
    return diag_const_on(st, st->re, st->im, st->len, g, local_cmask, hi);
}

// Lower one logical gate and run the resulting actions immediately (the unfused path).
int dist_apply_masked(spz_state *st, int kind, const double *p, int t0, int t1, uint64_t cmask, int target) {
    DistCtx *c = ctx_of(st);
    c->basis_hint = false;
    std::vector<spz_dist_action> acts;
    int rc = c->plan.lower(c->rank, kind, p, t0, t1, cmask, target, nullptr, acts);
    if (rc != SPZ_OK) { set_error("cannot lower gate kind %d target %d onto the sharded register", kind, target); return rc; }
    for (size_t ai = 0; ai < acts.size(); ++ai) {
        const spz_dist_action &a = acts[ai];
        // exchange immediately followed by the uncontrolled gate that asked for it: one fused kernel (opt-in)
        if (a.type == ACT_EXCHANGE && ai + 1 < acts.size() && acts[ai + 1].type == ACT_LOCAL_GATE &&
            dist_can_fuse_gate(st, acts[ai + 1].kind, cmask, acts[ai + 1].target, a.lq)) {
            GateK g;
            SPZ_TRY(resolve_gate(acts[ai + 1].kind, acts[ai + 1].p, &g));
            SPZ_TRY(dist_exchange_gate(st, a.gbit, a.lq, g));
            ++ai;
            continue;
        }
        switch (a.type) {
        case ACT_SKIP: break;
        case ACT_EXCHANGE: SPZ_TRY(dist_exchange(st, a.gbit, a.lq)); break;
        case ACT_LOCAL_GATE: { GateK g; SPZ_TRY(resolve_gate(a.kind, a.p, &g)); SPZ_TRY(launch_gate(st, g, a.cmask, a.target)); break; }
        case ACT_DIAG_CONST: { GateK g; SPZ_TRY(resolve_gate(a.kind, a.p, &g)); SPZ_TRY(dist_diag_const(st, g, a.cmask, a.hi)); break; }
        default: return SPZ_ERR_INVALID_ARG;
        }
    }
    return SPZ_OK;
}

int dist_lower(spz_state *st, int kind, const double *p, int t0, int t1, uint64_t cmask, int target, const uint64_t *next_use,
               std::vector<spz_dist_action> &acts) {
    DistCtx *c = ctx_of(st);
    c->basis_hint = false; // whatever is lowered here will be applied
    return c->plan.lower(c->rank, kind, p, t0, t1, cmask, target, next_use, acts);
}

// Reductions on a sharded register.  mode as in reduce_scalar(); `target` is a LOGICAL qubit.
int dist_reduce_scalar(spz_state *st, int mode, int target, double *out) {
    DistCtx *c = ctx_of(st);
    double local = 0.0;
    if (mode == 1) {
        SPZ_TRY(reduce_scalar(st, 1, 0, &local));
    } else {
        if (target < 0 || target >= c->plan.n) { set_error("target %d out of range", target); return SPZ_ERR_INVALID_ARG; }
        if (mode == 2 || mode == 3) { // <X>, <Y> pair amplitudes across the target bit: make it resident first
            std::vector<spz_dist_action> acts;
            SPZ_TRY(c->plan.ensure_local(target, nullptr, 0, acts, c->rank));
            for (const spz_dist_action &a : acts) SPZ_TRY(dist_exchange(st, a.gbit, a.lq));
        }
        const int pt = c->plan.perm[target];
        if (pt < c->plan.n_local) {
            SPZ_TRY(reduce_scalar(st, mode, pt, &local));
        } else { // global target: prob0 / <Z> are per-rank constants times the shard's norm
            const int bit = (c->rank >> (pt - c->plan.n_local)) & 1;
            double nrm = 0.0;
            SPZ_TRY(reduce_scalar(st, 1, 0, &nrm));
            if (mode == 0) local = bit ? 0.0 : nrm;
            else local = bit ? -nrm : nrm; // mode 4
        }
    }
    return dist_allreduce(st, nullptr, local, out, nullptr);
}

// <Z_t> for a list of targets: ONE read pass over the shard (reduce_z_all) serves every local target, a global target is the
// shard's norm with the sign of this rank's bit; then one scalar all-reduce per target.
int dist_reduce_z_multi(spz_state *st, const int32_t *targets, int k, double *out) {
    DistCtx *c = ctx_of(st);
    for (int i = 0; i < k; ++i)
        if (targets[i] < 0 || targets[i] >= c->plan.n) { set_error("target %d out of range", targets[i]); return SPZ_ERR_INVALID_ARG; }
    double loc[kZMaxBits + 1];
    SPZ_TRY(reduce_z_all(st, loc));
    for (int i = 0; i < k; ++i) {
        const int pt = c->plan.perm[targets[i]];
        double local;
        if (pt < c->plan.n_local) local = loc[0] - 2.0 * loc[1 + pt];
        else local = ((c->rank >> (pt - c->plan.n_local)) & 1) ? -loc[0] : loc[0];
        SPZ_TRY(dist_allreduce(st, nullptr, local, &out[i], nullptr));
    }
    return SPZ_OK;
}

// ---- free placement of a basis state ----------------------------------------------------------------------------------
// While the register is a computational basis state (|0..0> after State::new, spz_set_basis) the assignment of logical qubits to
// physical bits is free: relabelling costs two scalar writes (the single 1 moves).  execute() therefore picks the permutation
// from its op list before the first gate: the g qubits whose first non-diagonal use comes LATEST become the rank bits, so that
// everything before runs without communication and only those g qubits ever need an exchange (QFT-36 on 8 GPUs: 3 exchanges
// instead of 4; QFT-34 on 2: 1 instead of 2).  The hint that the state is a basis state is never trusted: every rank checks
// its shard (norm exactly 1 with the amplitude exactly 1 on the owner, norm exactly 0 elsewhere: one read pass) and the
// ranks agree through an all-reduce; on any doubt nothing is relabelled.  SPZ_DIST_PLACE=0 switches it off.
// first_use[q]: index of the first op with a non-diagonal action on logical qubit q (INT64_MAX: none).  dry: planning only
// (spz_debug_compile_sharded; the hint comes from SPZ_DEBUG_BASIS there).  *changed: the permutation is a new one.
static __global__ void k_store_double(double *p, double v) { *p = v; }

int dist_place_basis(spz_state *st, const int64_t *first_use, bool dry, bool *changed) {
    DistCtx *c = ctx_of(st);
    DistPlan &pl = c->plan;
    *changed = false;
    if (pl.g == 0) return SPZ_OK;
    if (const char *e = std::getenv("SPZ_DIST_PLACE")) if (e[0] == '0') return SPZ_OK;
    uint64_t x = c->basis_index;
    if (dry) {
        const char *e = std::getenv("SPZ_DEBUG_BASIS");
        if (!e) return SPZ_OK;
        x = std::strtoull(e, nullptr, 0);
    } else if (!c->basis_hint || !c->connected) {
        return SPZ_OK;
    }
    // the g logical qubits with the latest first use; among equals those that are global already, then the highest
    std::vector<int> order(pl.n);
    for (int q = 0; q < pl.n; ++q) order[q] = q;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (first_use[a] != first_use[b]) return first_use[a] > first_use[b];
        const bool ga = pl.is_global_phys(pl.perm[a]), gb = pl.is_global_phys(pl.perm[b]);
        if (ga != gb) return ga;
        return a > b;
    });
    uint64_t want = 0, have = 0;
    for (int k = 0; k < pl.g; ++k) want |= 1ull << order[k];
    for (int q = 0; q < pl.n; ++q) if (pl.is_global_phys(pl.perm[q])) have |= 1ull << q;
    // a qubit that is global now and first used no later than every wanted one gains nothing by moving: keep the set unless
    // it really postpones a use
    int64_t worst_have = INT64_MAX, worst_want = INT64_MAX;
    for (int q = 0; q < pl.n; ++q) {
        if ((have >> q) & 1ull) worst_have = std::min(worst_have, first_use[q]);
        if ((want >> q) & 1ull) worst_want = std::min(worst_want, first_use[q]);
    }
    if (want == have || worst_want <= worst_have) return SPZ_OK;
    auto phys_of = [&](uint64_t logical) {
        uint64_t p = 0;
        for (int q = 0; q < pl.n; ++q) if ((logical >> q) & 1ull) p |= 1ull << pl.perm[q];
        return p;
    };
    const uint64_t old_phys = phys_of(x);
    const int old_owner = (int)(old_phys >> pl.n_local);
    const uint64_t local_mask = (1ull << pl.n_local) - 1ull;
    if (!dry) { // is it really |x>?  (every rank takes part in the all-reduce, whatever it found)
        double nrm = -1.0, amp[2] = {0.0, 0.0};
        SPZ_TRY(reduce_scalar(st, 1, 0, &nrm));
        bool ok;
        if (old_owner == c->rank) {
            SPZ_CUDA(cudaMemcpyAsync(&amp[0], st->re + (old_phys & local_mask), sizeof(double), cudaMemcpyDeviceToHost, st->stream));
            SPZ_CUDA(cudaMemcpyAsync(&amp[1], st->im + (old_phys & local_mask), sizeof(double), cudaMemcpyDeviceToHost, st->stream));
            SPZ_CUDA(cudaStreamSynchronize(st->stream));
            ok = nrm == 1.0 && amp[0] == 1.0 && amp[1] == 0.0;
        } else {
            ok = nrm == 0.0;
        }
        double bad = 0.0;
        SPZ_TRY(dist_allreduce(st, nullptr, ok ? 0.0 : 1.0, &bad, nullptr));
        if (bad != 0.0) { c->basis_hint = false; return SPZ_OK; }
    }
    // relabel: wanted qubits that are local trade places with global qubits that are not wanted
    std::vector<int> in_q, out_q;
    for (int q = 0; q < pl.n; ++q) {
        if (((want >> q) & 1ull) && !((have >> q) & 1ull)) in_q.push_back(q);
        if (((have >> q) & 1ull) && !((want >> q) & 1ull)) out_q.push_back(q);
    }
    for (size_t i = 0; i < in_q.size() && i < out_q.size(); ++i) {
        const int a = in_q[i], b = out_q[i], pa = pl.perm[a], pb = pl.perm[b];
        pl.perm[a] = pb; pl.inv[pb] = a;
        pl.perm[b] = pa; pl.inv[pa] = b;
    }
    *changed = true;
    if (dry) return SPZ_OK;
    const uint64_t new_phys = phys_of(x);
    const int new_owner = (int)(new_phys >> pl.n_local);
    if (new_phys != old_phys) {
        if (old_owner == c->rank) k_store_double<<<1, 1, 0, st->stream>>>(st->re + (old_phys & local_mask), 0.0);
        if (new_owner == c->rank) k_store_double<<<1, 1, 0, st->stream>>>(st->re + (new_phys & local_mask), 1.0);
        SPZ_CUDA(cudaGetLastError());
    }
    c->n_placements += 1;
    return SPZ_OK;
}

int dist_collapse(spz_state *st, int target, int outcome, double scale) {
    DistCtx *c = ctx_of(st);
    c->basis_hint = false;
    const int pt = c->plan.perm[target];
    if (pt < c->plan.n_local) return launch_collapse(st, pt, outcome, 0, scale);
    const int bit = (c->rank >> (pt - c->plan.n_local)) & 1;
    if (bit == outcome) return launch_scale(st, scale);
    SPZ_TRY(join_pending(st)); // the memsets below bypass the launch_* helpers: an overlapped exchange may still be writing
    SPZ_CUDA(cudaMemsetAsync(st->re, 0, sizeof(double) * (size_t)st->len, st->stream));
    SPZ_CUDA(cudaMemsetAsync(st->im, 0, sizeof(double) * (size_t)st->len, st->stream));
    return SPZ_OK;
}

int dist_fill_basis(spz_state *st, uint64_t logical_index) {
    DistCtx *c = ctx_of(st);
    if (c->plan.n < 64 && (logical_index >> c->plan.n)) { set_error("basis index out of range"); return SPZ_ERR_INVALID_ARG; }
    uint64_t phys = 0;
    for (int q = 0; q < c->plan.n; ++q) if ((logical_index >> q) & 1ull) phys |= 1ull << c->plan.perm[q];
    const int owner = (int)(phys >> c->plan.n_local);
    // Every rank joins first: an overlapped exchange (second stream) may still be moving amplitudes of this shard, and the
    // memsets of the non-owner ranks do not go through a launch_* helper that would wait for it.
    SPZ_TRY(join_pending(st));
    c->basis_hint = false;
    if (owner == c->rank) {
        SPZ_TRY(launch_fill_basis(st, phys & ((1ull << c->plan.n_local) - 1ull)));
    } else {
        SPZ_CUDA(cudaMemsetAsync(st->re, 0, sizeof(double) * (size_t)st->len, st->stream));
        SPZ_CUDA(cudaMemsetAsync(st->im, 0, sizeof(double) * (size_t)st->len, st->stream));
    }
    c->basis_hint = true;
    c->basis_index = logical_index;
    return SPZ_OK;
}

int dist_init_random(spz_state *st, uint64_t seed) {
    DistCtx *c = ctx_of(st);
    c->basis_hint = false;
    const long long total_len = 1ll << c->plan.n;
    double *d_local = nullptr, *d_total = nullptr;
    SPZ_TRY(launch_rand_probs(st, seed, (long long)c->rank * st->len, &d_local));
    SPZ_TRY(dist_allreduce(st, d_local, 0.0, nullptr, &d_total));
    return launch_rand_finish(st, seed, (long long)c->rank * st->len, total_len, d_total);
}

// Sampling a sharded register: rank masses are all-gathered (one scalar all-reduce per rank, so every rank holds the
// bitwise-identical table), each shot is routed to the rank whose mass interval contains u * total, and that rank
// runs the single-GPU sampler on its shard with the rescaled uniform.  The CDF therefore walks PHYSICAL index order
// (rank-major); the outcome is reported as a logical basis index.  Shots owned by other ranks are returned as -1.
int dist_sample(spz_state *st, const double *u01, int64_t shots, int64_t *out_index) {
    DistCtx *c = ctx_of(st);
    double mine = 0.0;
    SPZ_TRY(reduce_scalar(st, 1, 0, &mine));
    double mass[kMaxRanks];
    for (int r = 0; r < c->world; ++r) SPZ_TRY(dist_allreduce(st, nullptr, r == c->rank ? mine : 0.0, &mass[r], nullptr));
    double cum[kMaxRanks + 1];
    cum[0] = 0.0;
    for (int r = 0; r < c->world; ++r) cum[r + 1] = cum[r] + mass[r];
    const double total = cum[c->world];
    std::vector<double> lu;
    std::vector<int64_t> where;
    for (int64_t k = 0; k < shots; ++k) {
        double u = u01[k];
        if (!(u >= 0.0)) u = 0.0;
        if (u >= 1.0) u = 0x1.fffffffffffffp-1;
        const double x = u * total;
        int owner = 0;
        while (owner + 1 < c->world && x >= cum[owner + 1]) ++owner;
        out_index[k] = -1;
        if (owner != c->rank || mass[owner] <= 0.0) continue;
        double v = (x - cum[owner]) / mass[owner];
        if (!(v >= 0.0)) v = 0.0;
        if (v >= 1.0) v = 0x1.fffffffffffffp-1;
        lu.push_back(v);
        where.push_back(k);
    }
    if (lu.empty()) return SPZ_OK;
    std::vector<int64_t> local(lu.size());
    SPZ_TRY(launch_sample(st, lu.data(), (int64_t)lu.size(), local.data()));
    for (size_t i = 0; i < lu.size(); ++i) {
        const uint64_t phys = ((uint64_t)c->rank << c->plan.n_local) | (uint64_t)local[i];
        uint64_t logical = 0;
        for (int q = 0; q < c->plan.n; ++q) logical |= ((phys >> c->plan.perm[q]) & 1ull) << q;
        out_index[where[i]] = (int64_t)logical;
    }
    return SPZ_OK;
}

void dist_destroy(spz_state *st) {
    DistCtx *c = ctx_of(st);
    if (!c) return;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || !c->connected || !c->ipc) continue;
        if (c->peer_re[r]) cudaIpcCloseMemHandle(c->peer_re[r]);
        if (c->peer_im[r]) cudaIpcCloseMemHandle(c->peer_im[r]);
        if (c->peer_ctrl[r]) cudaIpcCloseMemHandle(c->peer_ctrl[r]);
    }
    for (auto &e : c->pending) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : c->free_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (c->xstream) { cudaStreamSynchronize(c->xstream); cudaStreamDestroy(c->xstream); }
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    for (int k = 0; k < kMaxChunks; ++k) if (c->ev_chunk[k]) cudaEventDestroy(c->ev_chunk[k]);
    cudaFree(c->d_table);
    cudaFree(c->ctrl);
    delete c;
    st->dist = nullptr;
}

} // namespace spz

using namespace spz;

struct spz_dist_plan {
    DistPlan plan;
};

extern "C" {

int spz_dist_create(int n_qubits, int rank, int world, int device, spz_state **out) {
    if (!out) return SPZ_ERR_INVALID_ARG;
    *out = nullptr;
    if (world < 1 || world > kMaxRanks || (world & (world - 1)) || rank < 0 || rank >= world) {
        set_error("world size must be a power of two <= %d and 0 <= rank < world", kMaxRanks); return SPZ_ERR_INVALID_ARG;
    }
    int g = 0;
    while ((1 << g) < world) ++g;
    if (n_qubits - g < 2 || n_qubits > 48) { set_error("need at least 2 local qubits per rank"); return SPZ_ERR_INVALID_ARG; }
    spz_state *st = nullptr;
    SPZ_TRY(spz_create(n_qubits - g, device, &st));
    DistCtx *c = new DistCtx();
    c->plan.init(n_qubits, world);
    c->rank = rank; c->world = world;
    st->dist = c;
    cudaError_t e = cudaMalloc(&c->ctrl, sizeof(CtrlBlock));
    if (e == cudaSuccess) e = cudaMemset(c->ctrl, 0, sizeof(CtrlBlock));
    if (e == cudaSuccess) { // the exchange stream outranks the compute stream: its few CTAs are placed first
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming);
    for (int k = 0; k < kMaxChunks; ++k) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_chunk[k], cudaEventDisableTiming);
    if (const char *v = getenv("SPZ_XCHG_TMA")) c->xchg_tma = v[0] == '1';
    if (const char *v = getenv("SPZ_XCHG_TMA_CTAS")) { const int k = atoi(v); if (k >= 1 && k <= 148) c->xchg_tma_ctas = k; }
    if (c->xchg_tma) cudaFuncSetAttribute(k_exchange_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kXtStages * 4u * kXtRun * 8u + 64));
    if (const char *v = getenv("SPZ_XCHG_CHUNKS")) { const int k = atoi(v); if (k == 1 || k == 2 || k == 4 || k == 8) c->n_chunks = k; }
    if (getenv("SPZ_NO_OVERLAP")) c->overlap = false;
    if (const char *v = getenv("SPZ_XCHG_CTAS")) { const int k = atoi(v); if (k >= 1 && k <= 1024) c->xchg_ctas = k; }
    if (const char *v = getenv("SPZ_XG_CTAS")) { const int k = atoi(v); if (k >= 1 && k <= kMaxXgCtas) c->xg_ctas = k; }
    if (e != cudaSuccess) { int rc = cuda_fail(e, "cudaMalloc(ctrl)", __FILE__, __LINE__); spz_destroy(st); return rc; }
    c->peer_ctrl[rank] = c->ctrl; c->peer_re[rank] = st->re; c->peer_im[rank] = st->im;
    if (world == 1) { c->connected = true; if (upload_peer_table(c) != SPZ_OK) { spz_destroy(st); return SPZ_ERR_CUDA; } }
    int rc = dist_fill_basis(st, 0); // |0..0> lives on rank 0 only
    if (rc != SPZ_OK) { spz_destroy(st); return rc; }
    *out = st;
    return SPZ_OK;
}

int spz_dist_export(spz_state *st, void *blob) {
    if (!st || !st->dist || !blob) return SPZ_ERR_INVALID_ARG;
    SPZ_CUDA(cudaSetDevice(st->device));
    DistCtx *c = ctx_of(st);
    IpcBlob b;
    std::memset(&b, 0, sizeof b);
    SPZ_CUDA(cudaIpcGetMemHandle(&b.re, st->re));
    SPZ_CUDA(cudaIpcGetMemHandle(&b.im, st->im));
    SPZ_CUDA(cudaIpcGetMemHandle(&b.ctrl, c->ctrl));
    b.rank = c->rank; b.device = st->device; b.len = st->len;
    std::memcpy(blob, &b, sizeof b);
    return SPZ_OK;
}

int spz_dist_connect(spz_state *st, const void *blobs) {
    if (!st || !st->dist || !blobs) return SPZ_ERR_INVALID_ARG;
    SPZ_CUDA(cudaSetDevice(st->device));
    DistCtx *c = ctx_of(st);
    const IpcBlob *b = static_cast<const IpcBlob *>(blobs);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        if (b[r].rank != r || b[r].len != st->len) { set_error("blob %d does not describe rank %d's shard", r, r); return SPZ_ERR_COMM; }
        void *p = nullptr;
        // control blocks of every rank (scalar all-reduce); amplitude arrays only of the hypercube partners
        // rank ^ 2^k -- mapping a 137 GB shard of all 7 peers costs tens of seconds and is never used
        SPZ_CUDA(cudaIpcOpenMemHandle(&p, b[r].ctrl, cudaIpcMemLazyEnablePeerAccess)); c->peer_ctrl[r] = static_cast<CtrlBlock *>(p);
        const int diff = r ^ c->rank;
        if ((diff & (diff - 1)) != 0) continue;
        SPZ_CUDA(cudaIpcOpenMemHandle(&p, b[r].re, cudaIpcMemLazyEnablePeerAccess)); c->peer_re[r] = static_cast<double *>(p);
        SPZ_CUDA(cudaIpcOpenMemHandle(&p, b[r].im, cudaIpcMemLazyEnablePeerAccess)); c->peer_im[r] = static_cast<double *>(p);
    }
    SPZ_TRY(upload_peer_table(c));
    c->connected = true;
    c->ipc = true;
    return SPZ_OK;
}

int spz_dist_connect_local(spz_state **states, int world) {
    if (!states || world < 1 || world > kMaxRanks) return SPZ_ERR_INVALID_ARG;
    for (int r = 0; r < world; ++r) {
        if (!states[r] || !states[r]->dist) { set_error("state %d is not a shard", r); return SPZ_ERR_INVALID_ARG; }
        DistCtx *c = ctx_of(states[r]);
        if (c->rank != r || c->world != world || states[r]->len != states[0]->len) { set_error("state %d: wrong rank/world/length", r); return SPZ_ERR_INVALID_ARG; }
    }
    for (int r = 0; r < world; ++r) {
        DistCtx *c = ctx_of(states[r]);
        SPZ_CUDA(cudaSetDevice(states[r]->device));
        for (int q = 0; q < world; ++q) {
            if (q == r) continue;
            if (states[q]->device != states[r]->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(states[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                cudaGetLastError();
            }
            c->peer_re[q] = states[q]->re; c->peer_im[q] = states[q]->im; c->peer_ctrl[q] = ctx_of(states[q])->ctrl;
        }
        SPZ_TRY(upload_peer_table(c));
        c->connected = true;
        c->ipc = false;
    }
    return SPZ_OK;
}

int spz_dist_perm(const spz_state *st, int32_t *perm_out) {
    if (!st || !st->dist || !perm_out) return SPZ_ERR_INVALID_ARG;
    const DistCtx *c = ctx_of(st);
    for (int q = 0; q < c->plan.n; ++q) perm_out[q] = c->plan.perm[q];
    return SPZ_OK;
}

int spz_dist_local_qubits(const spz_state *st) { return st ? st->n : -1; }

// The Clone of a sharded register (core.rs:18 #[derive(Clone)]) is a collective: every rank creates and connects a second
// register (spz_dist_create / export / connect), then copies its own shard and the shared plan state into it.
int spz_dist_copy_from(spz_state *dst, const spz_state *csrc) {
    spz_state *src = const_cast<spz_state *>(csrc);
    if (!dst || !src || !dst->dist || !src->dist) { set_error("both handles must be shards"); return SPZ_ERR_INVALID_ARG; }
    DistCtx *d = ctx_of(dst), *s = ctx_of(src);
    d->basis_hint = false;
    if (dst->len != src->len || d->plan.n != s->plan.n || d->world != s->world || d->rank != s->rank) {
        set_error("shards of different registers or ranks"); return SPZ_ERR_INVALID_ARG;
    }
    SPZ_CUDA(cudaSetDevice(dst->device));
    SPZ_TRY(join_pending(const_cast<spz_state *>(src)));
    SPZ_TRY(join_pending(dst));
    SPZ_CUDA(cudaStreamSynchronize(src->stream));
    const size_t bytes = sizeof(double) * (size_t)src->len;
    SPZ_CUDA(cudaMemcpyAsync(dst->re, src->re, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    SPZ_CUDA(cudaMemcpyAsync(dst->im, src->im, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    SPZ_CUDA(cudaStreamSynchronize(dst->stream));
    const DistPlan &sp = s->plan;
    DistPlan &dp = d->plan;
    for (int q = 0; q < 64; ++q) { dp.perm[q] = sp.perm[q]; dp.inv[q] = sp.inv[q]; dp.last_use[q] = sp.last_use[q]; }
    dp.clock = sp.clock;
    dst->rng = src->rng;
    return SPZ_OK;
}

int spz_dist_stats(const spz_state *cst, double *out4) {
    if (!cst || !cst->dist || !out4) return SPZ_ERR_INVALID_ARG;
    spz_state *st = const_cast<spz_state *>(cst);
    DistCtx *c = ctx_of(st);
    SPZ_CUDA(cudaSetDevice(st->device));
    SPZ_TRY(join_pending(st)); // an overlapped exchange may still be running on the second stream
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    for (auto &e : c->pending) {
        float ms = 0.f;
        SPZ_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
        c->ms_accum += ms;
        c->free_events.push_back(e);
    }
    c->pending.clear();
    out4[0] = c->n_exchanges; out4[1] = c->bytes_sent; out4[2] = c->ms_accum; out4[3] = c->n_overlapped;
    return SPZ_OK;
}

int spz_dist_plan_create(int n_qubits, int world, spz_dist_plan **out) {
    if (!out || world < 1 || (world & (world - 1)) || n_qubits < 1 || n_qubits > 63) return SPZ_ERR_INVALID_ARG;
    spz_dist_plan *p = new spz_dist_plan();
    p->plan.init(n_qubits, world);
    if (p->plan.n_local < 1) { delete p; return SPZ_ERR_INVALID_ARG; }
    *out = p;
    return SPZ_OK;
}

int spz_dist_plan_destroy(spz_dist_plan *p) { delete p; return SPZ_OK; }

int spz_dist_plan_lower(spz_dist_plan *p, int rank, const spz_op *op, spz_dist_action *out, int max_out, int *n_out) {
    if (!p || !op || !out || !n_out) return SPZ_ERR_INVALID_ARG;
    uint64_t cmask = 0;
    switch (op->ctrl_kind) { // same resolution as spz_execute (circuit.rs:567-596)
    case SPZ_CTRL_NONE: break;
    case SPZ_CTRL_SINGLE: case SPZ_CTRL_ONES: cmask = op->ctrl_mask; break;
    case SPZ_CTRL_MIXED: cmask = op->ctrl_mask & ~op->zeros_mask; break;
    case SPZ_CTRL_SIGNED: // spz_execute lowers negative controls to X . op . X (three ops): expand before planning op by op
        if (op->zeros_mask) return SPZ_ERR_UNSUPPORTED;
        cmask = op->ctrl_mask;
        break;
    default: return SPZ_ERR_INVALID_ARG;
    }
    if (op->kind == SPZ_GATE_M || op->kind == SPZ_GATE_UNITARY || op->kind == SPZ_GATE_BITFLIP) return SPZ_ERR_UNSUPPORTED;
    std::vector<spz_dist_action> acts;
    SPZ_TRY(p->plan.lower(rank, op->kind, op->p, op->t0, op->t1, cmask, op->target, nullptr, acts));
    if ((int)acts.size() > max_out) return SPZ_ERR_INVALID_ARG;
    for (size_t i = 0; i < acts.size(); ++i) out[i] = acts[i];
    *n_out = (int)acts.size();
    return SPZ_OK;
}

int spz_dist_plan_perm(const spz_dist_plan *p, int32_t *perm_out) {
    if (!p || !perm_out) return SPZ_ERR_INVALID_ARG;
    for (int q = 0; q < p->plan.n; ++q) perm_out[q] = p->plan.perm[q];
    return SPZ_OK;
}

} // extern "C"
