// engine.h -- internal declarations shared by the CUDA translation units of libspinoza_b200.
// Nothing here is part of the ABI (see include/spinoza_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spinoza_b200.h"

namespace spz {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SPZ_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return ::spz::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define SPZ_TRY(call)                   \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != SPZ_OK) return rc__; \
    } while (0)

void count_launch(int n = 1);

// ---- gates resolved to scalars (host side) ----------------------------------------------------
// Scalars are computed once on the host exactly as each reference *_apply does, so that the device
// arithmetic can mirror the reference operation by operation:
//   P  : s0=cos(theta) s1=sin(theta)                         gates.rs:865
//   RX : s0=cos(theta/2) s1=-sin(theta/2)                    gates.rs:749-751
//   RY : s0=sin(theta/2) s1=cos(theta/2)                     gates.rs:1125-1126
//   RZ : s0=cos(theta/2) s1=sin(theta/2)  (d0=(c,-s) d1=(c,s)) gates.rs:970-973
//   U  : s0..s6 = a,k,l,q,r,s,t                              gates.rs:1286-1304
struct GateK {
    int kind;
    double s[7];
};
int resolve_gate(int kind, const double *p, GateK *out); // SPZ_ERR_UNSUPPORTED for M/SWAP/UNITARY/BITFLIP

inline bool is_diagonal_kind(int kind) { return kind == SPZ_GATE_Z || kind == SPZ_GATE_P || kind == SPZ_GATE_RZ; }

// ---- the state --------------------------------------------------------------------------------
struct Scratch {
    double *partials = nullptr; // device, reduction partials
    size_t n_partials = 0;
    double *h_result = nullptr; // pinned host, small
    void *samp = nullptr;       // sampler arena (stream-ordered allocation, grown on demand)
    size_t samp_bytes = 0;
};

} // namespace spz

struct spz_state {
    int n = 0;           // qubits held by this handle: the whole register, or the local qubits of a shard (len == 1 << n)
    int device = 0;
    int64_t len = 0;     // amplitudes held by this handle (2^n single-GPU)
    double *re = nullptr;
    double *im = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    spz::Scratch scratch;
    uint64_t rng = 0x853c49e6748fea9bULL; // splitmix64 state for M / BitFlipNoise draws
    // fused-execute scratch
    void *d_ops = nullptr;
    size_t d_ops_bytes = 0;
    size_t d_ops_cursor = 0;
    void *dist = nullptr; // spz::DistCtx* when this handle is one shard of a multi-GPU register (dist.cu)
    // A whole-state upload in flight on the copy stream (spz_upload_async): the state arrives in `chunks` contiguous pieces,
    // ev[k] fires when piece k is complete.  The pass that follows runs piece by piece behind the bus (see take_chunks).
    struct Arrival {
        bool pending = false;
        int chunks = 0;
        cudaStream_t copy = nullptr;
        cudaEvent_t ev[8] = {};
        cudaEvent_t ready = nullptr; // main stream -> copy stream: everything queued before the upload has finished
        // One stream per piece: a RUN of one-gate passes behind the upload is issued gate by gate, but must execute piece by
        // piece (all gates on piece 0 while piece 1 is on the bus, ...), which a single stream cannot express.
        cudaStream_t lane[8] = {};   // lane 0 has the highest priority: it finishes its run first, so its piece can leave first
        bool lanes_active = false;
        // "Streaming" between spz_upload_async and the spz_download that ends the round trip: runs of gates that act inside
        // the pieces keep going to the lanes (also after a join), so that the download of piece 0 can start while the other
        // pieces are still being computed.
        bool streaming = false;
    } arrival;
};

namespace spz {

// ---- kernel launchers (kernels_direct.cu) -------------------------------------------------------
// Apply resolved gate g to `target` for every amplitude whose index has all bits of ctrl_mask set.
int launch_gate(spz_state *st, const GateK &g, uint64_t ctrl_mask, int target);
// The same with signed controls: the qubits of neg_mask (a subset of ctrl_mask) must be 0 instead of 1 (extension).
int launch_gate_signed(spz_state *st, const GateK &g, uint64_t ctrl_mask, uint64_t neg_mask, int target);
int launch_swap(spz_state *st, int t0, int t1);

// ---- reductions and state utilities (kernels_reduce.cu) -----------------------------------------
int ensure_scratch(spz_state *st);
int launch_fill_basis(spz_state *st, uint64_t index);
int launch_init_random(spz_state *st, uint64_t seed);
// mode: 0 = sum |amp|^2 over amps with bit `target` clear (prob0); 1 = all amps (norm2; target ignored)
// 2/3/4 = <X>/<Y>/<Z> on `target`
int reduce_scalar(spz_state *st, int mode, int target, double *out);
// every qubit at once, one read pass: out[0] = sum |amp|^2, out[1 + t] = the mass at indices with bit t set (out: n + 1 doubles)
int reduce_z_all(spz_state *st, double *out);
// <X> (obs 0) or <Y> (obs 1) of k targets of a single-GPU register, up to twelve targets per read pass (kernels_xyall.cuh)
int reduce_xy_multi(spz_state *st, int obs, const int32_t *targets, int k, double *out);
int launch_collapse(spz_state *st, int target, int outcome, int reset, double scale);
// the same two steps of a measurement restricted to the indices where every qubit of zero_mask is 0 (the rest is known to be 0)
int reduce_prob0_sub(spz_state *st, int target, uint64_t zero_mask, double *out);
int launch_collapse_sub(spz_state *st, int target, int outcome, int reset, double scale, uint64_t zero_mask);
int launch_scale(spz_state *st, double scale); // every amplitude *= scale
// gen_random_state split in two so that a sharded register can insert its cross-rank sum between the halves
int launch_rand_probs(spz_state *st, uint64_t seed, long long index_offset, double **d_local_total);
int launch_rand_finish(spz_state *st, uint64_t seed, long long index_offset, long long total_len, const double *d_total);
int launch_sample(spz_state *st, const double *u01, int64_t shots, int64_t *out_index);

// ---- fused execution (kernels_tile.cu) ------------------------------------------------------------
// Micro-program of the fused tile kernel (kernels_tile.cu).  128 bytes per instruction.
enum { TI_LAYOUT = 0, TI_GATE = 1, TI_DIAG = 2, TI_RUN = 3 };

// One phase term of a merged run of diagonal gates (32 bytes): amplitudes whose tile base contains `outer`,
// whose thread bits contain `thr` and whose register index contains `m` are multiplied by (fr, fi).
struct TileTerm {
    uint64_t outer;
    uint32_t thr;
    uint32_t m;
    double fr, fi;
};
// Terms of one run that share (thr, m) form a group: their product depends only on the tile (through `outer`),
// so it is computed once per CTA and shared through shared memory.
struct TileGroup {
    uint32_t thr;
    uint32_t m;
    int first; // first term
    int count;
};
constexpr int kMaxTileGroups = 2048; // shared-memory budget: 20 bytes per group
struct TileInstr {
    // ---- chunk 0 (16 B): read for every instruction ----
    int op;               // TI_*
    int kind;             // spz_gate_kind
    int rpos;             // TI_GATE: which register bit (0..3) the target is.  TI_RUN: first group of the run
    uint32_t t_mask;      // TI_DIAG: t_where 1: tile-index mask; t_where 2: mask over k.  TI_GATE: set of k with controls set
    // ---- chunk 1 (16 B) ----
    uint32_t reg_cmask;   // controls on register bits (mask over k = 0..15)
    uint32_t thr_cmask;   // controls on thread bits (mask in tile-index space)
    int t_where;          // TI_DIAG target: 0 = outside the tile, 1 = thread bit, 2 = register bit
    int outer_target;     // t_where 0: absolute qubit, unless const_hi is set
    // ---- chunk 2 (16 B) ----
    // TI_LAYOUT: the 4 register-resident tile bits, ascending.
    // TI_RUN: rbit[0..3], reg_cmask, thr_cmask hold the GROUP counts of classes m = 0, 1, 2, 4, 8, other (in that
    // order); diagonal gates commute, so the host sorts a run's terms freely.
    int rbit[4];
    // ---- chunk 3 (16 B) ----
    uint64_t outer_cmask; // controls outside the tile (absolute positions)
    uint32_t const_hi;    // t_where 0: 0 = read the bit from the tile base, 1 = bit is 0, 2 = bit is 1 (a rank bit)
    int has_f0;           // merged mode: f0 is not the identity
    // ---- payload ----
    double s[7];          // gate scalars for the exact arithmetic (gate_math.cuh)
    double f0[2];         // merged mode: factor when all controls are set (RZ's d0)
    double f1[2];         // merged mode: extra factor when the target bit is set too (e^{i theta}, -1)
    double pad_;          // keeps the size a multiple of 16 bytes (staged with 128-bit copies)
};
static_assert(sizeof(TileInstr) % 16 == 0, "TileInstr is copied to shared memory in 16-byte units");
struct TilePlan {
    int tile_bits;       // T
    int low_bits;        // L: tile bits 0..L-1 are qubits 0..L-1
    int n_high;          // tile bits L.. are qubits high[0..n_high)
    int high[16];
};
int launch_tile_program(spz_state *st, const TilePlan &plan, const TileInstr *prog, int n_instr, const TileGroup *groups,
                        int n_groups, const TileTerm *terms, int n_terms, bool exact);
int max_tile_bits();
int min_tile_bits();
int tile_prepare(spz_state *st); // allocate the program ring buffer, set the kernel's shared-memory limit
// The program ring buffer of a state (device memory; a pass's program must stay intact until its kernel has run)
int tile_ring_alloc(spz_state *st, size_t bytes, char **slot);
// kernels_tile3.cu: the TMA tile kernel of merged mode (default; SPZ_TILE_V3=0 keeps every pass on k_tile)
struct Tile3Launch {
    alignas(64) unsigned char args[512]; // Tile3Args (holds two CUtensorMap)
    size_t smem;
};
bool tile3_enabled();
int prepare_tile3(spz_state *st, const TilePlan &plan, const TileInstr *prog, int n_instr, const TileGroup *groups, int n_groups,
                  const TileTerm *terms, int n_terms, Tile3Launch *out, bool *handled);
void run_tile3(spz_state *st, const Tile3Launch &l, unsigned first, unsigned count);

// ---- multi-GPU (dist.cu) ----------------------------------------------------------------------------
int dist_total_qubits(const spz_state *st);
int dist_apply_masked(spz_state *st, int kind, const double *p, int t0, int t1, uint64_t cmask, int target);
int dist_exchange(spz_state *st, int gbit, int lq);
// opt-in (SPZ_DIST_FUSE_GATE=1): the exchange and the uncontrolled non-diagonal gate on the arriving qubit in one kernel
bool dist_can_fuse_gate(const spz_state *st, int kind, uint64_t logical_cmask, int target, int lq);
int dist_exchange_gate(spz_state *st, int gbit, int lq, const GateK &g);
int dist_join(spz_state *st); // main stream waits for an overlapped exchange still in flight
bool dist_take_chunks(spz_state *st, int *n_chunks, cudaEvent_t *ev); // ev: room for 8 (see dist.cu)
int arrival_join(spz_state *st); // main stream waits for an asynchronous upload still in flight (abi.cu)
inline int join_pending(spz_state *st) {
    if (st->arrival.pending || st->arrival.lanes_active) { const int rc = arrival_join(st); if (rc != SPZ_OK) return rc; }
    return st->dist ? dist_join(st) : SPZ_OK;
}
// For the pass that directly follows a chunked arrival -- an asynchronous upload or an overlapped exchange: the state lands in
// *n_chunks contiguous pieces, ev[k] fires when piece k is complete (ev: room for 8).  False when nothing is in flight.
bool take_chunks(spz_state *st, int *n_chunks, cudaEvent_t *ev);
int dist_diag_const(spz_state *st, const GateK &g, uint64_t local_cmask, int hi);
bool lanes_diag_const(spz_state *st, const GateK &g, uint64_t local_cmask, int hi, int *rc); // kernels_direct.cu (streaming)
// the same constant-factor pass on one contiguous piece of the state (launched on st->stream; no join)
int diag_const_on(spz_state *st, double *re, double *im, long long len, const GateK &g, uint64_t cmask, int hi);
int dist_reduce_scalar(spz_state *st, int mode, int target, double *out);
int dist_reduce_z_multi(spz_state *st, const int32_t *targets, int k, double *out); // <Z> of k targets from one pass per shard
constexpr int kZMaxBits = 40; // reduce_z_all / kernels_zall.cuh
int dist_collapse(spz_state *st, int target, int outcome, double scale);
int dist_fill_basis(spz_state *st, uint64_t logical_index);
int dist_init_random(spz_state *st, uint64_t seed);
int dist_sample(spz_state *st, const double *u01, int64_t shots, int64_t *out_index);
// execute(): choose the qubit permutation of a register that is still a basis state from the op list (see dist.cu)
int dist_place_basis(spz_state *st, const int64_t *first_use, bool dry, bool *changed);
void dist_note_modified(spz_state *st); // the amplitudes were written by something dist.cu does not see (an upload)
void dist_destroy(spz_state *st);
void dist_debug_attach(spz_state *st, int n_total, int world, int rank); // host-only plan context for dry runs
void dist_debug_detach(spz_state *st);

} // namespace spz
