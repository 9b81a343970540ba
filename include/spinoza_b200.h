/*
 * spinoza_b200.h -- C ABI of the B200-native state-vector engine for Spinoza's gate path.
 *
 * The reference (QuState/spinoza) has no FFI seam: its boundary is the crate's public Rust API
 * (SURVEY.md section 8b).  Each entry point below names the reference item it replaces
 * (paths relative to /root/reference/spinoza/src/).  The Rust shim (rust/spinoza-b200), the C++
 * mirror (spinoza_b200/cpp/spinoza.hpp) and the Python mirror (spinoza_b200/__init__.py) bind
 * exactly these symbols; INTEGRATION.md shows the binding a Spinoza maintainer would add.
 *
 * Conventions
 *   - qubit t <-> bit t of the amplitude index; |0..0> is index 0 (core.rs:36).
 *   - amplitudes are split re/im f64 arrays (core.rs:20-24), resident in device memory.
 *   - every function returns an spz_status; nothing throws or aborts across the ABI
 *     (the reference panics: gates.rs:230,267,275,318; circuit.rs:597).
 *   - a handle is used by one caller at a time (like `&mut State`); work is stream-ordered on the
 *     state's stream, host-visible results (download, reductions, measure) synchronise.
 *   - there is NO CPU fallback: without a CUDA device spz_create fails with SPZ_ERR_NO_DEVICE.
 */
#ifndef SPINOZA_B200_H
#define SPINOZA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPZ_ABI_VERSION 1

#if defined(__GNUC__)
#define SPZ_API __attribute__((visibility("default")))
#else
#define SPZ_API
#endif

typedef struct spz_state spz_state; /* opaque: device SoA buffers + stream + scratch */

typedef enum {
    SPZ_OK = 0,
    SPZ_ERR_INVALID_ARG = 1, /* reference: assert!/debug_assert! failures, OOB */
    SPZ_ERR_UNSUPPORTED = 2, /* reference: todo!() / unimplemented!() */
    SPZ_ERR_CUDA = 3,
    SPZ_ERR_OOM = 4,
    SPZ_ERR_COMM = 5,
    SPZ_ERR_NO_DEVICE = 6
} spz_status;

/* Gate enum, gates.rs:44-74 (same order as the Rust enum). */
typedef enum {
    SPZ_GATE_H = 0,
    SPZ_GATE_M = 1,
    SPZ_GATE_X = 2,
    SPZ_GATE_Y = 3,
    SPZ_GATE_Z = 4,
    SPZ_GATE_P = 5,        /* p[0] = theta */
    SPZ_GATE_RX = 6,       /* p[0] = theta */
    SPZ_GATE_RY = 7,       /* p[0] = theta */
    SPZ_GATE_RZ = 8,       /* p[0] = theta */
    SPZ_GATE_SWAP = 9,     /* t0, t1 */
    SPZ_GATE_U = 10,       /* p = theta, phi, lambda */
    SPZ_GATE_UNITARY = 11, /* dense 2^m x 2^m fallback (unitaries.rs): out of scope -> SPZ_ERR_UNSUPPORTED */
    SPZ_GATE_BITFLIP = 12  /* p[0] = probability (gates.rs:1365-1374) */
} spz_gate_kind;

typedef struct {
    int32_t kind;   /* spz_gate_kind */
    int32_t t0, t1; /* SWAP operands */
    int32_t reserved;
    double p[3];
} spz_gate;

/* Controls enum, circuit.rs:55-70 */
typedef enum {
    SPZ_CTRL_NONE = 0, SPZ_CTRL_SINGLE = 1, SPZ_CTRL_ONES = 2,
    SPZ_CTRL_MIXED = 3, /* zeros are DROPPED from the mask, as mc_apply does (gates.rs:298-311) */
    SPZ_CTRL_SIGNED = 4 /* extension, not in the reference: zeros_mask (a subset of ctrl_mask) are true negative controls */
} spz_ctrl_kind;

/* QuantumTransformation, circuit.rs:113-120 (flattened; 64 bytes) */
typedef struct {
    int32_t kind;   /* spz_gate_kind */
    int32_t target;
    int32_t t0, t1; /* SWAP operands */
    double p[3];
    int32_t ctrl_kind; /* spz_ctrl_kind */
    int32_t reserved;
    uint64_t ctrl_mask;  /* set of control qubits */
    uint64_t zeros_mask; /* Mixed { zeros } */
} spz_op;

/* spz_execute flags */
#define SPZ_EXEC_NO_FUSE 0u  /* one kernel per gate, arithmetic bit-identical to spz_apply & co. */
#define SPZ_EXEC_FUSE 1u     /* batch runs of gates into on-chip tiles (one HBM pass per batch); runs of diagonal
                                gates are merged into phase accumulators (differs from gate-by-gate by a few ulp) */
#define SPZ_EXEC_KEEP_ORDER 4u /* with SPZ_EXEC_FUSE: do not reorder commuting gates across the op list when packing passes */
#define SPZ_EXEC_EXACT 2u    /* with SPZ_EXEC_FUSE: apply every gate with the reference arithmetic -> results are
                                bit-identical to SPZ_EXEC_NO_FUSE (slower on long diagonal runs) */

/* ---- library -------------------------------------------------------------------------------- */
SPZ_API int spz_abi_version(void);
SPZ_API const char *spz_last_error(void);          /* thread-local text of the last failure */
SPZ_API int spz_device_count(void);                /* 0 when no CUDA device is visible */
SPZ_API const char *spz_status_string(int status);

/* ---- State, core.rs:18-51 --------------------------------------------------------------------- */
SPZ_API int spz_create(int n_qubits, int device, spz_state **out); /* State::new (core.rs:32-42): |0..0> */
SPZ_API int spz_destroy(spz_state *st);                             /* Drop */
SPZ_API int spz_clone(const spz_state *st, spz_state **out);        /* #[derive(Clone)] core.rs:18 */
SPZ_API int spz_num_qubits(const spz_state *st);                    /* State.n */
SPZ_API int64_t spz_len(const spz_state *st);                       /* State::len core.rs:48 */
SPZ_API int spz_reset_zero(spz_state *st);                          /* back to |0..0> */
SPZ_API int spz_set_basis(spz_state *st, uint64_t index);           /* |index> */
SPZ_API int spz_init_random(spz_state *st, uint64_t seed);          /* utils.rs:168-201 recipe, counter-based RNG on device */
/* `state.reals[..] / state.imags[..]` access (tests index the Vecs directly, e.g. gates.rs:1541) */
SPZ_API int spz_upload(spz_state *st, const double *re, const double *im, int64_t offset, int64_t count);
SPZ_API int spz_download(const spz_state *st, double *re, double *im, int64_t offset, int64_t count);
/* Whole-state upload that returns at once (no counterpart in the reference, whose State lives in host memory: this is
   `State { reals, imags, n }` for a device-resident state, overlapped with what follows).  The state arrives in contiguous
   pieces on a copy stream and the gates / fused passes issued next run piece by piece behind the bus.  re / im: page-locked
   (spz_alloc_host), 2^n doubles each (the shard's length for a sharded register), untouched until the next spz_sync,
   spz_download or spz_upload. */
SPZ_API int spz_upload_async(spz_state *st, const double *re, const double *im);
SPZ_API int spz_sync(spz_state *st);
/* page-locked host buffers for spz_upload / spz_download at full PCIe speed (plain malloc memory also works) */
SPZ_API int spz_alloc_host(uint64_t bytes, void **out);
SPZ_API int spz_free_host(void *ptr);

/* ---- gates, gates.rs -------------------------------------------------------------------------- */
SPZ_API int spz_apply(spz_state *st, const spz_gate *gate, int target);                 /* apply   gates.rs:215 */
SPZ_API int spz_c_apply(spz_state *st, const spz_gate *gate, int control, int target);  /* c_apply gates.rs:257 */
SPZ_API int spz_cc_apply(spz_state *st, const spz_gate *gate, int control0, int control1, int target); /* gates.rs:272 */
/* mc_apply gates.rs:290; zeros may be NULL (None).  Controls listed in zeros are dropped from the mask,
   as the reference does (gates.rs:298-311). */
SPZ_API int spz_mc_apply(spz_state *st, const spz_gate *gate, const int32_t *controls, int n_controls,
                 const int32_t *zeros, int n_zeros, int target);
/* extension: any 1-qubit gate under any all-ones control mask */
SPZ_API int spz_mc_apply_mask(spz_state *st, const spz_gate *gate, uint64_t ctrl_mask, int target);
/* extension: signed controls -- every qubit of ones_mask must be 1 and every qubit of zeros_mask must be 0 (disjoint masks).
   What Controls::Mixed { zeros } (circuit.rs:61-69) describes and mc_apply does not do (gates.rs:298-311 drops the zeros). */
SPZ_API int spz_mc_apply_signed(spz_state *st, const spz_gate *gate, uint64_t ones_mask, uint64_t zeros_mask, int target);
SPZ_API int spz_iqft(spz_state *st, const int32_t *targets, int n_targets);             /* iqft core.rs:184 */

/* ---- QuantumCircuit::execute, circuit.rs:552-600 ---------------------------------------------- */
/* measured_mask / measured_vals are the two u64 of QubitTracker (circuit.rs:122-164), owned by the
   caller's circuit object and updated in place (may be NULL when the op list has no SPZ_GATE_M and no
   classically controlled gate).  Randomness for M / BitFlipNoise comes from the state's seeded
   generator (spz_set_seed); the reference uses unseeded thread_rng. */
SPZ_API int spz_execute(spz_state *st, const spz_op *ops, int64_t n_ops, uint32_t flags, uint64_t *measured_mask,
                uint64_t *measured_vals);
SPZ_API int spz_set_seed(spz_state *st, uint64_t seed);
/* Planning only (pure host code, no CUDA): how spz_execute would schedule the list.  out_order[k] = index of the op
   executed k-th, out_pass[k] = the HBM pass (kernel launch) it belongs to; entries past the number of scheduled ops
   are -1 (e.g. a BitFlipNoise that did not fire).  Used by the CPU tests of the scheduler. */
SPZ_API int spz_plan_fusion(int n_qubits, const spz_op *ops, int64_t n_ops, uint32_t flags, int32_t *out_order,
                            int32_t *out_pass, int32_t *out_n_passes);
/* Test hook (pure host code): the tile micro-program spz_execute would launch for pass `pass_index`, serialised into
   `out` (layout documented at the definition in csrc/abi.cu).  tests/test_tile_program.py interprets it in NumPy. */
/* Test hook (pure host code): every step rank `rank` of `world` would execute for the list -- fused passes, single ops,
   exchanges -- serialised (layout at the definition in csrc/abi.cu).  tests/test_dist_fused_cpu.py replays it in NumPy. */
SPZ_API int spz_debug_compile_sharded(int n_total, int world, int rank, const spz_op *ops, int64_t n_ops, uint32_t flags, void *out,
                                      int64_t out_bytes, int64_t *out_used);
SPZ_API int spz_debug_compile_pass(int n_qubits, const spz_op *ops, int64_t n_ops, uint32_t flags, int pass_index, void *out,
                                   int64_t out_bytes, int64_t *out_used);

/* ---- reductions: measurement.rs:12-92, core.rs:65-129,198-264 --------------------------------- */
SPZ_API int spz_prob0(spz_state *st, int target, double *out);   /* measurement.rs:16-29 */
SPZ_API int spz_norm2(spz_state *st, double *out);               /* sum |amp|^2 */
/* measure_qubit measurement.rs:12: forced_v = 0/1 (Some(v)) or -1 (None: draw Bernoulli(1-prob0) from
   the state's generator).  out_bit receives the outcome. */
SPZ_API int spz_measure_qubit(spz_state *st, int target, int reset, int forced_v, int *out_bit);
SPZ_API int spz_qubit_expectation_value(spz_state *st, int target, double *out); /* core.rs:198 */
/* xyz_expectation_value core.rs:222: observable in {'x','y','z'}, one value per target */
SPZ_API int spz_xyz_expectation_value(spz_state *st, char observable, const int32_t *targets, int n_targets, double *out);
/* Sampling (replaces reservoir_sampling core.rs:125): `shots` exact inverse-CDF draws; u01[k] in [0,1)
   supplied by the caller (so results are reproducible and checkable); out_index[k] = smallest i with
   cdf(i) > u01[k] * norm2.  On a sharded register every rank passes the same u01; a shot is answered by the rank
   that owns it (logical basis index) and comes back as -1 on the others (combine with a max). */
SPZ_API int spz_sample(spz_state *st, const double *u01, int64_t shots, int64_t *out_index);

/* ---- multi-GPU: one process per GPU, amplitudes sharded by the top log2(world) index bits ------ */
/* The reference has no distributed layer; this is the engine's own.  Every rank calls the
   same sequence of spz_* functions on its handle (SPMD).  Local-qubit gates and all diagonal gates need no
   communication; a non-diagonal gate on a global qubit first trades that qubit for a local one by a pairwise
   half-shard exchange written directly into the partner's HBM through CUDA IPC peer pointers over NVLink.
   The qubit permutation (spz_dist_perm) is the engine's to choose: exchanges update it, and spz_execute on a register that is
   still a computational basis state (spz_dist_create, spz_reset_zero, spz_set_basis) first picks it from the op list so that
   the qubits used last are the global ones (checked on the device before it is relied on; SPZ_DIST_PLACE=0 disables). */
typedef struct {
    int32_t type;    /* 0 skip, 1 local gate, 2 diagonal constant factor, 3 exchange */
    int32_t kind;    /* spz_gate_kind */
    int32_t target;  /* type 1: local physical bit */
    int32_t hi;      /* type 2: value of the (global) target bit on this rank */
    uint64_t cmask;  /* types 1,2: local physical control mask */
    int32_t gbit;    /* type 3: which bit of the rank is exchanged */
    int32_t lq;      /* type 3: local physical bit it trades places with */
    int32_t partner; /* type 3: rank ^ (1 << gbit) */
    int32_t grefs;   /* types 0,1,2: mask of the rank bits the lowering consulted (global controls / global diagonal target) */
    double p[3];
} spz_dist_action;

#define SPZ_IPC_BLOB_BYTES 256
SPZ_API int spz_dist_create(int n_qubits, int rank, int world, int device, spz_state **out);
SPZ_API int spz_dist_export(spz_state *st, void *blob);        /* SPZ_IPC_BLOB_BYTES of IPC handles */
SPZ_API int spz_dist_connect(spz_state *st, const void *blobs); /* world blobs, rank order (all-gathered by the host) */
/* same-process variant (one host thread per shard; shards on one or several GPUs): direct peer pointers, no IPC */
SPZ_API int spz_dist_connect_local(spz_state **states, int world);
SPZ_API int spz_dist_perm(const spz_state *st, int32_t *perm_out); /* logical qubit -> physical bit, n_qubits entries */
SPZ_API int spz_dist_local_qubits(const spz_state *st);
/* Clone of a sharded register (core.rs:18), a collective: every rank creates + connects a second register, then copies its
   shard, the qubit permutation and the measurement RNG state into it with this call. */
SPZ_API int spz_dist_copy_from(spz_state *dst, const spz_state *src);
SPZ_API int spz_dist_stats(const spz_state *st, double *out4);  /* exchanges, bytes sent, exchange ms, overlapped exchanges */
/* Host control plane between the processes of one node, without any framework (csrc/rendezvous.cu; no reference counterpart:
   the reference is single-process): an all-gather of small blobs and a barrier through files in a shared directory.
   dir = NULL: /dev/shm/spz_rdv_<MASTER_PORT>_<launcher pid>_<uid>, the same for every rank of one torchrun-style launch. */
typedef struct spz_rdv spz_rdv;
SPZ_API int spz_rdv_open(const char *dir, int rank, int world, spz_rdv **out);
SPZ_API int spz_rdv_allgather(spz_rdv *r, const void *mine, int64_t bytes, void *all); /* all: world * bytes, rank order */
SPZ_API int spz_rdv_barrier(spz_rdv *r);
SPZ_API int spz_rdv_close(spz_rdv *r);
SPZ_API int spz_dist_connect_rdv(spz_state *st, spz_rdv *r); /* spz_dist_export + all-gather + spz_dist_connect + barrier */
/* planning only (pure host code, no CUDA): what the CPU tests of the sharding logic drive */
typedef struct spz_dist_plan spz_dist_plan;
SPZ_API int spz_dist_plan_create(int n_qubits, int world, spz_dist_plan **out);
SPZ_API int spz_dist_plan_destroy(spz_dist_plan *p);
SPZ_API int spz_dist_plan_lower(spz_dist_plan *p, int rank, const spz_op *op, spz_dist_action *out, int max_out, int *n_out);
SPZ_API int spz_dist_plan_perm(const spz_dist_plan *p, int32_t *perm_out);

/* ---- instrumentation --------------------------------------------------------------------------- */
SPZ_API int spz_timer_start(spz_state *st);               /* cudaEventRecord on the state's stream */
SPZ_API int spz_timer_stop(spz_state *st, double *out_ms); /* record + synchronise + elapsed */
SPZ_API int64_t spz_launch_count(void);                   /* kernels launched by this library since load */
SPZ_API int spz_device_name(int device, char *buf, int buflen);
SPZ_API int spz_mem_info(int device, uint64_t *free_bytes, uint64_t *total_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SPINOZA_B200_H */
