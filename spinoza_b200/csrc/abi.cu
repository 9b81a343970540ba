// abi.cu -- the C ABI of libspinoza_b200 (include/spinoza_b200.h): state management, the gate entry points
// mirroring gates.rs:215-320, and QuantumCircuit::execute (circuit.rs:552-600) with the fusion scheduler.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "dist_plan.h"
#include "engine.h"

namespace spz {

// Device-side handshakes (dist.cu) spin until a peer's kernel arrives.  With CUDA's default lazy module loading
// the first launch of a kernel may need a context-wide synchronisation, which can never complete while such a
// spinning kernel is resident -> load every kernel when the context is created instead.  (Only matters when two
// shards share a GPU or a process; harmless otherwise.)  Runs at dlopen, before the first CUDA call.
// Likewise, streams of different shards must not share a hardware work queue (a queue whose head is a spinning
// handshake would block the very kernel it waits for): ask for the maximum number of connections (default 8).
__attribute__((constructor)) static void spz_force_eager_module_loading() {
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
}

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in `%s` at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    if (e == cudaErrorMemoryAllocation) return SPZ_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return SPZ_ERR_NO_DEVICE;
    return SPZ_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int arrival_join(spz_state *st) {
    spz_state::Arrival &a = st->arrival;
    if (a.lanes_active) { // the per-piece streams hand the state back to the main stream
        for (int k = 0; k < a.chunks; ++k) {
            SPZ_CUDA(cudaEventRecord(a.ev[k], a.lane[k]));
            SPZ_CUDA(cudaStreamWaitEvent(st->stream, a.ev[k], 0));
        }
        a.lanes_active = false;
    }
    if (a.pending) {
        SPZ_CUDA(cudaStreamWaitEvent(st->stream, a.ev[a.chunks - 1], 0));
        a.pending = false;
    }
    return SPZ_OK;
}

bool take_chunks(spz_state *st, int *n_chunks, cudaEvent_t *ev) {
    spz_state::Arrival &a = st->arrival;
    if (a.lanes_active) { arrival_join(st); return false; }
    if (a.pending) { // (an upload waits for everything before it, so no exchange can be in flight at the same time)
        *n_chunks = a.chunks;
        for (int k = 0; k < a.chunks; ++k) ev[k] = a.ev[k];
        a.pending = false;
        return true;
    }
    return st->dist && dist_take_chunks(st, n_chunks, ev);
}

#include "gate_resolve.inl" // int resolve_gate(kind, params, out): shared with the CPU emulation harness (tests/emu/)

static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double next_u01(spz_state *st) { return (double)(splitmix64(&st->rng) >> 11) * (1.0 / 9007199254740992.0); }

static int bind(const spz_state *st) {
    SPZ_CUDA(cudaSetDevice(st->device));
    return SPZ_OK;
}

#define SPZ_CHECK_STATE(st)                                              \
    do {                                                                 \
        if (!(st)) { ::spz::set_error("null state handle"); return SPZ_ERR_INVALID_ARG; } \
        SPZ_TRY(::spz::bind(st));                                        \
    } while (0)

static inline int total_qubits(const spz_state *st) { return st->dist ? dist_total_qubits(st) : st->n; }

// ---- one op, unfused ---------------------------------------------------------------------------------
static int apply_masked(spz_state *st, int kind, const double *p, uint64_t ctrl_mask, int target) {
    if (st->dist) {
        GateK probe;
        SPZ_TRY(resolve_gate(kind, p, &probe)); // same UNSUPPORTED behaviour as the single-GPU path
        return dist_apply_masked(st, kind, p, 0, 0, ctrl_mask, target);
    }
    GateK g;
    SPZ_TRY(resolve_gate(kind, p, &g));
    return launch_gate(st, g, ctrl_mask, target);
}

// Signed controls (extension; SURVEY 2.3 B4 -- the reference's Controls::Mixed only pretends to have them): ones must be 1,
// zeros must be 0.  One launch of the ordinary pair kernel on a single-GPU register; a sharded register conjugates with X on
// the zero-controls (correct wherever those qubits live; an X on a global qubit is an exchange).
static int apply_signed(spz_state *st, int kind, const double *p, uint64_t ones, uint64_t zeros, int target) {
    if (ones & zeros) { set_error("a qubit cannot be both a positive and a negative control (0x%llx)", (unsigned long long)(ones & zeros)); return SPZ_ERR_INVALID_ARG; }
    const int nq = total_qubits(st);
    if ((nq < 64 && ((ones | zeros) >> nq)) || target < 0 || target >= nq || (((ones | zeros) >> target) & 1ull)) {
        set_error("bad signed controls (ones 0x%llx, zeros 0x%llx, target %d, %d qubits)", (unsigned long long)ones, (unsigned long long)zeros, target, nq);
        return SPZ_ERR_INVALID_ARG;
    }
    if (!zeros) return apply_masked(st, kind, p, ones, target);
    if (!st->dist) {
        GateK g;
        SPZ_TRY(resolve_gate(kind, p, &g));
        return launch_gate_signed(st, g, ones | zeros, zeros, target);
    }
    for (int q = 0; q < nq; ++q) if ((zeros >> q) & 1ull) SPZ_TRY(apply_masked(st, SPZ_GATE_X, nullptr, 0, q));
    SPZ_TRY(apply_masked(st, kind, p, ones | zeros, target));
    for (int q = 0; q < nq; ++q) if ((zeros >> q) & 1ull) SPZ_TRY(apply_masked(st, SPZ_GATE_X, nullptr, 0, q));
    return SPZ_OK;
}

static int swap_impl(spz_state *st, int t0, int t1) {
    if (st->dist) return dist_apply_masked(st, SPZ_GATE_SWAP, nullptr, t0, t1, 0, 0); // relabel, no data moves
    return launch_swap(st, t0, t1);
}

// zero_mask: qubits known to be exactly |0> (measured with reset earlier in the same run of measurements, see execute_impl);
// single-GPU registers then visit only the indices where those bits are 0.
static int measure_impl(spz_state *st, int target, int reset, int forced_v, int *out_bit, uint64_t zero_mask = 0) {
    if (target < 0 || target >= total_qubits(st)) { set_error("target %d out of range", target); return SPZ_ERR_INVALID_ARG; }
    if (forced_v > 1) { set_error("forced outcome must be 0, 1 or -1"); return SPZ_ERR_INVALID_ARG; } // assert measurement.rs:32
    zero_mask &= ~(1ull << target);
    if (st->dist || st->n > kZMaxBits) zero_mask = 0;
    double prob0 = 0.0;
    if (st->dist) SPZ_TRY(dist_reduce_scalar(st, 0, target, &prob0));
    else if (zero_mask) SPZ_TRY(reduce_prob0_sub(st, target, zero_mask, &prob0));
    else SPZ_TRY(reduce_scalar(st, 0, target, &prob0));
    int val;
    if (forced_v >= 0) val = forced_v;
    else val = next_u01(st) < 1.0 - prob0 ? 1 : 0; // Binomial(1, 1 - prob0) measurement.rs:35-36
    double k;
    if (val == 0) k = 1.0 / std::sqrt(prob0);       // prob0.sqrt().recip() measurement.rs:40
    else k = 1.0 / std::sqrt(1.0 - prob0);          // measurement.rs:63-64
    if (st->dist) {
        SPZ_TRY(dist_collapse(st, target, val, k));
        if (val == 1 && reset) SPZ_TRY(apply_masked(st, SPZ_GATE_X, nullptr, 0, target)); // measurement.rs:87-89
    } else if (zero_mask) {
        SPZ_TRY(launch_collapse_sub(st, target, val, reset, k, zero_mask));
    } else {
        SPZ_TRY(launch_collapse(st, target, val, reset, k));
    }
    if (out_bit) *out_bit = val;
    return SPZ_OK;
}

// ---- fusion scheduler -----------------------------------------------------------------------------------
struct ROp { // an op resolved to what the kernels need (physical qubits of this handle's shard)
    int kind;
    int target, t2;
    uint64_t cmask;
    GateK g;
    int const_hi = -1; // >= 0: diagonal gate whose target is a rank bit of a sharded register; value of that bit
    double theta = 0.0; // first gate parameter (merged diagonal factors are computed from the angle itself)
    int src = -1;       // index of the originating op in the caller's list (planning output)
    // sharded registers, exchange-spanning windows (SPZ_DIST_WINDOW): see Fuser::schedule_dag
    bool skip = false;  // placeholder of an op this rank skips (a global control is 0 here): scheduled like the real op so
                        // that every rank forms the same passes, compiled to nothing
    uint32_t grefs = 0; // rank bits the lowering of this op consulted
};
constexpr int kRopExchange = 100; // ROp::kind of an exchange inside a window: target = local physical bit, t2 = rank bit

// Dry-run sink: instead of launching kernels, record which pass every op ends up in (spz_plan_fusion).
struct PlanSink {
    std::vector<int32_t> order, group;
    int n_groups = 0;
    // optional: the compiled micro-program of pass `capture_pass` (spz_debug_compile_pass)
    int capture_pass = -1;
    bool captured = false, captured_direct = false;
    TilePlan cap_plan{};
    std::vector<TileInstr> cap_prog;
    std::vector<TileGroup> cap_groups;
    std::vector<TileTerm> cap_terms;
    // optional: everything a (sharded) execute would do, in order (spz_debug_compile_sharded)
    struct Step {
        int type = 0; // 0 = tile program, 1 = single op, 2 = exchange, 3 = measurement
        TilePlan plan{};
        std::vector<TileInstr> prog;
        std::vector<TileGroup> groups;
        std::vector<TileTerm> terms;
        ROp op{};
        int gbit = 0, lq = 0;
    };
    bool capture_all = false;
    std::vector<Step> steps;
    void take(const std::vector<ROp> &ops) {
        for (const ROp &o : ops) { order.push_back(o.src); group.push_back(n_groups); }
        ++n_groups;
        if (capture_all && ops.size() == 1) {
            Step st;
            st.type = ops[0].kind == SPZ_GATE_M ? 3 : 1;
            st.op = ops[0];
            steps.push_back(std::move(st));
        }
    }
    void exchange(int gbit, int lq) {
        if (!capture_all) return;
        Step st;
        st.type = 2; st.gbit = gbit; st.lq = lq;
        steps.push_back(std::move(st));
    }
    int32_t relabel_perm[64];
    void relabel(const int32_t *perm) { // the basis state was re-placed before the first op (dist_place_basis): type 4
        if (!capture_all) return;
        for (int q = 0; q < 64; ++q) relabel_perm[q] = perm[q];
        Step st;
        st.type = 4;
        steps.push_back(std::move(st));
    }
};

struct Fuser {
    spz_state *st;
    int T;      // tile bits (<= n)
    int Lmin;   // smallest contiguous run we accept (coalescing)
    std::vector<ROp> ops;
    uint64_t high_set = 0;
    int low_need = 0;
    bool exact = false;
    PlanSink *sink = nullptr;

    Fuser(spz_state *s) : st(s) {
        T = std::min(max_tile_bits(), s->n);
        // 2^Lmin contiguous amplitudes per tile segment: 6 (512 bytes per array) keeps every access a run of full 128-byte
        // lines.  SPZ_TILE_LMIN=4|5 trades segment length for up to 8 (7) arbitrary high qubits per pass instead of 6 --
        // an experiment knob: the kernels are correct for any L >= 4, only coalescing changes.
        int lmin = 6;
        if (const char *e = std::getenv("SPZ_TILE_LMIN")) if (e[0] >= '4' && e[0] <= '6' && !e[1]) lmin = e[0] - '0';
        Lmin = std::min(lmin, T);
    }
    int n_high() const { return __builtin_popcountll(high_set); }

    // schedule_dag() chooses the pass's tile qubits before filling it: high_set is then final and must not grow
    bool tile_fixed = false;

    bool try_add_target(int q, uint64_t &hs, int &ln) const {
        const int Lcur = T - __builtin_popcountll(hs);
        if (q < Lcur) { ln = std::max(ln, q + 1); return true; }
        if ((hs >> q) & 1ull) return true;
        if (tile_fixed) return false;
        const int Lnew = Lcur - 1;
        if (Lnew < std::max(ln, Lmin) || __builtin_popcountll(hs) + 1 > 8) return false;
        hs |= 1ull << q;
        return true;
    }

    bool fits(const ROp &op, uint64_t &hs, int &ln) const {
        hs = high_set; ln = low_need;
        // low_need records every target accepted as a low tile bit, so L can never shrink below it
        if (op.kind == SPZ_GATE_SWAP) {
            // Either operand may have to become a high tile bit, and taking the larger one first can make room that the
            // other order does not (SWAP(11, 13) in an empty 12-bit tile: 11 is low only while there is no high bit).
            if (try_add_target(op.target, hs, ln) && try_add_target(op.t2, hs, ln)) return true;
            hs = high_set; ln = low_need;
            return try_add_target(op.t2, hs, ln) && try_add_target(op.target, hs, ln);
        }
        if (is_diagonal_kind(op.kind)) return true; // includes const_hi ops
        return try_add_target(op.target, hs, ln);
    }

    // Compile the group into the tile kernel's micro-program (see kernels_tile.cu).
    int compile(const TilePlan &plan, std::vector<TileInstr> &prog, std::vector<TileGroup> &groups,
                std::vector<TileTerm> &terms) const {
        auto tile_bit = [&](int q) -> int {
            if (q < plan.low_bits) return q;
            for (int i = 0; i < plan.n_high; ++i) if (plan.high[i] == q) return plan.low_bits + i;
            return -1;
        };
        // SWAP(a, b) = CX(a,b) CX(b,a) CX(a,b) (utils.rs:204-208): exact, and each CX is a register butterfly
        std::vector<ROp> lops;
        lops.reserve(ops.size() + 8);
        for (const ROp &o : ops) {
            if (o.kind != SPZ_GATE_SWAP) { lops.push_back(o); continue; }
            ROp cx{};
            cx.kind = SPZ_GATE_X; cx.g.kind = SPZ_GATE_X;
            cx.target = o.t2; cx.cmask = 1ull << o.target; lops.push_back(cx);
            cx.target = o.target; cx.cmask = 1ull << o.t2; lops.push_back(cx);
            cx.target = o.t2; cx.cmask = 1ull << o.target; lops.push_back(cx);
        }
        int R[4] = {0, 1, 2, 3};
        bool haveR = false;
        // merged mode: the diagonal gates of the current run, as phase terms bucketed by class
        // (m = 0, 1, 2, 4, 8, other); diagonal gates commute, so a run may be reordered freely
        std::vector<TileTerm> bucket[6];
        auto close_run = [&]() {
            size_t total = 0;
            for (auto &b : bucket) total += b.size();
            if (!total) return;
            TileInstr r{};
            r.op = TI_RUN;
            r.rpos = (int)groups.size();
            int counts[6];
            for (int c = 0; c < 6; ++c) {
                auto &b = bucket[c];
                // terms with the same (thr, m) become one group (their product is a per-tile constant)
                std::stable_sort(b.begin(), b.end(), [](const TileTerm &x, const TileTerm &y) {
                    return x.thr != y.thr ? x.thr < y.thr : x.m < y.m;
                });
                counts[c] = 0;
                for (size_t i = 0; i < b.size();) {
                    size_t j = i;
                    while (j < b.size() && b[j].thr == b[i].thr && b[j].m == b[i].m) ++j;
                    TileGroup g{b[i].thr, b[i].m, (int)terms.size(), (int)(j - i)};
                    groups.push_back(g);
                    terms.insert(terms.end(), b.begin() + i, b.begin() + j);
                    ++counts[c];
                    i = j;
                }
                b.clear();
            }
            r.rbit[0] = counts[0]; r.rbit[1] = counts[1]; r.rbit[2] = counts[2]; r.rbit[3] = counts[3];
            r.reg_cmask = (uint32_t)counts[4]; r.thr_cmask = (uint32_t)counts[5];
            prog.push_back(r);
        };
        auto add_term = [&](uint64_t outer, uint32_t thr, uint32_t m, double fr, double fi) {
            TileTerm t{outer, thr, m, fr, fi};
            const int cls = m == 0 ? 0 : m == 1 ? 1 : m == 2 ? 2 : m == 4 ? 3 : m == 8 ? 4 : 5;
            bucket[cls].push_back(t);
        };
        auto idx_in_R = [&](int b) -> int { for (int i = 0; i < 4; ++i) if (R[i] == b) return i; return -1; };
        auto choose_layout = [&](size_t from, int must) {
            int pick[4], np = 0;
            auto add = [&](int b) { for (int i = 0; i < np; ++i) if (pick[i] == b) return; if (np < 4) pick[np++] = b; };
            if (must >= 0) add(must);
            for (size_t j = from; j < lops.size() && np < 4; ++j) {
                const ROp &o = lops[j];
                if (is_diagonal_kind(o.kind) || o.const_hi >= 0) continue;
                const int b = tile_bit(o.target);
                if (b >= 0) add(b);
            }
            for (int b = T - 1; b >= 0 && np < 4; --b) add(b);
            std::sort(pick, pick + 4);
            close_run(); // term masks are relative to the layout
            for (int i = 0; i < 4; ++i) R[i] = pick[i];
            haveR = true;
            TileInstr li{};
            li.op = TI_LAYOUT;
            for (int i = 0; i < 4; ++i) li.rbit[i] = R[i];
            prog.push_back(li);
        };
        for (size_t i = 0; i < lops.size(); ++i) {
            const ROp &o = lops[i];
            const bool diag = is_diagonal_kind(o.kind);
            const int tb = o.const_hi >= 0 ? -1 : tile_bit(o.target);
            if (!diag && tb < 0) { set_error("internal: non-diagonal op left outside its tile"); return SPZ_ERR_INVALID_ARG; }
            if (!diag) { if (!haveR || idx_in_R(tb) < 0) choose_layout(i, tb); }
            else if (!haveR) choose_layout(i, -1);
            TileInstr t{};
            t.op = diag ? TI_DIAG : TI_GATE;
            t.kind = o.kind;
            t.outer_target = -1;
            for (int q = 0; q < st->n; ++q) {
                if (!((o.cmask >> q) & 1ull)) continue;
                const int b = tile_bit(q);
                if (b < 0) t.outer_cmask |= 1ull << q;
                else if (idx_in_R(b) >= 0) t.reg_cmask |= 1u << idx_in_R(b);
                else t.thr_cmask |= 1u << b;
            }
            for (int j = 0; j < 7; ++j) t.s[j] = o.g.s[j];
            if (!diag) {
                close_run(); // the gates of a run commute with each other, not with this butterfly
                t.rpos = idx_in_R(tb);
                for (unsigned k = 0; k < 16; ++k) // register indices whose register-bit controls are all set
                    if ((k & t.reg_cmask) == t.reg_cmask) t.t_mask |= 1u << k;
            } else {
                if (o.const_hi >= 0) { t.t_where = 0; t.const_hi = (uint32_t)(o.const_hi + 1); }
                else if (tb < 0) { t.t_where = 0; t.outer_target = o.target; }
                else if (idx_in_R(tb) >= 0) { t.t_where = 2; t.t_mask = 1u << idx_in_R(tb); }
                else { t.t_where = 1; t.t_mask = 1u << tb; }
                t.f0[0] = 1.0; t.f0[1] = 0.0;
                if (o.kind == SPZ_GATE_Z) { t.f1[0] = -1.0; t.f1[1] = 0.0; }
                else { t.f1[0] = std::cos(o.theta); t.f1[1] = std::sin(o.theta); } // P and RZ: e^{i theta} on target-bit-1
                if (o.kind == SPZ_GATE_RZ) { t.has_f0 = 1; t.f0[0] = o.g.s[0]; t.f0[1] = -o.g.s[1]; } // d0 = (cos, -sin)(theta/2)
                if (!exact) { // fold into the run instead of emitting an instruction
                    if (t.has_f0) add_term(t.outer_cmask, t.thr_cmask, t.reg_cmask, t.f0[0], t.f0[1]);
                    if (t.t_where == 0) {
                        if (t.const_hi == 2) add_term(t.outer_cmask, t.thr_cmask, t.reg_cmask, t.f1[0], t.f1[1]);
                        else if (t.const_hi == 0) add_term(t.outer_cmask | (1ull << t.outer_target), t.thr_cmask, t.reg_cmask, t.f1[0], t.f1[1]);
                    } else if (t.t_where == 1) add_term(t.outer_cmask, t.thr_cmask | t.t_mask, t.reg_cmask, t.f1[0], t.f1[1]);
                    else add_term(t.outer_cmask, t.thr_cmask, t.reg_cmask | t.t_mask, t.f1[0], t.f1[1]);
                    continue;
                }
            }
            prog.push_back(t);
        }
        close_run();
        return SPZ_OK;
    }

    // Launch the current group (ops / high_set / low_need) and reset it.
    int emit_group() {
        // placeholders (ops another rank applies and this one skips) took part in the scheduling only
        ops.erase(std::remove_if(ops.begin(), ops.end(), [](const ROp &o) { return o.skip; }), ops.end());
        if (ops.empty()) { high_set = 0; low_need = 0; return SPZ_OK; }
        if (sink) {
            if (sink->capture_pass == sink->n_groups) { // compile exactly as the launch path below would
                sink->captured = true;
                if (ops.size() == 1) {
                    sink->captured_direct = true; // a single op goes to the direct kernel: no micro-program
                } else {
                    TilePlan plan{};
                    plan.n_high = n_high();
                    plan.tile_bits = T;
                    plan.low_bits = T - plan.n_high;
                    int k = 0;
                    for (int q = 0; q < 64; ++q) if ((high_set >> q) & 1ull) plan.high[k++] = q;
                    sink->cap_plan = plan;
                    SPZ_TRY(compile(plan, sink->cap_prog, sink->cap_groups, sink->cap_terms));
                }
            }
            if (sink->capture_all && ops.size() > 1) {
                PlanSink::Step step;
                step.type = 0;
                step.plan.n_high = n_high();
                step.plan.tile_bits = T;
                step.plan.low_bits = T - step.plan.n_high;
                int k = 0;
                for (int q = 0; q < 64; ++q) if ((high_set >> q) & 1ull) step.plan.high[k++] = q;
                SPZ_TRY(compile(step.plan, step.prog, step.groups, step.terms));
                sink->steps.push_back(std::move(step));
            }
            sink->take(ops); ops.clear(); high_set = 0; low_need = 0; return SPZ_OK;
        }
        int rc = SPZ_OK;
        // A group of one op goes to the one-gate-per-pass kernels, which run at the HBM roofline.  SPZ_TILE_MIN_OPS=k
        // (default 2) sends groups of fewer than k ops the same way, one launch each in group order: a knob for tuning
        // against the tile kernel's fixed cost per pass (results are identical either way).
        size_t min_tile_ops = 2;
        if (const char *e = std::getenv("SPZ_TILE_MIN_OPS")) { const int v = std::atoi(e); if (v >= 2 && v <= 16) min_tile_ops = (size_t)v; }
        if (ops.size() < min_tile_ops) {
            for (size_t i = 0; i < ops.size() && rc == SPZ_OK; ++i) {
                const ROp &o = ops[i];
                if (o.const_hi >= 0) rc = dist_diag_const(st, o.g, o.cmask, o.const_hi);
                else if (o.kind != SPZ_GATE_SWAP) rc = launch_gate(st, o.g, o.cmask, o.target);
                else rc = launch_swap(st, o.target, o.t2);
            }
        } else {
            TilePlan plan{};
            plan.n_high = n_high();
            plan.tile_bits = T;
            plan.low_bits = T - plan.n_high;
            int k = 0;
            for (int q = 0; q < 64; ++q) if ((high_set >> q) & 1ull) plan.high[k++] = q;
            std::vector<TileInstr> prog;
            std::vector<TileGroup> groups;
            std::vector<TileTerm> terms;
            prog.reserve(ops.size() + 16);
            rc = compile(plan, prog, groups, terms);
            if (rc == SPZ_OK)
                rc = launch_tile_program(st, plan, prog.data(), (int)prog.size(), groups.data(), (int)groups.size(), terms.data(),
                                         (int)terms.size(), exact);
        }
        ops.clear();
        high_set = 0;
        low_need = 0;
        return rc;
    }

    // Append to the current group, closing it first when the op does not fit (order-preserving greedy).
    int add(const ROp &op) {
        uint64_t hs; int ln;
        if (!fits(op, hs, ln) || ops.size() >= (size_t)kMaxTileGroups / 2) {
            SPZ_TRY(emit_group());
            if (!fits(op, hs, ln)) { set_error("internal: op does not fit an empty tile"); return SPZ_ERR_INVALID_ARG; }
        }
        high_set = hs; low_need = ln;
        ops.push_back(op);
        return SPZ_OK;
    }

    // ---- scheduling window ---------------------------------------------------------------------------------
    // Ops are buffered and grouped at flush points (end of the list, a measurement, an exchange -- unless the window is
    // allowed to span exchanges, SPZ_DIST_WINDOW).  In the default
    // fused mode the window is scheduled as a dependency DAG: two ops need their program order only if they share
    // a qubit on which at least one of them is not "Z-like" (diagonal gates and controls are Z-like: block diagonal
    // in that qubit's computational basis).  Ready ops are packed into the current tile in program order, so e.g.
    // the rotations of the next layer on qubits the tile already holds join the pass instead of opening a new one.
    // Reordering commuting gates is exact mathematically but changes rounding in the last place, so EXACT mode
    // (bit-identical to unfused) keeps strict program order.
    std::vector<ROp> pending;
    bool reorder = true;

    int push(const ROp &op) {
        pending.push_back(op);
        if (pending.size() >= 4096) return flush();
        return SPZ_OK;
    }

    int flush() {
        int rc = (exact || !reorder || pending.size() < 3) ? schedule_in_order() : schedule_dag();
        pending.clear();
        return rc;
    }

    // an exchange queued in the window (sharded registers): everything scheduled so far must be on the device first.
    // `gate`: the uncontrolled non-diagonal op that asked for the exchange, to be applied by the exchange kernel itself
    // (SPZ_DIST_FUSE_GATE), or null.
    int run_exchange(const ROp &x, const ROp *gate) {
        SPZ_TRY(emit_group());
        if (sink) {
            sink->exchange(x.t2, x.target);
            if (gate) sink->take(std::vector<ROp>{*gate}); // the dry run describes the pair as two steps
            return SPZ_OK;
        }
        return gate ? dist_exchange_gate(st, x.t2, x.target, gate->g) : dist_exchange(st, x.t2, x.target);
    }
    // Is pending[gi] the op that asked for exchange pending[e], and may the exchange kernel apply it?  It is the next op
    // of the window (same source op), uncontrolled in every sense, and the exchange is all it still waits for: everything
    // earlier on its qubit consulted the rank bit the qubit sat on, so it already precedes the exchange.
    bool fusable_trigger(int e, int gi, int undone_preds) const {
        if (gi >= (int)pending.size()) return false;
        const ROp &x = pending[e], &g = pending[gi];
        return g.src == x.src && g.kind != kRopExchange && !g.skip && g.cmask == 0 && g.grefs == 0 && g.const_hi < 0 && undone_preds == 1 &&
               !is_diagonal_kind(g.kind) && dist_can_fuse_gate(st, g.kind, 0, g.target, x.target);
    }

    int schedule_in_order() {
        for (size_t i = 0; i < pending.size(); ++i) {
            const ROp &op = pending[i];
            if (op.kind == kRopExchange) {
                const bool fuse_g = fusable_trigger((int)i, (int)i + 1, 1);
                SPZ_TRY(run_exchange(op, fuse_g ? &pending[i + 1] : nullptr));
                if (fuse_g) ++i;
            } else {
                SPZ_TRY(add(op));
            }
        }
        return emit_group();
    }

    int schedule_dag() {
        const int N = (int)pending.size();
        std::vector<int> npred(N, 0), stamp(N, -1);
        std::vector<std::vector<int>> succ(N);
        int last_w[64];
        int last_exchange = -1;
        std::vector<int> zl[64]; // Z-like ops on the qubit since its last writer
        for (int &w : last_w) w = -1;
        for (int i = 0; i < N; ++i) {
            const ROp &op = pending[i];
            const bool diag = is_diagonal_kind(op.kind);
            // Index bits 56.. stand for the bits of the rank: an op whose lowering consulted rank bit b (a global control, a
            // diagonal target on a global qubit) is Z-like on 56 + b, and an exchange, which moves a local bit into that
            // rank bit, writes both.  So an op may cross an exchange exactly when it touches neither -- and because
            // skipped ops stay in the window as placeholders, every rank builds the same graph and takes the same
            // decisions, which is what keeps the two partners of an exchange consistent.  Exchanges keep their order.
            uint64_t xmask = 0, zmask = op.cmask | ((uint64_t)op.grefs << 56);
            if (op.kind == kRopExchange) {
                xmask = (1ull << op.target) | (1ull << (56 + op.t2));
                zmask = 0;
                if (last_exchange >= 0 && stamp[last_exchange] != i) { stamp[last_exchange] = i; succ[last_exchange].push_back(i); ++npred[i]; }
                last_exchange = i;
            }
            else if (op.kind == SPZ_GATE_SWAP) xmask = (1ull << op.target) | (1ull << op.t2);
            else if (!diag) xmask = 1ull << op.target;
            else if (op.const_hi < 0 && op.target >= 0) zmask |= 1ull << op.target;
            auto dep = [&](int j) {
                if (stamp[j] == i) return;
                stamp[j] = i;
                succ[j].push_back(i);
                ++npred[i];
            };
            for (int q = 0; q < 64; ++q) {
                if ((xmask >> q) & 1ull) {
                    if (last_w[q] >= 0) dep(last_w[q]);
                    for (int r : zl[q]) dep(r);
                } else if ((zmask >> q) & 1ull) {
                    if (last_w[q] >= 0) dep(last_w[q]);
                }
            }
            for (int q = 0; q < 64; ++q) {
                if ((xmask >> q) & 1ull) { last_w[q] = i; zl[q].clear(); }
                else if ((zmask >> q) & 1ull) zl[q].push_back(i);
            }
        }
        std::vector<char> done(N, 0);
        int remaining = N;
        // ---- choice of the tile for one pass ----
        // For a candidate tile (low qubits 0..L-1 plus a high set) one forward scan over the window counts the ops that
        // could run in it: an op runs if its non-diagonal targets are tile qubits and all its predecessors run (or have
        // run).  The tile is grown greedily for every admissible L -- add the high qubit with the largest gain -- and the
        // best (L, high set) wins; ties go to the longer contiguous segment.  Without this the first ready ops in program
        // order fix the tile, which on layered circuits strands most passes with a dozen gates.
        const int n_local = st->n; // qubits held by this handle (for a shard: its local qubits)
        std::vector<std::vector<int>> pred(N);
        for (int j = 0; j < N; ++j) for (int i : succ[j]) pred[i].push_back(j);
        std::vector<uint64_t> xm(N, 0);
        for (int i = 0; i < N; ++i) {
            const ROp &op = pending[i];
            if (op.kind == kRopExchange) xm[i] = ~0ull; // never part of a pass: its successors wait for it
            else if (op.kind == SPZ_GATE_SWAP) xm[i] = (1ull << op.target) | (1ull << op.t2);
            else if (!is_diagonal_kind(op.kind)) xm[i] = 1ull << op.target;
        }
        std::vector<char> runs(N, 0);
        int first_undone = 0;
        // Choosing costs ~0.1-0.2 ms of host time per pass (hidden behind the previous pass's kernel on large registers, but
        // more than a whole pass on small ones), so by default it is used from 24 local qubits up.  SPZ_TILE_SELECT=0|1
        // forces it off / on.
        bool select_tiles = n_local >= 24;
        if (const char *e = std::getenv("SPZ_TILE_SELECT")) select_tiles = e[0] != '0';
        const int kScan = 512; // evaluation horizon in ops (the rest of the window waits for a later pass)
        const int lstep = 3;   // candidate segment lengths: L = T, T-3, ... and Lmin
        auto count_tile = [&](uint64_t tile) -> int {
            int c = 0;
            const int end = std::min(N, first_undone + kScan);
            for (int i = first_undone; i < end; ++i) {
                runs[i] = 0;
                if (done[i]) continue;
                if (xm[i] & ~tile) continue;
                bool ok = true;
                for (int p : pred[i]) if (!done[p] && !(p >= first_undone && runs[p])) { ok = false; break; }
                if (!ok) continue;
                runs[i] = 1;
                ++c;
            }
            return c;
        };
        auto choose_tile = [&](uint64_t &best_high) -> int {
            while (first_undone < N && done[first_undone]) ++first_undone;
            uint64_t cand = 0; // qubits some undone non-diagonal op within the horizon targets
            const int end = std::min(N, first_undone + kScan);
            for (int i = first_undone; i < end; ++i) if (!done[i] && pending[i].kind != kRopExchange) cand |= xm[i];
            int best_count = -1;
            best_high = 0;
            for (int L = T; L >= Lmin; L -= (L - lstep >= Lmin || L == Lmin ? lstep : L - Lmin)) {
                const uint64_t low = L >= 64 ? ~0ull : ((1ull << L) - 1ull);
                uint64_t high = 0;
                int cur = count_tile(low);
                for (int h = 0; h < T - L; ++h) {
                    int gain_best = -1, q_best = -1;
                    for (int q = L; q < n_local; ++q) {
                        if (!((cand >> q) & 1ull) || ((high >> q) & 1ull)) continue;
                        const int c = count_tile(low | high | (1ull << q));
                        if (c > gain_best) { gain_best = c; q_best = q; }
                    }
                    if (q_best < 0) { // no useful qubit left: pad (the kernel derives L from the number of high qubits)
                        for (int q = n_local - 1; q >= L; --q) if (!((high >> q) & 1ull)) { q_best = q; break; }
                        if (q_best < 0) break;
                        gain_best = cur;
                    }
                    high |= 1ull << q_best;
                    cur = gain_best;
                }
                if (__builtin_popcountll(high) != T - L) continue;
                if (cur > best_count) { best_count = cur; best_high = high; } // strict: ties keep the longer segment
            }
            return best_count;
        };
        int next_exchange = 0; // exchanges run in window order, as soon as everything they depend on has run: the ops that
                               // remain commute with them and can only gain company from what the exchange unlocks
        while (remaining > 0) {
            while (next_exchange < N && (pending[next_exchange].kind != kRopExchange || done[next_exchange])) ++next_exchange;
            if (next_exchange < N && npred[next_exchange] == 0) {
                const int gi = next_exchange + 1;
                const bool fuse_g = gi < N && !done[gi] && fusable_trigger(next_exchange, gi, npred[gi]);
                SPZ_TRY(run_exchange(pending[next_exchange], fuse_g ? &pending[gi] : nullptr));
                done[next_exchange] = 1;
                --remaining;
                for (int sidx : succ[next_exchange]) --npred[sidx];
                if (fuse_g) {
                    done[gi] = 1;
                    --remaining;
                    for (int sidx : succ[gi]) --npred[sidx];
                }
                continue;
            }
            {
                // One-qubit growth cannot see an op that needs two new high qubits at once (a SWAP between them); if no
                // tile admits any op, fall back to letting the first ready ops claim their qubits, which always progresses.
                uint64_t h = 0;
                const bool chosen = select_tiles && choose_tile(h) > 0;
                high_set = chosen ? h : 0;
                low_need = 0;
                tile_fixed = chosen;
            }
            // Fill one group.  Ready ops are taken in ascending scans (successors always have larger indices, and
            // tile capacity only shrinks).  Non-diagonal targets are taken in CLUSTERS of at most 4 distinct qubits
            // -- the register layout of the tile kernel -- so that the compiled program changes layout once per
            // cluster instead of once per gate; a scan that takes nothing opens the next cluster, and a fresh
            // cluster that still takes nothing closes the group.
            uint64_t cluster = 0;
            bool fresh_cluster = true;
            for (;;) {
                bool progress = false;
                for (int i = 0; i < N; ++i) {
                    if (done[i] || npred[i] != 0 || pending[i].kind == kRopExchange) continue;
                    if (ops.size() >= (size_t)kMaxTileGroups / 2) break;
                    const ROp &op = pending[i];
                    uint64_t want = 0; // non-diagonal targets this op needs register-resident
                    if (op.kind == SPZ_GATE_SWAP) want = (1ull << op.target) | (1ull << op.t2);
                    else if (!is_diagonal_kind(op.kind)) want = 1ull << op.target;
                    if (want && __builtin_popcountll(cluster | want) > 4) continue;
                    uint64_t hs; int ln;
                    if (!fits(op, hs, ln)) continue;
                    high_set = hs; low_need = ln;
                    cluster |= want;
                    ops.push_back(op);
                    done[i] = 1;
                    --remaining;
                    for (int sidx : succ[i]) --npred[sidx];
                    progress = true;
                }
                if (progress) { fresh_cluster = false; continue; }
                if (fresh_cluster || ops.size() >= (size_t)kMaxTileGroups / 2 || remaining == 0) break;
                cluster = 0;
                fresh_cluster = true;
            }
            tile_fixed = false;
            if (ops.empty()) {
                set_error("internal: scheduler made no progress"); return SPZ_ERR_INVALID_ARG; }
            SPZ_TRY(emit_group());
        }
        return SPZ_OK;
    }
};

} // namespace spz

using namespace spz;

extern "C" {

int spz_abi_version(void) { return SPZ_ABI_VERSION; }
const char *spz_last_error(void) { return g_err; }

const char *spz_status_string(int status) {
    switch (status) {
    case SPZ_OK: return "ok";
    case SPZ_ERR_INVALID_ARG: return "invalid argument";
    case SPZ_ERR_UNSUPPORTED: return "unsupported gate/control combination";
    case SPZ_ERR_CUDA: return "CUDA error";
    case SPZ_ERR_OOM: return "out of device memory";
    case SPZ_ERR_COMM: return "communication error";
    case SPZ_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
    }
}

int spz_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

int spz_device_name(int device, char *buf, int buflen) {
    cudaDeviceProp prop;
    SPZ_CUDA(cudaGetDeviceProperties(&prop, device));
    snprintf(buf, (size_t)buflen, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    return SPZ_OK;
}

int spz_mem_info(int device, uint64_t *free_bytes, uint64_t *total_bytes) {
    SPZ_CUDA(cudaSetDevice(device));
    size_t f = 0, t = 0;
    SPZ_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return SPZ_OK;
}

int spz_create(int n_qubits, int device, spz_state **out) {
    if (!out) { set_error("null out pointer"); return SPZ_ERR_INVALID_ARG; }
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > 40) { set_error("n_qubits must be in 1..40 (assert!(n > 0) core.rs:33)"); return SPZ_ERR_INVALID_ARG; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: libspinoza_b200 has no CPU fallback");
        return SPZ_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (%d visible)", device, count); return SPZ_ERR_INVALID_ARG; }
    SPZ_CUDA(cudaSetDevice(device));
    spz_state *st = new (std::nothrow) spz_state();
    if (!st) return SPZ_ERR_OOM;
    st->n = n_qubits;
    st->device = device;
    st->len = (int64_t)1 << n_qubits;
    const size_t bytes = sizeof(double) * (size_t)st->len;
    auto fail = [&](cudaError_t err, const char *what) {
        int rc = cuda_fail(err, what, __FILE__, __LINE__);
        spz_destroy(st);
        return rc;
    };
    if ((e = cudaMalloc(&st->re, bytes)) != cudaSuccess) return fail(e, "cudaMalloc(re)");
    if ((e = cudaMalloc(&st->im, bytes)) != cudaSuccess) return fail(e, "cudaMalloc(im)");
    if ((e = cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&st->ev0)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&st->ev1)) != cudaSuccess) return fail(e, "cudaEventCreate");
    // All scratch is allocated here, never lazily: cudaMalloc synchronises the whole device, which must not
    // happen while another shard's handshake kernel is spinning on the same GPU.
    int rc = ensure_scratch(st);
    if (rc == SPZ_OK) rc = tile_prepare(st);
    if (rc == SPZ_OK) rc = launch_fill_basis(st, 0);
    if (rc != SPZ_OK) { spz_destroy(st); return rc; }
    *out = st;
    return SPZ_OK;
}

int spz_destroy(spz_state *st) {
    if (!st) return SPZ_OK;
    cudaSetDevice(st->device);
    if (st->arrival.copy) { cudaStreamSynchronize(st->arrival.copy); cudaStreamDestroy(st->arrival.copy); }
    for (cudaStream_t l : st->arrival.lane) if (l) { cudaStreamSynchronize(l); cudaStreamDestroy(l); }
    for (cudaEvent_t e : st->arrival.ev) if (e) cudaEventDestroy(e);
    if (st->arrival.ready) cudaEventDestroy(st->arrival.ready);
    st->arrival.pending = false;
    if (st->dist) dist_join(st);
    if (st->stream) cudaStreamSynchronize(st->stream);
    if (st->dist) dist_destroy(st);
    cudaFree(st->re);
    cudaFree(st->im);
    cudaFree(st->scratch.partials);
    if (st->scratch.samp) cudaFree(st->scratch.samp); // (allocated stream-ordered; the stream has been synchronised above)
    if (st->scratch.h_result) cudaFreeHost(st->scratch.h_result);
    cudaFree(st->d_ops);
    if (st->ev0) cudaEventDestroy(st->ev0);
    if (st->ev1) cudaEventDestroy(st->ev1);
    if (st->stream) cudaStreamDestroy(st->stream);
    cudaGetLastError();
    delete st;
    return SPZ_OK;
}

int spz_clone(const spz_state *src, spz_state **out) {
    if (!src || !out) { set_error("null argument"); return SPZ_ERR_INVALID_ARG; }
    if (src->dist) { set_error("clone of a sharded register is not supported"); return SPZ_ERR_UNSUPPORTED; }
    SPZ_TRY(spz_create(src->n, src->device, out));
    spz_state *dst = *out;
    const size_t bytes = sizeof(double) * (size_t)src->len;
    SPZ_CUDA(cudaStreamSynchronize(src->stream));
    SPZ_CUDA(cudaMemcpyAsync(dst->re, src->re, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    SPZ_CUDA(cudaMemcpyAsync(dst->im, src->im, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    SPZ_CUDA(cudaStreamSynchronize(dst->stream));
    dst->rng = src->rng;
    return SPZ_OK;
}

int spz_num_qubits(const spz_state *st) { return st ? total_qubits(st) : -1; }
int64_t spz_len(const spz_state *st) { return st ? st->len : -1; }

int spz_reset_zero(spz_state *st) { SPZ_CHECK_STATE(st); return st->dist ? dist_fill_basis(st, 0) : launch_fill_basis(st, 0); }
int spz_set_basis(spz_state *st, uint64_t index) { SPZ_CHECK_STATE(st); return st->dist ? dist_fill_basis(st, index) : launch_fill_basis(st, index); }
int spz_init_random(spz_state *st, uint64_t seed) { SPZ_CHECK_STATE(st); return st->dist ? dist_init_random(st, seed) : launch_init_random(st, seed); }
int spz_set_seed(spz_state *st, uint64_t seed) { if (!st) return SPZ_ERR_INVALID_ARG; st->rng = seed; return SPZ_OK; }

int spz_upload(spz_state *st, const double *re, const double *im, int64_t offset, int64_t count) {
    SPZ_CHECK_STATE(st);
    if (offset < 0 || count < 0 || offset + count > st->len) { set_error("upload range [%lld, +%lld) outside the state", (long long)offset, (long long)count); return SPZ_ERR_INVALID_ARG; }
    SPZ_TRY(join_pending(st));
    st->arrival.streaming = false;
    dist_note_modified(st);
    const size_t bytes = sizeof(double) * (size_t)count;
    if (re) SPZ_CUDA(cudaMemcpyAsync(st->re + offset, re, bytes, cudaMemcpyHostToDevice, st->stream));
    if (im) SPZ_CUDA(cudaMemcpyAsync(st->im + offset, im, bytes, cudaMemcpyHostToDevice, st->stream));
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    return SPZ_OK;
}

// Whole-state upload that returns at once: K pieces on a copy stream, each with its event, so that the gates issued next run
// piece by piece behind the bus (launch_gate / launch_tile_program take the chunk events) instead of after the last byte.
// The host buffers must be page-locked (spz_alloc_host) and stay untouched until the next spz_sync / spz_download.
int spz_upload_async(spz_state *st, const double *re, const double *im) {
    if (!st || !re || !im) { set_error("spz_upload_async: null argument"); return SPZ_ERR_INVALID_ARG; }
    dist_note_modified(st);
    SPZ_CUDA(cudaSetDevice(st->device));
    SPZ_TRY(join_pending(st));
    spz_state::Arrival &a = st->arrival;
    // pieces: more of them shorten the lag of the first / last piece behind the bus, but every piece bit is a qubit whose
    // non-diagonal gates need the whole state (SPZ_UPLOAD_CHUNKS = 2, 4 or 8; measured in profiles/round2_summary.md)
    static const int want_chunks = [] { const char *e = std::getenv("SPZ_UPLOAD_CHUNKS"); const int v = e ? std::atoi(e) : 0; return (v == 2 || v == 4 || v == 8) ? v : 4; }();
    int K = want_chunks;
    while (K > 1 && (st->len >> 12) < K) K >>= 1;
    if (!a.copy) {
        SPZ_CUDA(cudaStreamCreateWithFlags(&a.copy, cudaStreamNonBlocking));
        for (cudaEvent_t &e : a.ev) SPZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        SPZ_CUDA(cudaEventCreateWithFlags(&a.ready, cudaEventDisableTiming));
        int least = 0, greatest = 0;
        SPZ_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest)); // (numerically: greatest <= least)
        for (int k = 0; k < 8; ++k) SPZ_CUDA(cudaStreamCreateWithPriority(&a.lane[k], cudaStreamNonBlocking, std::min(least, greatest + k)));
    }
    SPZ_CUDA(cudaEventRecord(a.ready, st->stream));
    SPZ_CUDA(cudaStreamWaitEvent(a.copy, a.ready, 0));
    const size_t piece = (size_t)st->len / (size_t)K;
    for (int k = 0; k < K; ++k) {
        SPZ_CUDA(cudaMemcpyAsync(st->re + k * piece, re + k * piece, sizeof(double) * piece, cudaMemcpyHostToDevice, a.copy));
        SPZ_CUDA(cudaMemcpyAsync(st->im + k * piece, im + k * piece, sizeof(double) * piece, cudaMemcpyHostToDevice, a.copy));
        SPZ_CUDA(cudaEventRecord(a.ev[k], a.copy));
    }
    a.chunks = K;
    a.pending = true;
    a.streaming = true;
    return SPZ_OK;
}

int spz_download(const spz_state *st, double *re, double *im, int64_t offset, int64_t count) {
    SPZ_CHECK_STATE(st);
    if (offset < 0 || count < 0 || offset + count > st->len) { set_error("download range [%lld, +%lld) outside the state", (long long)offset, (long long)count); return SPZ_ERR_INVALID_ARG; }
    {
        // The end of a streamed round trip (spz_upload_async ... gates ... spz_download of the whole state): every piece leaves
        // from its own lane as soon as that lane has finished its run, while the later pieces are still being computed.
        spz_state::Arrival &a = const_cast<spz_state *>(st)->arrival;
        a.streaming = false;
        if (a.lanes_active && !a.pending && offset == 0 && count == st->len && re && im) {
            const size_t piece = (size_t)st->len / (size_t)a.chunks;
            for (int k = 0; k < a.chunks; ++k) {
                SPZ_CUDA(cudaMemcpyAsync(re + k * piece, st->re + k * piece, sizeof(double) * piece, cudaMemcpyDeviceToHost, a.lane[k]));
                SPZ_CUDA(cudaMemcpyAsync(im + k * piece, st->im + k * piece, sizeof(double) * piece, cudaMemcpyDeviceToHost, a.lane[k]));
            }
            for (int k = 0; k < a.chunks; ++k) SPZ_CUDA(cudaStreamSynchronize(a.lane[k]));
            a.lanes_active = false;
            return SPZ_OK;
        }
    }
    SPZ_TRY(join_pending(const_cast<spz_state *>(st)));
    const size_t bytes = sizeof(double) * (size_t)count;
    if (re) SPZ_CUDA(cudaMemcpyAsync(re, st->re + offset, bytes, cudaMemcpyDeviceToHost, st->stream));
    if (im) SPZ_CUDA(cudaMemcpyAsync(im, st->im + offset, bytes, cudaMemcpyDeviceToHost, st->stream));
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    return SPZ_OK;
}

int spz_sync(spz_state *st) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(join_pending(st));
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    return SPZ_OK;
}

int spz_alloc_host(uint64_t bytes, void **out) {
    if (!out) return SPZ_ERR_INVALID_ARG;
    *out = nullptr;
    SPZ_CUDA(cudaMallocHost(out, (size_t)bytes));
    return SPZ_OK;
}

int spz_free_host(void *ptr) {
    if (ptr) SPZ_CUDA(cudaFreeHost(ptr));
    return SPZ_OK;
}

// ---- gates ---------------------------------------------------------------------------------------------
int spz_apply(spz_state *st, const spz_gate *gate, int target) {
    SPZ_CHECK_STATE(st);
    if (!gate) { set_error("null gate"); return SPZ_ERR_INVALID_ARG; }
    switch (gate->kind) {
    case SPZ_GATE_SWAP: return swap_impl(st, gate->t0, gate->t1); // gates.rs:225 (ignores `target`)
    case SPZ_GATE_BITFLIP: { // bit_flip_noise_apply gates.rs:1365-1374
        if (target < 0 || target >= total_qubits(st)) { set_error("target %d out of range", target); return SPZ_ERR_INVALID_ARG; }
        const double eps = next_u01(st);
        if (eps <= gate->p[0]) return apply_masked(st, SPZ_GATE_X, nullptr, 0, target);
        return SPZ_OK; }
    case SPZ_GATE_M: case SPZ_GATE_UNITARY: // unimplemented!() gates.rs:230
        set_error("apply: gate kind %d is not applicable here (reference: unimplemented!())", gate->kind);
        return SPZ_ERR_UNSUPPORTED;
    default: return apply_masked(st, gate->kind, gate->p, 0, target);
    }
}

static int controlled_ok(const spz_gate *gate, const char *fn) {
    if (!gate) { set_error("null gate"); return SPZ_ERR_INVALID_ARG; }
    switch (gate->kind) {
    case SPZ_GATE_M: case SPZ_GATE_SWAP: case SPZ_GATE_UNITARY: case SPZ_GATE_BITFLIP: // todo!() gates.rs:267,275,318
        set_error("%s: gate kind %d has no controlled form (reference: todo!())", fn, gate->kind);
        return SPZ_ERR_UNSUPPORTED;
    default: return SPZ_OK;
    }
}

int spz_c_apply(spz_state *st, const spz_gate *gate, int control, int target) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(controlled_ok(gate, "c_apply"));
    if (control < 0 || control >= total_qubits(st) || control == target) { set_error("bad control %d (target %d, %d qubits)", control, target, total_qubits(st)); return SPZ_ERR_INVALID_ARG; }
    return apply_masked(st, gate->kind, gate->p, 1ull << control, target);
}

int spz_cc_apply(spz_state *st, const spz_gate *gate, int c0, int c1, int target) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(controlled_ok(gate, "cc_apply"));
    if (c0 < 0 || c1 < 0 || c0 >= total_qubits(st) || c1 >= total_qubits(st) || c0 == target || c1 == target) { set_error("bad controls (%d,%d)", c0, c1); return SPZ_ERR_INVALID_ARG; }
    return apply_masked(st, gate->kind, gate->p, (1ull << c0) | (1ull << c1), target);
}

int spz_mc_apply_mask(spz_state *st, const spz_gate *gate, uint64_t ctrl_mask, int target) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(controlled_ok(gate, "mc_apply"));
    return apply_masked(st, gate->kind, gate->p, ctrl_mask, target);
}

int spz_mc_apply_signed(spz_state *st, const spz_gate *gate, uint64_t ones_mask, uint64_t zeros_mask, int target) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(controlled_ok(gate, "mc_apply_signed"));
    return apply_signed(st, gate->kind, gate->p, ones_mask, zeros_mask, target);
}

int spz_mc_apply(spz_state *st, const spz_gate *gate, const int32_t *controls, int n_controls, const int32_t *zeros,
                 int n_zeros, int target) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(controlled_ok(gate, "mc_apply"));
    if (n_controls < 0 || (n_controls && !controls) || !(total_qubits(st) > n_controls)) { // debug_assert gates.rs:297
        set_error("mc_apply needs n_qubits > n_controls"); return SPZ_ERR_INVALID_ARG;
    }
    uint64_t mask = 0; // gates.rs:298-311: controls listed in `zeros` are dropped from the mask
    for (int i = 0; i < n_controls; ++i) {
        const int c = controls[i];
        if (c < 0 || c >= total_qubits(st) || c == target) { set_error("bad control %d", c); return SPZ_ERR_INVALID_ARG; }
        bool skip = false;
        for (int j = 0; j < n_zeros; ++j) if (zeros && zeros[j] == c) skip = true;
        if (!skip) mask |= 1ull << c;
    }
    return apply_masked(st, gate->kind, gate->p, mask, target);
}

static int execute_impl(spz_state *st, const spz_op *ops, int64_t n_ops, uint32_t flags, uint64_t *measured_mask,
                        uint64_t *measured_vals, PlanSink *sink);

int spz_iqft(spz_state *st, const int32_t *targets, int m) {
    SPZ_CHECK_STATE(st);
    if (m < 0 || (m && !targets)) return SPZ_ERR_INVALID_ARG;
    // The m + m(m-1)/2 gates of core.rs:184-191 as ONE op list through the fused scheduler: a run of them shares one HBM pass
    // (QFT-30: 4 passes instead of 465).  Default: merged mode, like QuantumCircuit::execute -- within 1e-12 of the gate-by-gate
    // loop (the documented fused rounding difference; 30 qubits: 38 ms against ~700 ms).  SPZ_IQFT_FUSE=exact keeps every gate's
    // reference arithmetic (bit-identical to the loop, 550 ms at 30 qubits: the exact tile kernel pays for every controlled
    // phase on its own), SPZ_IQFT_FUSE=0 keeps the loop; so do small registers, where a pass is launch latency anyway.
    const char *iqft_env = std::getenv("SPZ_IQFT_FUSE");
    const bool fuse_iqft = !(iqft_env && iqft_env[0] == '0');
    const uint32_t iqft_flags = SPZ_EXEC_FUSE | ((iqft_env && iqft_env[0] == 'e') ? SPZ_EXEC_EXACT : 0u);
    if (fuse_iqft && m >= 3 && st->n >= 16) {
        const int nq = total_qubits(st);
        std::vector<spz_op> ops;
        ops.reserve((size_t)m + (size_t)m * (m - 1) / 2);
        for (int j = m - 1; j >= 0; --j) {
            if (targets[j] < 0 || targets[j] >= nq) { set_error("bad iqft targets"); return SPZ_ERR_INVALID_ARG; }
            spz_op h{};
            h.kind = SPZ_GATE_H; h.target = targets[j]; h.ctrl_kind = SPZ_CTRL_NONE;
            ops.push_back(h);
            for (int k = j - 1; k >= 0; --k) {
                if (targets[j] == targets[k] || targets[k] < 0 || targets[k] >= nq) { set_error("bad iqft targets"); return SPZ_ERR_INVALID_ARG; }
                spz_op c{};
                c.kind = SPZ_GATE_P; c.target = targets[k]; c.p[0] = -3.14159265358979323846 / std::ldexp(1.0, j - k);
                c.ctrl_kind = SPZ_CTRL_SINGLE; c.ctrl_mask = 1ull << targets[j];
                ops.push_back(c);
            }
        }
        return execute_impl(st, ops.data(), (int64_t)ops.size(), iqft_flags, nullptr, nullptr, nullptr);
    }
    // core.rs:184-191
    for (int j = m - 1; j >= 0; --j) {
        SPZ_TRY(apply_masked(st, SPZ_GATE_H, nullptr, 0, targets[j]));
        for (int k = j - 1; k >= 0; --k) {
            const double ang = -3.14159265358979323846 / std::ldexp(1.0, j - k);
            if (targets[j] == targets[k] || targets[j] < 0 || targets[j] >= total_qubits(st)) { set_error("bad iqft targets"); return SPZ_ERR_INVALID_ARG; }
            SPZ_TRY(apply_masked(st, SPZ_GATE_P, &ang, 1ull << targets[j], targets[k]));
        }
    }
    return SPZ_OK;
}

// ---- execute -------------------------------------------------------------------------------------------
static int execute_impl(spz_state *st, const spz_op *ops, int64_t n_ops, uint32_t flags, uint64_t *measured_mask,
                        uint64_t *measured_vals, PlanSink *sink) {
    if (n_ops < 0 || (n_ops && !ops)) { set_error("bad op list"); return SPZ_ERR_INVALID_ARG; }
    uint64_t local_m = 0, local_v = 0;
    uint64_t *mm = measured_mask ? measured_mask : &local_m;
    uint64_t *mv = measured_vals ? measured_vals : &local_v;
    const bool fuse = (flags & SPZ_EXEC_FUSE) != 0 && st->n >= min_tile_bits();
    Fuser fuser(st);
    fuser.exact = (flags & SPZ_EXEC_EXACT) != 0;
    fuser.reorder = (flags & SPZ_EXEC_KEEP_ORDER) == 0;
    fuser.sink = sink;

    const int nq = total_qubits(st);
    int64_t cur = 0; // index of the op being emitted (for the exchange look-ahead)
    // distance from op `cur` to the next non-diagonal use of every logical qubit (sharded registers only): the positions of
    // every qubit's non-diagonal uses are collected once, and a cursor per qubit moves forward with `cur` -- O(n_ops) for the
    // whole list instead of a rescan of the remaining ops for every emitted op
    std::vector<int64_t> uses[64];
    size_t use_cursor[64] = {0};
    if (st->dist)
        for (int64_t j = 0; j < n_ops; ++j) {
            const spz_op &o = ops[j];
            if (o.kind == SPZ_GATE_SWAP || is_diagonal_kind(o.kind) || o.kind == SPZ_GATE_UNITARY) continue;
            if (o.target >= 0 && o.target < 64) uses[o.target].push_back(j);
        }
    if (st->dist && n_ops > 0) { // a register that is still a basis state: its qubit permutation is free, choose it from the list
        int64_t first_use[64];
        for (int q = 0; q < 64; ++q) first_use[q] = uses[q].empty() ? INT64_MAX : uses[q][0];
        bool changed = false;
        SPZ_TRY(dist_place_basis(st, first_use, sink != nullptr, &changed));
        if (changed && sink) {
            int32_t perm[64];
            for (int q = 0; q < 64; ++q) perm[q] = q;
            spz_dist_perm(st, perm);
            sink->relabel(perm);
        }
    }
    auto next_uses = [&](uint64_t *nu) {
        for (int q = 0; q < 64; ++q) {
            size_t &c = use_cursor[q];
            while (c < uses[q].size() && uses[q][c] <= cur) ++c;
            nu[q] = c < uses[q].size() ? (uint64_t)(uses[q][c] - cur) : UINT64_MAX;
        }
    };

    // Sharded registers: keep exchanges (and the ops this rank skips, as placeholders) inside the scheduling window instead
    // of closing the window at every exchange.  On by default (SPZ_DIST_WINDOW=0 restores the old windows): checked on the CPU by
    // tests/test_dist_fused_cpu.py, on one GPU by the local-group tests and on two GPUs over NVLink against the oracle
    // (bench.py's `parity` extra; QFT-31 on 2 GPUs 65.5 -> 61.5 ms, profiles/round2_summary.md).
    bool window_exchanges = false;
    if (fuse && st->dist && fuser.reorder && !fuser.exact) {
        window_exchanges = true;
        if (const char *e = std::getenv("SPZ_DIST_WINDOW")) window_exchanges = e[0] != '0';
    }

    auto emit_local = [&](int kind, const double *p, uint64_t cmask, int target, int t2, int const_hi, bool skip = false,
                          uint32_t grefs = 0) -> int {
        if (!fuse && !sink) {
            if (kind == SPZ_GATE_SWAP) return launch_swap(st, target, t2);
            GateK g;
            SPZ_TRY(resolve_gate(kind, p, &g));
            if (const_hi >= 0) return dist_diag_const(st, g, cmask, const_hi);
            return launch_gate(st, g, cmask, target);
        }
        ROp r{};
        r.kind = kind; r.target = target; r.t2 = t2; r.cmask = cmask; r.const_hi = const_hi;
        r.theta = p ? p[0] : 0.0;
        r.src = (int)cur;
        r.skip = skip; r.grefs = grefs;
        if (!fuse) { // dry run of the unfused path: one pass per op
            if (kind != SPZ_GATE_SWAP) SPZ_TRY(resolve_gate(kind, p, &r.g));
            sink->take(std::vector<ROp>{r});
            return SPZ_OK;
        }
        if (kind == SPZ_GATE_SWAP) { if (target == t2) return SPZ_OK; r.g.kind = kind; }
        else SPZ_TRY(resolve_gate(kind, p, &r.g));
        return fuser.push(r);
    };

    auto emit = [&](int kind, const double *p, uint64_t cmask, int target, int t2) -> int {
        if (kind != SPZ_GATE_SWAP) {
            if (target < 0 || target >= nq) { set_error("target %d out of range", target); return SPZ_ERR_INVALID_ARG; }
            if ((cmask >> target) & 1ull) { set_error("target %d is also a control", target); return SPZ_ERR_INVALID_ARG; }
            if (nq < 64 && (cmask >> nq)) { set_error("control outside the register"); return SPZ_ERR_INVALID_ARG; }
        } else if (target < 0 || t2 < 0 || target >= nq || t2 >= nq) {
            set_error("swap operands out of range"); return SPZ_ERR_INVALID_ARG;
        }
        if (!st->dist) return emit_local(kind, p, cmask, target, t2, -1);
        // sharded register: lower the logical op to per-rank actions (dist_plan.h)
        if (kind != SPZ_GATE_SWAP) { GateK probe; SPZ_TRY(resolve_gate(kind, p, &probe)); }
        uint64_t nu[64];
        next_uses(nu);
        std::vector<spz_dist_action> acts;
        int rc = dist_lower(st, kind, p, target, t2, cmask, target, nu, acts);
        if (rc != SPZ_OK) { set_error("cannot lower gate kind %d onto the sharded register", kind); return rc; }
        for (size_t ai = 0; ai < acts.size(); ++ai) {
            const spz_dist_action &a = acts[ai];
            // exchange + the uncontrolled gate that asked for it as one kernel (opt-in; not inside exchange-spanning windows,
            // where the exchange is a node of the scheduling graph)
            if (!window_exchanges && a.type == ACT_EXCHANGE && ai + 1 < acts.size() && acts[ai + 1].type == ACT_LOCAL_GATE &&
                dist_can_fuse_gate(st, acts[ai + 1].kind, cmask, acts[ai + 1].target, a.lq)) {
                GateK g;
                SPZ_TRY(resolve_gate(acts[ai + 1].kind, acts[ai + 1].p, &g));
                SPZ_TRY(fuser.flush());
                if (sink) { // dry run: an exchange step followed by the gate as a step of its own describes the same thing
                    sink->exchange(a.gbit, a.lq);
                    ROp r{};
                    r.kind = acts[ai + 1].kind; r.target = acts[ai + 1].target; r.g = g; r.src = (int)cur;
                    sink->take(std::vector<ROp>{r});
                } else {
                    SPZ_TRY(dist_exchange_gate(st, a.gbit, a.lq, g));
                }
                ++ai;
                continue;
            }
            switch (a.type) {
            case ACT_SKIP:
                if (window_exchanges) // placeholder with the real op's shape
                    SPZ_TRY(emit_local(a.kind, a.p, a.cmask, a.target < 0 ? 0 : a.target, 0, a.target < 0 ? a.hi : -1, true, (uint32_t)a.grefs));
                break;
            case ACT_EXCHANGE:
                if (window_exchanges) {
                    ROp x{};
                    x.kind = kRopExchange; x.target = a.lq; x.t2 = a.gbit; x.src = (int)cur;
                    SPZ_TRY(fuser.push(x));
                    break;
                }
                SPZ_TRY(fuser.flush());
                if (sink) sink->exchange(a.gbit, a.lq); // dry run: the plan's permutation has already been updated
                else SPZ_TRY(dist_exchange(st, a.gbit, a.lq));
                break;
            case ACT_LOCAL_GATE: SPZ_TRY(emit_local(a.kind, a.p, a.cmask, a.target, 0, -1, false, (uint32_t)a.grefs)); break;
            case ACT_DIAG_CONST: SPZ_TRY(emit_local(a.kind, a.p, a.cmask, 0, 0, a.hi, false, (uint32_t)a.grefs)); break;
            default: return SPZ_ERR_INVALID_ARG;
            }
        }
        return SPZ_OK;
    };

    // A run of consecutive M ops (typically "measure everything" at the end of a circuit): every measurement resets its qubit
    // to |0>, so the later ones of the run visit only the subspace where the earlier qubits are 0 (kernels_reduce.cu).
    uint64_t run_zero = 0;
    int64_t run_end = -2; // index of the last op of the run so far (a measurement that was skipped as already measured counts)
    for (int64_t i = 0; i < n_ops; ++i) {
        const spz_op &op = ops[i];
        const int kind = op.kind;
        cur = i;
        if (kind == SPZ_GATE_UNITARY) { // transform_u / c_transform_u (circuit.rs:555-558): dense fallback, out of scope
            set_error("execute: dense Unitary gates are not supported by the B200 engine");
            return SPZ_ERR_UNSUPPORTED;
        }
        if (op.ctrl_kind == SPZ_CTRL_NONE) {
            if (kind == SPZ_GATE_M) { // circuit.rs:559-566
                if (op.target < 0 || op.target >= total_qubits(st)) { set_error("target %d out of range", op.target); return SPZ_ERR_INVALID_ARG; }
                if (!((*mm >> op.target) & 1ull)) {
                    SPZ_TRY(fuser.flush());
                    if (sink) { ROp m{}; m.kind = SPZ_GATE_M; m.target = op.target; m.src = (int)i; sink->take(std::vector<ROp>{m}); continue; }
                    int bit = 0;
                    if (run_end != i - 1) run_zero = 0; // some other op ran since the last measurement: nothing is known to be 0
                    SPZ_TRY(measure_impl(st, op.target, 1, -1, &bit, run_zero));
                    run_zero |= 1ull << op.target; // measured with reset: |0> from here on
                    run_end = i;
                    *mm |= 1ull << op.target;
                    *mv &= ~(1ull << op.target);
                    *mv |= (uint64_t)bit << op.target;
                } else if (run_end == i - 1) {
                    run_end = i; // measured before (nothing happens, circuit.rs:560): the run goes on
                }
            } else if (kind == SPZ_GATE_BITFLIP) { // gates.rs:1365-1374
                const double eps = next_u01(st);
                if (eps <= op.p[0]) SPZ_TRY(emit(SPZ_GATE_X, nullptr, 0, op.target, 0));
            } else if (kind == SPZ_GATE_SWAP) {
                SPZ_TRY(emit(SPZ_GATE_SWAP, nullptr, 0, op.t0, op.t1)); // apply() ignores target for SWAP, gates.rs:225
            } else { // circuit.rs:567-569
                SPZ_TRY(emit(kind, op.p, 0, op.target, 0));
            }
            continue;
        }
        if (kind == SPZ_GATE_M || kind == SPZ_GATE_SWAP || kind == SPZ_GATE_BITFLIP) {
            set_error("execute: gate kind %d has no controlled form (reference: todo!())", kind);
            return SPZ_ERR_UNSUPPORTED;
        }
        if (op.ctrl_kind == SPZ_CTRL_SINGLE) { // circuit.rs:570-578
            if (__builtin_popcountll(op.ctrl_mask) != 1) { set_error("Single control needs exactly one control bit"); return SPZ_ERR_INVALID_ARG; }
            const int c = __builtin_ctzll(op.ctrl_mask);
            if ((*mm >> c) & 1ull) {
                if ((*mv >> c) & 1ull) SPZ_TRY(emit(kind, op.p, 0, op.target, 0)); // classical control
            } else {
                SPZ_TRY(emit(kind, op.p, op.ctrl_mask, op.target, 0));
            }
        } else if (op.ctrl_kind == SPZ_CTRL_ONES) { // circuit.rs:579-587 (X only in the reference)
            SPZ_TRY(emit(kind, op.p, op.ctrl_mask, op.target, 0));
        } else if (op.ctrl_kind == SPZ_CTRL_MIXED) { // circuit.rs:588-596 -> mc_apply drops the zeros (gates.rs:298-311)
            SPZ_TRY(emit(kind, op.p, op.ctrl_mask & ~op.zeros_mask, op.target, 0));
        } else if (op.ctrl_kind == SPZ_CTRL_SIGNED) { // extension: zeros_mask are true negative controls (a subset of ctrl_mask)
            const uint64_t zeros = op.zeros_mask;
            if (zeros & ~op.ctrl_mask) { set_error("Signed controls: zeros_mask must be a subset of ctrl_mask"); return SPZ_ERR_INVALID_ARG; }
            if (nq < 64 && (op.ctrl_mask >> nq)) { set_error("control outside the register"); return SPZ_ERR_INVALID_ARG; }
            if (zeros && !fuse && !st->dist) { // one launch of the pair kernel (a dry run records one pass of the same shape)
                if (sink) SPZ_TRY(emit(kind, op.p, op.ctrl_mask, op.target, 0));
                else SPZ_TRY(apply_signed(st, kind, op.p, op.ctrl_mask & ~zeros, zeros, op.target));
            } else { // X on the zero-controls around the all-ones form: the X's ride in the same fused passes
                for (int q = 0; q < nq; ++q) if ((zeros >> q) & 1ull) SPZ_TRY(emit(SPZ_GATE_X, nullptr, 0, q, 0));
                SPZ_TRY(emit(kind, op.p, op.ctrl_mask, op.target, 0));
                for (int q = 0; q < nq; ++q) if ((zeros >> q) & 1ull) SPZ_TRY(emit(SPZ_GATE_X, nullptr, 0, q, 0));
            }
        } else {
            set_error("bad ctrl_kind %d", op.ctrl_kind);
            return SPZ_ERR_INVALID_ARG;
        }
    }
    return fuser.flush();
}

int spz_execute(spz_state *st, const spz_op *ops, int64_t n_ops, uint32_t flags, uint64_t *measured_mask,
                uint64_t *measured_vals) {
    SPZ_CHECK_STATE(st);
    return execute_impl(st, ops, n_ops, flags, measured_mask, measured_vals, nullptr);
}

int spz_plan_fusion(int n_qubits, const spz_op *ops, int64_t n_ops, uint32_t flags, int32_t *out_order, int32_t *out_pass,
                    int32_t *out_n_passes) {
    if (n_qubits < 1 || n_qubits > 40 || !out_order || !out_pass || !out_n_passes) { set_error("bad arguments"); return SPZ_ERR_INVALID_ARG; }
    spz_state dummy; // never touches a device: only the qubit count and the seeded generator are read
    dummy.n = n_qubits;
    dummy.len = (int64_t)1 << n_qubits;
    PlanSink sink;
    SPZ_TRY(execute_impl(&dummy, ops, n_ops, flags, nullptr, nullptr, &sink));
    if (sink.order.size() > (size_t)n_ops) { // the output arrays hold one entry per op of the caller's list
        set_error("plan: the list expands to %zu scheduled ops (negative controls become X . op . X in fused passes); expand it before planning", sink.order.size());
        return SPZ_ERR_UNSUPPORTED;
    }
    for (size_t i = 0; i < sink.order.size(); ++i) { out_order[i] = sink.order[i]; out_pass[i] = sink.group[i]; }
    for (size_t i = sink.order.size(); i < (size_t)n_ops; ++i) { out_order[i] = -1; out_pass[i] = -1; }
    *out_n_passes = sink.n_groups;
    return SPZ_OK;
}

// Debug / test hook (pure host code): the micro-program spz_execute would launch for pass `pass_index` of the list.
// Layout of `out` (int32 header then raw structs, all host-endian):
//   [0] status: 0 = tile program, 1 = single op on the direct kernel, 2 = no such pass
//   [1] tile_bits [2] low_bits [3] n_high [4..11] high[8] [12] n_instr [13] n_groups [14] n_terms [15] sizeof(TileInstr)
//   then n_instr TileInstr, n_groups TileGroup (16 B), n_terms TileTerm (32 B).
int spz_debug_compile_pass(int n_qubits, const spz_op *ops, int64_t n_ops, uint32_t flags, int pass_index, void *out,
                           int64_t out_bytes, int64_t *out_used) {
    if (n_qubits < 1 || n_qubits > 40 || !out || !out_used || pass_index < 0) { set_error("bad arguments"); return SPZ_ERR_INVALID_ARG; }
    spz_state dummy;
    dummy.n = n_qubits;
    dummy.len = (int64_t)1 << n_qubits;
    PlanSink sink;
    sink.capture_pass = pass_index;
    SPZ_TRY(execute_impl(&dummy, ops, n_ops, flags, nullptr, nullptr, &sink));
    const size_t need = 16 * sizeof(int32_t) + sizeof(TileInstr) * sink.cap_prog.size() + sizeof(TileGroup) * sink.cap_groups.size() +
                        sizeof(TileTerm) * sink.cap_terms.size();
    if ((int64_t)need > out_bytes) { set_error("output buffer too small: need %zu bytes", need); return SPZ_ERR_INVALID_ARG; }
    int32_t hdr[16] = {0};
    hdr[0] = !sink.captured ? 2 : sink.captured_direct ? 1 : 0;
    hdr[1] = sink.cap_plan.tile_bits; hdr[2] = sink.cap_plan.low_bits; hdr[3] = sink.cap_plan.n_high;
    for (int k = 0; k < 8; ++k) hdr[4 + k] = sink.cap_plan.high[k];
    hdr[12] = (int32_t)sink.cap_prog.size(); hdr[13] = (int32_t)sink.cap_groups.size(); hdr[14] = (int32_t)sink.cap_terms.size();
    hdr[15] = (int32_t)sizeof(TileInstr);
    char *w = static_cast<char *>(out);
    std::memcpy(w, hdr, sizeof hdr); w += sizeof hdr;
    if (!sink.cap_prog.empty()) { std::memcpy(w, sink.cap_prog.data(), sizeof(TileInstr) * sink.cap_prog.size()); w += sizeof(TileInstr) * sink.cap_prog.size(); }
    if (!sink.cap_groups.empty()) { std::memcpy(w, sink.cap_groups.data(), sizeof(TileGroup) * sink.cap_groups.size()); w += sizeof(TileGroup) * sink.cap_groups.size(); }
    if (!sink.cap_terms.empty()) { std::memcpy(w, sink.cap_terms.data(), sizeof(TileTerm) * sink.cap_terms.size()); w += sizeof(TileTerm) * sink.cap_terms.size(); }
    *out_used = (int64_t)need;
    return SPZ_OK;
}

// Debug / test hook (pure host code): everything spz_execute would do on rank `rank` of a register of n_total qubits
// sharded over `world` ranks -- fused passes as micro-programs, single ops, exchanges -- in order.  Serialisation:
//   int64 n_steps, then per step an int32 header h[20] and a payload of h[16] bytes:
//     h[0] = 0 tile program : h[1..3] = tile_bits, low_bits, n_high; h[4..11] = high[8]; h[12..14] = #instr, #groups, #terms;
//                             h[15] = sizeof(TileInstr); payload = instr | groups | terms
//     h[0] = 1 single op    : h[1..4] = kind, target, t2, const_hi; payload = u64 cmask, double s[7]
//     h[0] = 2 exchange     : h[1] = bit of the rank, h[2] = local physical bit it trades places with
//     h[0] = 3 measurement  : h[1] = target
//     h[0] = 4 re-placement : the register was a basis state (SPZ_DEBUG_BASIS) and its qubits were relabelled before the first op
//                             (dist_place_basis); payload = the new permutation, 64 x int32
//   then int32 perm[64] (logical -> physical after the last op).  world = 1 describes an unsharded register.
int spz_debug_compile_sharded(int n_total, int world, int rank, const spz_op *ops, int64_t n_ops, uint32_t flags, void *out,
                              int64_t out_bytes, int64_t *out_used) {
    if (n_total < 1 || n_total > 40 || world < 1 || (world & (world - 1)) || rank < 0 || rank >= world || !out || !out_used) {
        set_error("bad arguments"); return SPZ_ERR_INVALID_ARG;
    }
    int g = 0;
    while ((1 << g) < world) ++g;
    if (n_total - g < 1) { set_error("more ranks than amplitudes"); return SPZ_ERR_INVALID_ARG; }
    spz_state dummy;
    dummy.n = n_total - g;
    dummy.len = (int64_t)1 << dummy.n;
    if (world > 1) dist_debug_attach(&dummy, n_total, world, rank);
    PlanSink sink;
    sink.capture_all = true;
    int rc = execute_impl(&dummy, ops, n_ops, flags, nullptr, nullptr, &sink);
    int32_t perm[64];
    for (int q = 0; q < 64; ++q) perm[q] = q;
    if (world > 1) { spz_dist_perm(&dummy, perm); dist_debug_detach(&dummy); }
    SPZ_TRY(rc);
    char *w = static_cast<char *>(out), *end = w + out_bytes;
    auto put = [&](const void *src, size_t bytes) -> bool {
        if (w + bytes > end) return false;
        std::memcpy(w, src, bytes);
        w += bytes;
        return true;
    };
    const int64_t n_steps = (int64_t)sink.steps.size();
    bool ok = put(&n_steps, sizeof n_steps);
    for (const PlanSink::Step &st : sink.steps) {
        int32_t h[20] = {0};
        h[0] = st.type;
        if (st.type == 0) {
            h[1] = st.plan.tile_bits; h[2] = st.plan.low_bits; h[3] = st.plan.n_high;
            for (int k = 0; k < 8; ++k) h[4 + k] = st.plan.high[k];
            h[12] = (int32_t)st.prog.size(); h[13] = (int32_t)st.groups.size(); h[14] = (int32_t)st.terms.size();
            h[15] = (int32_t)sizeof(TileInstr);
            h[16] = (int32_t)(sizeof(TileInstr) * st.prog.size() + sizeof(TileGroup) * st.groups.size() + sizeof(TileTerm) * st.terms.size());
            ok = ok && put(h, sizeof h);
            if (!st.prog.empty()) ok = ok && put(st.prog.data(), sizeof(TileInstr) * st.prog.size());
            if (!st.groups.empty()) ok = ok && put(st.groups.data(), sizeof(TileGroup) * st.groups.size());
            if (!st.terms.empty()) ok = ok && put(st.terms.data(), sizeof(TileTerm) * st.terms.size());
        } else if (st.type == 1) {
            h[1] = st.op.kind; h[2] = st.op.target; h[3] = st.op.t2; h[4] = st.op.const_hi;
            h[16] = (int32_t)(sizeof(uint64_t) + 7 * sizeof(double));
            ok = ok && put(h, sizeof h) && put(&st.op.cmask, sizeof(uint64_t)) && put(st.op.g.s, 7 * sizeof(double));
        } else if (st.type == 4) { // the basis state was re-placed: payload = the new permutation (64 x int32)
            h[16] = (int32_t)sizeof sink.relabel_perm;
            ok = ok && put(h, sizeof h) && put(sink.relabel_perm, sizeof sink.relabel_perm);
        } else {
            h[1] = st.type == 2 ? st.gbit : st.op.target;
            h[2] = st.lq;
            ok = ok && put(h, sizeof h);
        }
    }
    ok = ok && put(perm, sizeof perm);
    if (!ok) { set_error("output buffer too small"); return SPZ_ERR_INVALID_ARG; }
    *out_used = (int64_t)(w - static_cast<char *>(out));
    return SPZ_OK;
}

// ---- reductions --------------------------------------------------------------------------------------------
int spz_prob0(spz_state *st, int target, double *out) {
    SPZ_CHECK_STATE(st);
    if (!out) return SPZ_ERR_INVALID_ARG;
    return st->dist ? dist_reduce_scalar(st, 0, target, out) : reduce_scalar(st, 0, target, out);
}

int spz_norm2(spz_state *st, double *out) {
    SPZ_CHECK_STATE(st);
    if (!out) return SPZ_ERR_INVALID_ARG;
    return st->dist ? dist_reduce_scalar(st, 1, 0, out) : reduce_scalar(st, 1, 0, out);
}

int spz_measure_qubit(spz_state *st, int target, int reset, int forced_v, int *out_bit) {
    SPZ_CHECK_STATE(st);
    return measure_impl(st, target, reset, forced_v, out_bit);
}

int spz_qubit_expectation_value(spz_state *st, int target, double *out) {
    SPZ_CHECK_STATE(st);
    if (!out) return SPZ_ERR_INVALID_ARG;
    double p0 = 0.0;
    if (st->dist) SPZ_TRY(dist_reduce_scalar(st, 0, target, &p0));
    else SPZ_TRY(reduce_scalar(st, 0, target, &p0));
    *out = 2.0 * p0 - 1.0; // core.rs:217-218
    return SPZ_OK;
}

int spz_xyz_expectation_value(spz_state *st, char observable, const int32_t *targets, int n_targets, double *out) {
    SPZ_CHECK_STATE(st);
    int mode;
    switch (observable) { // panic!("observable {observable} not supported") core.rs:223-225
    case 'x': mode = 2; break;
    case 'y': mode = 3; break;
    case 'z': mode = 4; break;
    default: set_error("observable %c not supported", observable); return SPZ_ERR_INVALID_ARG;
    }
    if (n_targets < 0 || (n_targets && (!targets || !out))) return SPZ_ERR_INVALID_ARG;
    // 'z' on several targets: one read pass keeps the mass of every index bit (kernels_zall.cuh), instead of one pass per target
    if (mode == 4 && n_targets >= 2 && st->n >= 2 && st->n <= kZMaxBits) {
        if (st->dist) return dist_reduce_z_multi(st, targets, n_targets, out);
        for (int i = 0; i < n_targets; ++i)
            if (targets[i] < 0 || targets[i] >= st->n) { set_error("target %d out of range", targets[i]); return SPZ_ERR_INVALID_ARG; }
        double all[kZMaxBits + 1];
        SPZ_TRY(reduce_z_all(st, all));
        for (int i = 0; i < n_targets; ++i) out[i] = all[0] - 2.0 * all[1 + targets[i]];
        return SPZ_OK;
    }
    // 'x' / 'y' on three or more targets of a single-GPU register: the targets share tiles staged in shared memory, up to
    // twelve per read pass (kernels_xyall.cuh)
    if ((mode == 2 || mode == 3) && n_targets >= 3 && !st->dist && st->n >= 7 && st->n <= 40)
        return reduce_xy_multi(st, mode - 2, targets, n_targets, out);
    for (int i = 0; i < n_targets; ++i) {
        if (st->dist) SPZ_TRY(dist_reduce_scalar(st, mode, targets[i], &out[i]));
        else SPZ_TRY(reduce_scalar(st, mode, targets[i], &out[i]));
    }
    return SPZ_OK;
}

int spz_sample(spz_state *st, const double *u01, int64_t shots, int64_t *out_index) {
    SPZ_CHECK_STATE(st);
    if (shots < 0 || (shots && (!u01 || !out_index))) return SPZ_ERR_INVALID_ARG;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t layout");
    if (st->dist) return dist_sample(st, u01, shots, out_index); // shots owned by other ranks come back as -1
    return launch_sample(st, u01, shots, out_index);
}

// ---- instrumentation -----------------------------------------------------------------------------------------
int spz_timer_start(spz_state *st) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(join_pending(st));
    SPZ_CUDA(cudaEventRecord(st->ev0, st->stream));
    return SPZ_OK;
}

int spz_timer_stop(spz_state *st, double *out_ms) {
    SPZ_CHECK_STATE(st);
    SPZ_TRY(join_pending(st));
    SPZ_CUDA(cudaEventRecord(st->ev1, st->stream));
    SPZ_CUDA(cudaEventSynchronize(st->ev1));
    float ms = 0.f;
    SPZ_CUDA(cudaEventElapsedTime(&ms, st->ev0, st->ev1));
    if (out_ms) *out_ms = (double)ms;
    return SPZ_OK;
}

int64_t spz_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

} // extern "C"
