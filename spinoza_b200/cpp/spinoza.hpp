// spinoza.hpp -- header-only C++17 mirror of Spinoza's public Rust API for the gate-application path, over the
// C ABI of libspinoza_b200.so (include/spinoza_b200.h).
//
// The reference is Rust and no Rust toolchain exists in this environment, so this is the compiled-language host
// side above the C ABI (the Rust shim in rust/spinoza-b200 is the same thing in the reference's own language).
// Names, argument order and semantics follow /root/reference/spinoza/src/{core,gates,circuit,measurement}.rs; each
// item cites its counterpart.  Where the reference panics (todo!/unimplemented!/assert!) this throws spinoza::Error.
//
//   spinoza::State s(3);                              // State::new(3)                      core.rs:32
//   spinoza::apply(spinoza::Gate::H(), s, 0);         // apply(Gate::H, &mut state, 0)      gates.rs:215
//   spinoza::c_apply(spinoza::Gate::P(0.5), s, 0, 1); // c_apply(Gate::P(0.5), &mut s, 0, 1) gates.rs:257
//   spinoza::QuantumRegister q(3); spinoza::QuantumCircuit qc({&q}); qc.h(0); qc.cx(0, 1); qc.execute();
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/spinoza_b200.h"

namespace spinoza {

using Float = double; // math.rs:11-15 (feature "double")
constexpr Float PI = 3.14159265358979323846;

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};
inline void check(int status) {
    if (status != SPZ_OK) throw Error(status, std::string(spz_status_string(status)) + ": " + spz_last_error());
}

// ---- Gate, gates.rs:44-92 -----------------------------------------------------------------------------------
struct Amplitude { // math.rs
    Float re, im;
};

struct Gate {
    spz_gate g{};
    static Gate make(int kind, Float a = 0, Float b = 0, Float c = 0, int t0 = 0, int t1 = 0) {
        Gate x; x.g.kind = kind; x.g.p[0] = a; x.g.p[1] = b; x.g.p[2] = c; x.g.t0 = t0; x.g.t1 = t1; return x;
    }
    static Gate H() { return make(SPZ_GATE_H); }
    static Gate M() { return make(SPZ_GATE_M); }
    static Gate X() { return make(SPZ_GATE_X); }
    static Gate Y() { return make(SPZ_GATE_Y); }
    static Gate Z() { return make(SPZ_GATE_Z); }
    static Gate P(Float theta) { return make(SPZ_GATE_P, theta); }
    static Gate RX(Float theta) { return make(SPZ_GATE_RX, theta); }
    static Gate RY(Float theta) { return make(SPZ_GATE_RY, theta); }
    static Gate RZ(Float theta) { return make(SPZ_GATE_RZ, theta); }
    static Gate SWAP(int t0, int t1) { return make(SPZ_GATE_SWAP, 0, 0, 0, t0, t1); }
    static Gate U(Float theta, Float phi, Float lambda) { return make(SPZ_GATE_U, theta, phi, lambda); }
    static Gate BitFlipNoise(Float prob) { return make(SPZ_GATE_BITFLIP, prob); }

    Gate inverse() const { // gates.rs:78-92
        switch (g.kind) {
        case SPZ_GATE_H: case SPZ_GATE_X: case SPZ_GATE_Y: case SPZ_GATE_Z: case SPZ_GATE_SWAP: return *this;
        case SPZ_GATE_P: case SPZ_GATE_RX: case SPZ_GATE_RY: case SPZ_GATE_RZ: return make(g.kind, -g.p[0]);
        case SPZ_GATE_U: return make(SPZ_GATE_U, -g.p[0], -g.p[2], -g.p[1]);
        default: throw Error(SPZ_ERR_UNSUPPORTED, "Gate::inverse: unimplemented!() (gates.rs:86)");
        }
    }
    // Gate::to_matrix, gates.rs:95-190: the 2 x 2 matrix, row-major
    std::array<Amplitude, 4> to_matrix() const {
        const Float r = 0.70710678118654752440;
        auto half = [&](Float &sn, Float &cs) { sn = std::sin(g.p[0] / 2); cs = std::cos(g.p[0] / 2); };
        Float sn = 0, cs = 0;
        switch (g.kind) {
        case SPZ_GATE_H: return {{{r, 0}, {r, 0}, {r, 0}, {-r, 0}}};
        case SPZ_GATE_X: return {{{0, 0}, {1, 0}, {1, 0}, {0, 0}}};
        case SPZ_GATE_Y: return {{{0, 0}, {0, -1}, {0, 1}, {0, 0}}};
        case SPZ_GATE_Z: return {{{1, 0}, {0, 0}, {0, 0}, {-1, 0}}};
        case SPZ_GATE_P: return {{{1, 0}, {0, 0}, {0, 0}, {std::cos(g.p[0]), std::sin(g.p[0])}}};
        case SPZ_GATE_RX: half(sn, cs); return {{{cs, 0}, {0, -sn}, {0, -sn}, {cs, 0}}};
        case SPZ_GATE_RY: half(sn, cs); return {{{cs, 0}, {-sn, 0}, {sn, 0}, {cs, 0}}};
        case SPZ_GATE_RZ: half(sn, cs); return {{{cs, -sn}, {0, 0}, {0, 0}, {cs, sn}}};
        case SPZ_GATE_U: {
            half(sn, cs);
            const Float phi = g.p[1], lam = g.p[2];
            return {{{cs, 0}, {-std::cos(lam) * sn, -std::sin(lam) * sn}, {std::cos(phi) * sn, std::sin(phi) * sn},
                     {std::cos(phi + lam) * cs, std::sin(phi + lam) * cs}}};
        }
        default: throw Error(SPZ_ERR_UNSUPPORTED, "Gate::to_matrix: unimplemented!() (gates.rs:188)");
        }
    }
};

// ---- State, core.rs:18-51 -----------------------------------------------------------------------------------
// Device-resident.  `reals()` / `imags()` download (the reference exposes the Vecs directly).
class State {
  public:
    explicit State(std::size_t n, int device = 0) { check(spz_create((int)n, device, &h_)); } // State::new
    State(const State &o) { check(spz_clone(o.h_, &h_)); }                                    // Clone
    State(State &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    State &operator=(State o) { std::swap(h_, o.h_); return *this; }
    ~State() { if (h_) spz_destroy(h_); }

    std::uint8_t n() const { return (std::uint8_t)spz_num_qubits(h_); }
    std::size_t len() const { return (std::size_t)spz_len(h_); } // core.rs:48
    std::vector<Float> reals() const { std::vector<Float> v(len()); check(spz_download(h_, v.data(), nullptr, 0, (int64_t)v.size())); return v; }
    std::vector<Float> imags() const { std::vector<Float> v(len()); check(spz_download(h_, nullptr, v.data(), 0, (int64_t)v.size())); return v; }
    std::pair<Float, Float> amp(std::size_t i) const { Float r, m; check(spz_download(h_, &r, &m, (int64_t)i, 1)); return {r, m}; }
    void set(const std::vector<Float> &re, const std::vector<Float> &im) { check(spz_upload(h_, re.data(), im.data(), 0, (int64_t)re.size())); }
    // whole-state upload that returns at once (page-locked buffers of len() doubles, untouched until the next sync / download):
    // the gates issued next follow the state piece by piece as it arrives
    void upload_async(const Float *re_pinned, const Float *im_pinned) { check(spz_upload_async(h_, re_pinned, im_pinned)); }
    void download_into(Float *re, Float *im) const { check(spz_download(h_, re, im, 0, (int64_t)len())); }
    void sync() { check(spz_sync(h_)); }
    void set_seed(std::uint64_t seed) { check(spz_set_seed(h_, seed)); }
    spz_state *handle() const { return h_; }
    static State adopt(spz_state *h) { State s; s.h_ = h; return s; } // a handle made elsewhere (spinoza::dist)

  private:
    State() = default;
    spz_state *h_ = nullptr;
};

// ---- gates.rs:215-320 -----------------------------------------------------------------------------------------
inline void apply(const Gate &gate, State &state, std::size_t target) { check(spz_apply(state.handle(), &gate.g, (int)target)); }
inline void c_apply(const Gate &gate, State &state, std::size_t control, std::size_t target) {
    check(spz_c_apply(state.handle(), &gate.g, (int)control, (int)target));
}
inline void cc_apply(const Gate &gate, State &state, std::size_t control0, std::size_t control1, std::size_t target) {
    check(spz_cc_apply(state.handle(), &gate.g, (int)control0, (int)control1, (int)target));
}
// mc_apply(gate, state, controls, zeros: Option<HashSet<usize>>, target): pass nullptr for None
inline void mc_apply(const Gate &gate, State &state, const std::vector<std::size_t> &controls,
                     const std::set<std::size_t> *zeros, std::size_t target) {
    std::vector<int32_t> c(controls.begin(), controls.end()), z;
    if (zeros) z.assign(zeros->begin(), zeros->end());
    check(spz_mc_apply(state.handle(), &gate.g, c.data(), (int)c.size(), zeros ? z.data() : nullptr, (int)z.size(), (int)target));
}
// extension (not in the reference): ones must be 1, zeros must be 0 -- true negative controls, one pass
inline void mc_apply_signed(const Gate &gate, State &state, const std::vector<std::size_t> &ones,
                            const std::vector<std::size_t> &zeros, std::size_t target) {
    uint64_t om = 0, zm = 0;
    for (std::size_t q : ones) om |= 1ull << q;
    for (std::size_t q : zeros) zm |= 1ull << q;
    check(spz_mc_apply_signed(state.handle(), &gate.g, om, zm, (int)target));
}
inline void iqft(State &state, const std::vector<std::size_t> &targets) { // core.rs:184
    std::vector<int32_t> t(targets.begin(), targets.end());
    check(spz_iqft(state.handle(), t.data(), (int)t.size()));
}

// ---- measurement.rs:12, core.rs:198-264 ----------------------------------------------------------------------------
// v: -1 mirrors `None`
inline std::uint8_t measure_qubit(State &state, std::size_t target, bool reset, int v = -1) {
    int bit = 0;
    check(spz_measure_qubit(state.handle(), (int)target, reset ? 1 : 0, v, &bit));
    return (std::uint8_t)bit;
}
inline Float qubit_expectation_value(const State &state, std::size_t target) {
    Float out = 0;
    check(spz_qubit_expectation_value(state.handle(), (int)target, &out));
    return out;
}
inline std::vector<Float> xyz_expectation_value(char observable, const State &state, const std::vector<std::size_t> &targets) {
    std::vector<int32_t> t(targets.begin(), targets.end());
    std::vector<Float> out(t.size());
    check(spz_xyz_expectation_value(state.handle(), observable, t.data(), (int)t.size(), out.data()));
    return out;
}

// ---- core.rs:65-129: Reservoir / reservoir_sampling ------------------------------------------------------------------
// Same construction and read-out as the reference; the filling is the engine's exact inverse-CDF sampler (spz_sample, one read
// pass over the device state) instead of num_tests rounds of weighted replacement, so every entry is an exact draw from
// |amplitude|^2 whatever num_tests is.
class Reservoir {
  public:
    explicit Reservoir(std::size_t k, std::uint64_t seed = 0x9E3779B97F4A7C15ull) : entries_(k, 0), seed_(seed) {}
    void sampling(const State &state, std::size_t /*num_tests*/) {
        std::vector<Float> u(entries_.size());
        for (auto &x : u) { // splitmix64, 53 bits per draw
            seed_ += 0x9E3779B97F4A7C15ull;
            std::uint64_t z = seed_;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            x = (Float)((z ^ (z >> 31)) >> 11) * (1.0 / 9007199254740992.0);
        }
        std::vector<int64_t> idx(u.size());
        check(spz_sample(state.handle(), u.data(), (int64_t)u.size(), idx.data()));
        for (std::size_t i = 0; i < idx.size(); ++i) entries_[i] = (std::size_t)idx[i];
    }
    std::map<std::size_t, std::size_t> get_outcome_count() const { // core.rs:115-121
        std::map<std::size_t, std::size_t> samples;
        for (auto e : entries_) ++samples[e];
        return samples;
    }
    const std::vector<std::size_t> &entries() const { return entries_; }

  private:
    std::vector<std::size_t> entries_;
    std::uint64_t seed_;
};
inline Reservoir reservoir_sampling(const State &state, std::size_t reservoir_size, std::size_t num_tests) { // core.rs:125
    Reservoir reservoir(reservoir_size);
    reservoir.sampling(state, num_tests);
    return reservoir;
}

// ---- sharded registers (no counterpart in the reference; include/spinoza_b200.h "multi-GPU") ---------------------------------
// A shard IS a State: apply / QuantumCircuit::execute / measure_qubit / the reductions are the same calls, made by every rank in
// the same order.
namespace dist {
class Rendezvous { // the host-side control plane between the processes of a node (csrc/rendezvous.cu)
  public:
    Rendezvous(int rank, int world, const char *dir = nullptr) : rank(rank), world(world) { check(spz_rdv_open(dir, rank, world, &h_)); }
    Rendezvous(const Rendezvous &) = delete;
    ~Rendezvous() { if (h_) spz_rdv_close(h_); }
    void barrier() { check(spz_rdv_barrier(h_)); }
    Float max_float(Float x) {
        std::vector<Float> all((std::size_t)world);
        check(spz_rdv_allgather(h_, &x, sizeof x, all.data()));
        Float m = all[0];
        for (Float v : all) m = v > m ? v : m;
        return m;
    }
    spz_rdv *handle() const { return h_; }
    const int rank, world;

  private:
    spz_rdv *h_ = nullptr;
};
// this rank's shard of an n-qubit register, connected to the other processes' shards over CUDA IPC
inline State sharded_state(std::size_t n, Rendezvous &rdv, int device) {
    spz_state *h = nullptr;
    check(spz_dist_create((int)n, rdv.rank, rdv.world, device, &h));
    State s = State::adopt(h);
    check(spz_dist_connect_rdv(h, rdv.handle()));
    return s;
}
// all `world` shards in one process (plain device pointers, same kernels)
inline std::vector<State> local_group(std::size_t n, int world, const std::vector<int> &devices = {0}) {
    std::vector<spz_state *> hs((std::size_t)world, nullptr);
    std::vector<State> out;
    for (int r = 0; r < world; ++r) {
        check(spz_dist_create((int)n, r, world, devices[(std::size_t)r % devices.size()], &hs[(std::size_t)r]));
        out.push_back(State::adopt(hs[(std::size_t)r]));
    }
    check(spz_dist_connect_local(hs.data(), world));
    return out;
}
inline std::vector<std::size_t> perm(const State &s, std::size_t n_total) { // logical qubit -> physical index bit
    std::vector<int32_t> p(n_total);
    check(spz_dist_perm(s.handle(), p.data()));
    return std::vector<std::size_t>(p.begin(), p.end());
}
inline std::size_t local_qubits(const State &s) { return (std::size_t)spz_dist_local_qubits(s.handle()); }
inline Float norm2(const State &s) { Float v = 0; check(spz_norm2(s.handle(), &v)); return v; } // a collective
struct Stats { double exchanges, bytes_sent, exchange_ms, overlapped; };
inline Stats stats(const State &s) { double o[4]; check(spz_dist_stats(s.handle(), o)); return {o[0], o[1], o[2], o[3]}; }
} // namespace dist

// ---- circuit.rs ---------------------------------------------------------------------------------------------------
struct QuantumRegister { // circuit.rs:13-51
    std::vector<std::size_t> q;
    explicit QuantumRegister(std::size_t size) { if (!size) throw Error(SPZ_ERR_INVALID_ARG, "assert!(size > 0)"); for (std::size_t i = 0; i < size; ++i) q.push_back(i); }
    std::size_t operator[](std::size_t i) const { return q[i]; }
    std::size_t len() const { return q.size(); }
    void update_shift(std::size_t shift) { for (auto &x : q) x += shift; }
    std::size_t get_shift() const { return q[0]; }
};

struct Controls { // circuit.rs:55-109
    int kind = SPZ_CTRL_NONE;
    std::vector<std::size_t> controls;
    std::set<std::size_t> zeros;
    static Controls None() { return {}; }
    static Controls Single(std::size_t c) { Controls x; x.kind = SPZ_CTRL_SINGLE; x.controls = {c}; return x; }
    static Controls Ones(std::vector<std::size_t> cs) { Controls x; x.kind = SPZ_CTRL_ONES; x.controls = std::move(cs); return x; }
    static Controls Mixed(std::vector<std::size_t> cs, std::set<std::size_t> zs) { Controls x; x.kind = SPZ_CTRL_MIXED; x.controls = std::move(cs); x.zeros = std::move(zs); return x; }
    // extension: zs (a subset of cs) are true negative controls
    static Controls Signed(std::vector<std::size_t> cs, std::set<std::size_t> zs) { Controls x = Mixed(std::move(cs), std::move(zs)); x.kind = SPZ_CTRL_SIGNED; return x; }
    static Controls from(const std::vector<std::size_t> &cs, const std::set<std::size_t> *zs) { // circuit.rs:73-86
        if (zs) return Mixed(cs, *zs);
        if (cs.empty()) return None();
        if (cs.size() == 1) return Single(cs[0]);
        return Ones(cs);
    }
    Controls new_with_control(std::size_t control, std::size_t shift) const { // circuit.rs:97-108
        std::vector<std::size_t> cs;
        for (auto c : controls) cs.push_back(c + shift);
        cs.push_back(control);
        if (zeros.empty()) return from(cs, nullptr);
        std::set<std::size_t> zs;
        for (auto z : zeros) zs.insert(z + shift);
        return from(cs, &zs);
    }
};

struct QuantumTransformation { // circuit.rs:113-120
    Gate gate;
    std::size_t target;
    Controls controls;
};

class QuantumCircuit { // circuit.rs:168-601
  public:
    std::vector<QuantumTransformation> transformations;
    State state;
    std::vector<std::size_t> quantum_registers_info;
    bool fuse = true, exact = false;

    explicit QuantumCircuit(std::vector<QuantumRegister *> registers, int device = 0) : state(shift_all(registers), device) {
        for (auto *r : registers) quantum_registers_info.push_back(r->len());
    }
    explicit QuantumCircuit(State s) : state(std::move(s)) {}

    const State &get_statevector() const { return state; }
    void inverse() { // circuit.rs:206-211
        std::vector<QuantumTransformation> r(transformations.rbegin(), transformations.rend());
        for (auto &t : r) t.gate = t.gate.inverse();
        transformations = std::move(r);
    }
    void add(QuantumTransformation t) { transformations.push_back(std::move(t)); }
    void measure(std::size_t t) { add({Gate::M(), t, Controls::None()}); }
    void swap(std::size_t t0, std::size_t t1) { add({Gate::SWAP((int)t0, (int)t1), 0, Controls::None()}); }
    void x(std::size_t t) { add({Gate::X(), t, Controls::None()}); }
    void y(std::size_t t) { add({Gate::Y(), t, Controls::None()}); }
    void z(std::size_t t) { add({Gate::Z(), t, Controls::None()}); }
    void h(std::size_t t) { add({Gate::H(), t, Controls::None()}); }
    void p(Float a, std::size_t t) { add({Gate::P(a), t, Controls::None()}); }
    void rx(Float a, std::size_t t) { add({Gate::RX(a), t, Controls::None()}); }
    void ry(Float a, std::size_t t) { add({Gate::RY(a), t, Controls::None()}); }
    void rz(Float a, std::size_t t) { add({Gate::RZ(a), t, Controls::None()}); }
    void u(Float th, Float ph, Float la, std::size_t t) { add({Gate::U(th, ph, la), t, Controls::None()}); }
    void cx(std::size_t c, std::size_t t) { add({Gate::X(), t, Controls::Single(c)}); }
    void ccx(std::size_t c1, std::size_t c2, std::size_t t) { add({Gate::X(), t, Controls::Ones({c1, c2})}); }
    void ch(std::size_t c, std::size_t t) { add({Gate::H(), t, Controls::Single(c)}); }
    void cy(std::size_t c, std::size_t t) { add({Gate::Y(), t, Controls::Single(c)}); }
    void cp(Float a, std::size_t c, std::size_t t) { add({Gate::P(a), t, Controls::Single(c)}); }
    void crx(Float a, std::size_t c, std::size_t t) { add({Gate::RX(a), t, Controls::Single(c)}); }
    void cry(Float a, std::size_t c, std::size_t t) { add({Gate::RY(a), t, Controls::Single(c)}); }
    void crz(Float a, std::size_t c, std::size_t t) { add({Gate::RZ(a), t, Controls::Single(c)}); }
    void cu(Float th, Float ph, Float la, std::size_t c, std::size_t t) { add({Gate::U(th, ph, la), t, Controls::Single(c)}); }
    void bit_flip_noise(Float prob, std::size_t t) { add({Gate::BitFlipNoise(prob), t, Controls::None()}); }
    void iqft(const std::vector<std::size_t> &targets) { // circuit.rs:438-445
        for (std::size_t j = targets.size(); j-- > 0;) {
            h(targets[j]);
            for (std::size_t k = j; k-- > 0;) cp(-PI / std::ldexp(1.0, (int)(j - k)), targets[j], targets[k]);
        }
    }
    void append(const QuantumCircuit &c, const QuantumRegister &reg) { // circuit.rs:448-460
        for (const auto &t : c.transformations) add({t.gate, reg.get_shift() + t.target, t.controls});
    }
    void c_append(const QuantumCircuit &c, std::size_t ctl, const QuantumRegister &reg) { // circuit.rs:463-476
        if (ctl >= reg.get_shift() && ctl < reg.get_shift() + reg.len()) throw Error(SPZ_ERR_INVALID_ARG, "control inside the register");
        for (const auto &t : c.transformations) add({t.gate, reg.get_shift() + t.target, t.controls.new_with_control(ctl, reg.get_shift())});
    }
    void mc_append(const QuantumCircuit &c, const std::vector<std::size_t> &ctls, const QuantumRegister &reg) { // circuit.rs:479-511
        for (auto ctl : ctls) {
            if (ctl >= reg.get_shift() && ctl < reg.get_shift() + reg.len()) throw Error(SPZ_ERR_INVALID_ARG, "control inside the register");
            for (const auto &t : c.transformations) add({t.gate, reg.get_shift() + t.target, t.controls.new_with_control(ctl, reg.get_shift())});
        }
    }
    bool is_qubit_measured(std::size_t q) const { return (measured_ >> q) & 1; }
    int get_qubit_measured_val(std::size_t q) const { return is_qubit_measured(q) ? (int)((vals_ >> q) & 1) : -1; }

    void execute() { // circuit.rs:552-600; drains the list
        std::vector<spz_op> ops(transformations.size());
        for (std::size_t i = 0; i < ops.size(); ++i) {
            const auto &t = transformations[i];
            spz_op &o = ops[i];
            o = spz_op{};
            o.kind = t.gate.g.kind; o.target = (int32_t)t.target; o.t0 = t.gate.g.t0; o.t1 = t.gate.g.t1;
            for (int j = 0; j < 3; ++j) o.p[j] = t.gate.g.p[j];
            o.ctrl_kind = t.controls.kind;
            for (auto c : t.controls.controls) o.ctrl_mask |= 1ull << c;
            for (auto z : t.controls.zeros) o.zeros_mask |= 1ull << z;
        }
        transformations.clear();
        const uint32_t flags = fuse ? (SPZ_EXEC_FUSE | (exact ? SPZ_EXEC_EXACT : 0u)) : SPZ_EXEC_NO_FUSE;
        check(spz_execute(state.handle(), ops.data(), (int64_t)ops.size(), flags, &measured_, &vals_));
    }

  private:
    std::uint64_t measured_ = 0, vals_ = 0; // QubitTracker circuit.rs:122-164
    static std::size_t shift_all(std::vector<QuantumRegister *> &regs) { // circuit.rs:181-190
        std::size_t bits = 0;
        for (auto *r : regs) { r->update_shift(bits); bits += r->len(); }
        return bits;
    }
};

} // namespace spinoza
