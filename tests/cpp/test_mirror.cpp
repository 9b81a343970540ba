// C++ host mirror (spinoza_b200/cpp/spinoza.hpp) exercised the way the reference's own tests exercise the Rust API.
// Expected numbers are the reference's golden vectors (file:line cited).  Exit 0 + "CPP_MIRROR_OK" on success,
// exit 3 when no CUDA device is visible (the engine has no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../spinoza_b200/cpp/spinoza.hpp"

using namespace spinoza;

static int fails = 0;
static void close_to(double a, double e, double eps, const char *what) { // utils.rs:163-165
    if (!(std::fabs(a - e) < eps)) { std::printf("FAIL %s: %.17g vs %.17g\n", what, a, e); ++fails; }
}

int main() {
    if (spz_device_count() == 0) {
        try { State s(3); } catch (const Error &e) { std::printf("no device: %s\n", e.what()); return e.status == SPZ_ERR_NO_DEVICE ? 3 : 4; }
        return 4;
    }
    { // h_gate_3_qubits gates.rs:1532-1544
        State s(3);
        for (std::size_t t = 0; t < 3; ++t) apply(Gate::H(), s, t);
        auto re = s.reals(); auto im = s.imags();
        for (int i = 0; i < 8; ++i) { close_to(re[i], 0.35355339059327384, 1e-10, "h re"); close_to(im[i], 0.0, 1e-10, "h im"); }
    }
    { // rx_gate_3_qubits gates.rs:1651-1682
        State s(3);
        for (std::size_t t = 0; t < 3; ++t) apply(Gate::RX(1.0), s, t);
        close_to(s.amp(0).first, 0.6758712218347053, 1e-10, "rx0");
        close_to(s.amp(1).second, -0.3692301313020644, 1e-10, "rx1");
        close_to(s.amp(7).second, 0.11019540730213864, 1e-10, "rx7");
    }
    { // qcbm_3_qubits gates.rs:1499-1529, 1787-1795
        const std::size_t n = 3;
        State s(n);
        for (std::size_t i = 0; i < n; ++i) { apply(Gate::RX(1.0), s, i); apply(Gate::RZ(1.0), s, i); }
        for (std::size_t i = 0; i + 1 < n; ++i) c_apply(Gate::X(), s, i, i + 1);
        for (int d = 0; d < 9; ++d) {
            for (std::size_t i = 0; i < n; ++i) { apply(Gate::RZ(1.0), s, i); apply(Gate::RX(1.0), s, i); apply(Gate::RZ(1.0), s, i); }
            for (std::size_t i = 0; i + 1 < n; ++i) c_apply(Gate::X(), s, i, i + 1);
        }
        for (std::size_t i = 0; i < n; ++i) { apply(Gate::RZ(1.0), s, i); apply(Gate::RX(1.0), s, i); }
        close_to(s.amp(0).first, 0.18037770683997864, 1e-10, "qcbm re0");
        close_to(s.amp(0).second, -0.17626993141958947, 1e-10, "qcbm im0");
        close_to(s.amp(7).first, 0.014503954556966365, 1e-10, "qcbm re7");
        close_to(s.amp(7).second, -0.11198008105074927, 1e-10, "qcbm im7");
    }
    { // append_value_encoding circuit.rs:1076-1113 through the circuit builder, fused
        const std::size_t n = 3; const double v = 4.0;
        QuantumRegister qr(n);
        QuantumCircuit qc({&qr});
        for (std::size_t t = 0; t < n; ++t) qc.h(t);
        for (std::size_t t = 0; t < n; ++t) qc.p(2.0 * PI / std::ldexp(1.0, (int)t + 1) * v, t);
        qc.iqft({2, 1, 0});
        qc.execute();
        auto re = qc.state.reals(); auto im = qc.state.imags();
        for (int i = 0; i < 8; ++i) { close_to(re[i], i == 4 ? 1.0 : 0.0, 1e-4, "venc re"); close_to(im[i], 0.0, 1e-4, "venc im"); }
    }
    { // controlled_u circuit.rs:1235-1250: circuit == functional, bit for bit
        QuantumRegister qr(3);
        QuantumCircuit qc({&qr});
        qc.fuse = false;
        qc.cu(1.0, 2.0, 3.0, 0, 1);
        qc.execute();
        State s(3);
        c_apply(Gate::U(1.0, 2.0, 3.0), s, 0, 1);
        if (qc.state.reals() != s.reals() || qc.state.imags() != s.imags()) { std::printf("FAIL controlled_u\n"); ++fails; }
    }
    { // xyz_exp_val core.rs:294-301 and measure twice circuit.rs:825-889
        State s(1);
        apply(Gate::RX(0.54), s, 0); apply(Gate::RY(0.12), s, 0);
        close_to(xyz_expectation_value('z', s, {0})[0], 0.8515405859048367, 1e-4, "expz");
        close_to(qubit_expectation_value(s, 0), 0.8515405859048367, 1e-4, "qev");
        QuantumRegister qr(4);
        QuantumCircuit qc({&qr});
        for (std::size_t t = 0; t < 4; ++t) { qc.h(t); qc.measure(t); }
        qc.execute();
        int bits[4];
        for (std::size_t t = 0; t < 4; ++t) bits[t] = qc.get_qubit_measured_val(t);
        for (std::size_t t = 0; t < 4; ++t) qc.measure(t);
        qc.execute();
        for (std::size_t t = 0; t < 4; ++t) if (qc.get_qubit_measured_val(t) != bits[t] || bits[t] < 0) { std::printf("FAIL remeasure\n"); ++fails; }
    }
    { // unsupported combination -> error, not abort (gates.rs:267)
        State s(2);
        bool threw = false;
        try { c_apply(Gate::SWAP(0, 1), s, 0, 1); } catch (const Error &e) { threw = e.status == SPZ_ERR_UNSUPPORTED; }
        if (!threw) { std::printf("FAIL unsupported\n"); ++fails; }
        bool inv = false;
        try { Gate::M().inverse(); } catch (const Error &) { inv = true; } // gates.rs:2050-2055
        if (!inv) { std::printf("FAIL m_inverse\n"); ++fails; }
    }
    if (fails) { std::printf("%d failures\n", fails); return 1; }
    std::printf("CPP_MIRROR_OK\n");
    return 0;
}
