"""The reference's own unit tests, scenario by scenario, through this engine's mirror of the reference API on the GPU.

Each test below restates (does not copy) one `#[test]` of /root/reference/spinoza/src/{circuit,core,gates}.rs with the same
construction and the same assertion -- and, where the reference only checks that nothing panics, the stronger one available
here: equality with the CPU oracle.  Unfused execution is compared bit for bit, fused (the mirror's default) within 1e-12.
Scenarios already covered elsewhere are listed at the bottom with the test that holds them.
"""
import math

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import Gate, QuantumCircuit, QuantumRegister, QuantumTransformation, Controls
from tests.test_gpu_parity import oracle_ops_from, to_gpu

pytestmark = pytest.mark.gpu
PI = math.pi
MODES = [dict(fuse=False), dict(fuse=True), dict(fuse=True, exact=True)]


def run_both(build, n, init=None, modes=MODES):
    """build(qc) fills a circuit; it runs on the GPU in every mode and on the oracle.  Returns the oracle state."""
    cpu = init.clone() if init is not None else orc.State(n)
    first = True
    for kw in modes:
        qc = QuantumCircuit(QuantumRegister(n), **kw)
        if init is not None:
            qc.state = to_gpu(init)
        build(qc)
        if first:
            orc.execute(cpu, oracle_ops_from(qc))
            first = False
        qc.execute()
        re, im = qc.state.download()
        if kw.get("exact") or not kw["fuse"]:
            assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags), kw
        else:
            assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12, kw
    return cpu


def test_circuit_value_encoding_runs():  # circuit.rs:623-641 (v = 2.4, n = 3: only "does not panic" in the reference)
    def build(qc):
        for i in range(3):
            qc.h(i)
        for i in range(3):
            qc.p(2.0 * PI / (2.0 ** (i + 1)) * 2.4, i)
        qc.iqft([2, 1, 0])
    run_both(build, 3)


def test_circuit_z_gate():  # circuit.rs:644-656
    def build(qc):
        qc.h(0); qc.h(1); qc.z(0)
    cpu = run_both(build, 2)
    assert np.allclose(cpu.reals, [0.5, -0.5, 0.5, -0.5], atol=1e-15)


def test_circuit_crx_and_cy_on_even_pairs():  # circuit.rs:659-675, 722-738
    for gate in ("crx", "cy"):
        def build(qc):
            for i in range(3):
                qc.h(i)
            for t in range(0, 2, 2):
                qc.crx(3.043, t, t + 1) if gate == "crx" else qc.cy(t, t + 1)
        run_both(build, 3)


def test_circuit_ch_equals_functional_c_apply():  # circuit.rs:678-697 (assert_eq on reals and imags)
    n = 3
    s = sb.State(n)
    for t in range(n):
        sb.apply(Gate.H, s, t)
    sb.c_apply(Gate.H, s, 0, 1)
    qc = QuantumCircuit(QuantumRegister(n), fuse=False)
    for t in range(n):
        qc.h(t)
    qc.ch(0, 1)
    qc.execute()
    assert np.array_equal(qc.state.reals, s.reals) and np.array_equal(qc.state.imags, s.imags)
    # gates.rs:2100-2133: the values themselves
    for i in range(0, 8, 4):
        assert np.allclose(s.reals[i:i + 4], [0.353553391, 0.5, 0.353553391, 0.0], atol=1e-4)
        assert np.allclose(s.imags[i:i + 4], 0.0, atol=1e-4)


def test_circuit_crz_equals_functional_c_apply():  # circuit.rs:700-719, gates.rs:2136-2171
    n = 3
    s = sb.State(n)
    for t in range(n):
        sb.apply(Gate.H, s, t)
    sb.c_apply(Gate.RZ(PI / 2.0), s, 0, 1)
    qc = QuantumCircuit(QuantumRegister(n), fuse=False)
    for t in range(n):
        qc.h(t)
    qc.crz(PI / 2.0, 0, 1)
    qc.execute()
    assert np.array_equal(qc.state.reals, s.reals) and np.array_equal(qc.state.imags, s.imags)
    for i in range(0, 8, 4):
        assert np.allclose(s.reals[i:i + 4], [0.353553391, 0.25, 0.353553391, 0.25], atol=1e-4)
        assert np.allclose(s.imags[i:i + 4], [0.0, -0.25, 0.0, 0.25], atol=1e-4)


def test_circuit_ccx_and_get_statevector():  # circuit.rs:741-754
    def build(qc):
        qc.h(0); qc.h(1); qc.ccx(0, 1, 2)
    cpu = run_both(build, 3)
    want = np.zeros(8); want[[0, 1, 2, 7]] = 0.5
    assert np.allclose(cpu.reals, want, atol=1e-15)
    qc = QuantumCircuit(QuantumRegister(3))
    build(qc)
    qc.execute()
    assert np.allclose(qc.get_statevector().amps(), want, atol=1e-12)


def test_circuit_x_gate_test_is_an_rx():  # circuit.rs:757-768 (named x_gate, applies RX(pi/2))
    def build(qc):
        qc.h(0); qc.h(1); qc.rx(PI / 2.0, 0)
    run_both(build, 2)


def test_circuit_swap_all_qubits_equals_three_cx():  # circuit.rs:892-926 (assert_eq)
    n = 9
    init = orc.gen_random_state(n, 9)
    assert abs(orc.norm2(init) - 1.0) < 1e-3
    cpu = init.clone()
    for i in range(n >> 1):  # utils.rs:204-208: swap = CX(a,b) CX(b,a) CX(a,b)
        a, b = i, n - 1 - i
        for c, t in ((a, b), (b, a), (a, b)):
            orc.c_apply(orc.X, cpu, c, t)
    for kw in MODES:
        qc = QuantumCircuit(QuantumRegister(n), **kw)
        qc.state = to_gpu(init)
        for i in range(n >> 1):
            qc.swap(i, n - 1 - i)
        qc.execute()
        assert np.array_equal(qc.state.reals, cpu.reals) and np.array_equal(qc.state.imags, cpu.imags), kw


def test_circuit_inverse_iqft_round_trip():  # circuit.rs:929-960 (1e-3 in the reference; 1e-12 here)
    for n in (2, 13):
        init = orc.gen_random_state(n, 31 + n)
        for kw in MODES:
            qc = QuantumCircuit(QuantumRegister(n), **kw)
            qc.state = to_gpu(init)
            targets = list(reversed(range(n)))
            qc.iqft(targets)
            qc.execute()
            qc.iqft(targets)
            qc.inverse()
            qc.execute()
            assert np.max(np.abs(qc.state.amps() - init.amps())) <= 1e-12, (n, kw)


def test_circuit_inverse_equals_hand_written_inverse():  # circuit.rs:963-989 (assert_eq)
    for kw in MODES:
        qc1 = QuantumCircuit(QuantumRegister(2), **kw)
        qc1.h(0); qc1.p(PI / 4.0, 1); qc1.inverse(); qc1.execute()
        qc2 = QuantumCircuit(QuantumRegister(2), **kw)
        qc2.p(-(PI / 4.0), 1); qc2.h(0); qc2.execute()
        assert np.array_equal(qc1.state.reals, qc2.state.reals) and np.array_equal(qc1.state.imags, qc2.state.imags)


def gate_to_circuit(gate, n, target):  # circuit.rs:1004-1013
    qc = QuantumCircuit(QuantumRegister(n))
    qc.add(QuantumTransformation(gate, target, Controls.none()))
    return qc


def iqft_circuit_from_controlled_append(n, multi_control):  # circuit.rs:1028-1053
    qr = QuantumRegister(n)
    out = QuantumCircuit(qr)
    targets = list(reversed(range(n)))
    for j in reversed(range(n)):
        out.append(gate_to_circuit(Gate.H, n, targets[j]), qr)
        for k in reversed(range(j)):
            one = QuantumRegister(1)
            one.update_shift(targets[k])
            g = gate_to_circuit(Gate.P(-PI / (2.0 ** (j - k))), 1, 0)
            if multi_control:
                out.mc_append(g, [targets[j]], one)
            else:
                out.c_append(g, targets[j], one)
    return out


def test_circuit_append_two_registers():  # circuit.rs:1056-1073
    for kw in MODES:
        qr0, qr1 = QuantumRegister(1), QuantumRegister(1)
        qc = QuantumCircuit(qr0, qr1, **kw)
        qc.append(gate_to_circuit(Gate.H, 1, 0), qr0)
        qc.append(gate_to_circuit(Gate.H, 1, 0), qr1)
        qc.execute()
        assert np.allclose(qc.state.reals, 0.5, atol=1e-4) and np.allclose(qc.state.imags, 0.0, atol=1e-4)


@pytest.mark.parametrize("how", ["append", "c_append", "mc_append"])
def test_circuit_value_encoding_through_appended_iqft(how):  # circuit.rs:1076-1192: |v> = |4> at 1e-4
    n, v = 3, 4.0
    for kw in MODES:
        qr = QuantumRegister(n)
        qc = QuantumCircuit(qr, **kw)
        for t in range(n):
            qc.append(gate_to_circuit(Gate.H, n, t), qr)
        for t in range(n):
            qc.append(gate_to_circuit(Gate.P(2.0 * PI / (2.0 ** (t + 1)) * v), n, t), qr)
        if how == "append":
            sub = QuantumCircuit(QuantumRegister(n))
            sub.iqft(list(reversed(range(n))))
        else:
            sub = iqft_circuit_from_controlled_append(n, how == "mc_append")
        qc.append(sub, qr)
        qc.execute()
        want = np.zeros(1 << n); want[int(v)] = 1.0
        assert np.max(np.abs(qc.state.reals - want)) < 1e-4 and np.max(np.abs(qc.state.imags)) < 1e-4, (how, kw)


def test_circuit_bit_flip_noise_extremes():  # circuit.rs:1195-1232, gates.rs:2174-2195
    init = orc.gen_random_state(1, 5)
    for kw in MODES:
        qc1 = QuantumCircuit(QuantumRegister(1), **kw); qc1.state = to_gpu(init)
        qc2 = QuantumCircuit(QuantumRegister(1), **kw); qc2.state = to_gpu(init)
        qc2.bit_flip_noise(0.0, 0)
        qc2.execute()                                  # probability 0: nothing changes
        assert np.array_equal(qc2.state.reals, init.reals) and np.array_equal(qc2.state.imags, init.imags)
        qc1.x(0); qc2.bit_flip_noise(1.0, 0)           # probability 1: the X gate
        qc1.execute(); qc2.execute()
        assert np.array_equal(qc1.state.reals, qc2.state.reals) and np.array_equal(qc1.state.imags, qc2.state.imags)
    s = to_gpu(init)
    sb.apply(Gate.BitFlipNoise(0.0), s, 0)
    assert np.array_equal(s.reals, init.reals)
    sb.apply(Gate.BitFlipNoise(1.0), s, 0)
    assert s.reals[0] == init.reals[1] and s.reals[1] == init.reals[0] and s.imags[0] == init.imags[1]


def test_circuit_controlled_u_equals_functional():  # circuit.rs:1235-1251 (assert_eq), gates.rs:2198-2230 (values at 1e-5)
    qc = QuantumCircuit(QuantumRegister(3), fuse=False)
    qc.cu(1.0, 2.0, 3.0, 0, 1)
    qc.execute()
    s = sb.State(3)
    sb.c_apply(Gate.U(1.0, 2.0, 3.0), s, 0, 1)
    assert np.array_equal(qc.state.reals, s.reals) and np.array_equal(qc.state.imags, s.imags)
    s = sb.State(3)
    for t in range(3):
        sb.apply(Gate.H, s, t)
    sb.c_apply(Gate.U(1.0, 2.0, 3.0), s, 0, 1)
    assert np.allclose(s.reals, [0.35355339, 0.47807852, 0.35355339, 0.01747458] * 2, atol=1e-5)
    assert np.allclose(s.imags, [0.0, -0.0239202, 0.0, -0.14339942] * 2, atol=1e-5)


def test_core_encoded_integers_through_reservoir_sampling():  # core.rs:272-291
    n = 3
    state = sb.State(n)
    hist = sb.reservoir_sampling(state, len(state), len(state) * 10_000).get_outcome_count()
    assert hist == {0: len(state)}
    for i in range(1, 1 << n):
        state.set_basis(i)
        hist = sb.reservoir_sampling(state, len(state), len(state) * 10_000).get_outcome_count()
        assert hist == {i: len(state)}


def test_core_expectation_values():  # core.rs:294-337
    state = sb.State(1)
    sb.apply(Gate.RX(0.54), state, 0)
    sb.apply(Gate.RY(0.12), state, 0)
    z = sb.xyz_expectation_value("z", state, [0])[0]
    assert abs(z - 0.8515405859048367) < 1e-4                                  # :299-300
    with pytest.raises(sb.SpinozaError):
        sb.xyz_expectation_value("a", state, [0])                              # :303-310 should_panic
    cpu = orc.State(1)
    orc.apply(orc.RX, cpu, 0, (0.54,)); orc.apply(orc.RY, cpu, 0, (0.12,))
    for axis in "xy":                                                          # :313-326 (no assertion in the reference)
        got = sb.xyz_expectation_value(axis, state, [0])[0]
        assert abs(got - orc.xyz_expectation_value(axis, cpu, [0])[0]) < 1e-12
    assert abs(sb.qubit_expectation_value(state, 0) - z) < 1e-4                # :329-337


def test_gates_x_and_y_on_20_qubits():  # gates.rs:1563-1576, 1595-1608: |0..0> -> |1..1> (times i^20 = 1 for Y)
    n = 20
    for gate in (Gate.X, Gate.Y):
        s = sb.State(n)
        for t in range(n):
            sb.apply(gate, s, t)
        re, im = s.download()
        assert re[-1] == 1.0 and np.count_nonzero(re) == 1 and np.count_nonzero(im) == 0


def test_gates_value_encoding_20_qubits():  # gates.rs:1770-1784: H, P(2 pi v / 2^(i+1)), iqft -> |v>, v = 2.4 is not an integer
    n, v = 20, 2.4
    for mode in (None, "0"):
        import os
        if mode is None:
            os.environ.pop("SPZ_IQFT_FUSE", None)
        else:
            os.environ["SPZ_IQFT_FUSE"] = mode
        try:
            s, cpu = sb.State(n), orc.State(n)
            for i in range(n):
                sb.apply(Gate.H, s, i); orc.apply(orc.H, cpu, i)
            for i in range(n):
                ang = 2.0 * PI / (2.0 ** (i + 1)) * v
                sb.apply(Gate.P(ang), s, i); orc.apply(orc.P, cpu, i, (ang,))
            targets = list(reversed(range(n)))
            sb.iqft(s, targets); orc.iqft(cpu, targets)
            re, im = s.download()
            assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12
            assert abs(sb.norm2(s) - 1.0) < 1e-10
        finally:
            os.environ.pop("SPZ_IQFT_FUSE", None)


# Held elsewhere: register_shift -> test_host_logic.test_register_shift; all_gates_as_transformations (n = 17) ->
# test_gpu_parity.test_all_gates_as_transformations_n17; measure (n = 21, twice) -> test_measure_all_twice_gives_identical_bits;
# h/x/y/z/p/rx/ry/rz/u 3-qubit known answers, QCBM-3 / QCBM-20 -> test_golden_vectors_through_the_gpu, test_qcbm_20_qubits_golden
# and tests/test_oracle_golden.py; swap_9_qubits / swap_all_qubits (gates.rs) -> test_swap_matches_oracle_and_three_cx;
# *_inverse -> test_host_logic.test_gate_inverse_matrices; measurement.rs -> test_measure_qubit_known_state / _forced_outcomes;
# openqasm.rs -> test_host_logic.test_qasm_*; unitaries.rs, math.rs, utils.rs colour maps, config.rs: off the path.
