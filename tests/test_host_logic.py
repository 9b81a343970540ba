"""Host-side mirror of circuit.rs / gates.rs / openqasm.rs that needs no GPU: IR building, inverse,
append/c_append/mc_append, QFT definition, Gate matrices, QASM import."""
import math
from pathlib import Path

import numpy as np
import pytest

import spinoza_b200 as sb
from spinoza_b200 import Controls, Gate, QuantumCircuit, QuantumRegister

PI = math.pi
QASM = Path(__file__).resolve().parent / "golden"


def test_register_shift():  # circuit.rs:613-620
    qr = QuantumRegister(4)
    assert qr.get_shift() == 0
    qr.update_shift(4)
    assert qr.get_shift() == 4


def test_circuit_new_shifts_registers():  # circuit.rs:181-197
    a, b = QuantumRegister(2), QuantumRegister(3)
    qc = QuantumCircuit(a, b)
    assert (a.get_shift(), b.get_shift()) == (0, 2)
    assert qc.quantum_registers_info == [2, 3] and qc.n_qubits == 5


@pytest.mark.parametrize("g", [Gate.H, Gate.X, Gate.Y, Gate.Z, Gate.P(2.03), Gate.RX(2.03), Gate.RZ(3.03),
                               Gate.RY(3.03), Gate.U(1.0, 2.0, 3.0)])
def test_gate_inverse_matrices(g):  # gates.rs:1904-2045
    ident = g.to_matrix() @ g.inverse().to_matrix()
    assert np.allclose(ident, np.eye(2), atol=1e-3)
    assert np.allclose(ident, np.eye(2), atol=1e-14)


def test_inverse_of_m_and_swap_matrix_panic():  # gates.rs:2050-2062
    with pytest.raises(sb.SpinozaError):
        Gate.M.inverse()
    with pytest.raises(sb.SpinozaError):
        Gate.SWAP(0, 1).inverse().to_matrix()


def test_iqft_gate_list():  # circuit.rs:438-445
    qc = QuantumCircuit(QuantumRegister(3))
    qc.iqft([2, 1, 0])
    got = [(t.gate, t.target, t.controls.controls) for t in qc.transformations]
    want = [(Gate.H, 0, []), (Gate.P(-PI / 2), 1, [0]), (Gate.P(-PI / 4), 2, [0]), (Gate.H, 1, []),
            (Gate.P(-PI / 2), 2, [1]), (Gate.H, 2, [])]
    assert got == want


def test_qft_definition():  # starts H(n-1), CP(pi/2, c=n-2, t=n-1), H(n-2), ...
    n = 5
    qc = QuantumCircuit(QuantumRegister(n))
    qc.qft()
    tr = qc.transformations
    assert len(tr) == n + n * (n - 1) // 2
    assert (tr[0].gate, tr[0].target) == (Gate.H, n - 1)
    assert (tr[1].gate, tr[1].target, tr[1].controls.controls) == (Gate.P(PI / 2), n - 1, [n - 2])
    assert (tr[2].gate, tr[2].target) == (Gate.H, n - 2)
    assert (tr[-1].gate, tr[-1].target) == (Gate.H, 0)


def test_inverse_reverses_and_inverts():  # circuit.rs:963-989
    qc = QuantumCircuit(QuantumRegister(2))
    qc.h(0)
    qc.p(PI / 4, 1)
    qc.inverse()
    assert [(t.gate, t.target) for t in qc.transformations] == [(Gate.P(-PI / 4), 1), (Gate.H, 0)]


def test_controls_new_with_control():  # circuit.rs:97-108
    assert Controls.none().new_with_control(3, 0).kind == Controls.SINGLE
    c = Controls.single(1).new_with_control(4, 2)
    assert c.kind == Controls.ONES and c.controls == [3, 4]
    m = Controls.mixed([0, 1], {1}).new_with_control(5, 1)
    assert m.kind == Controls.MIXED and m.controls == [1, 2, 5] and m.zeros == {2}


def test_append_shifts_target_not_controls():  # circuit.rs:448-460
    inner = QuantumCircuit(QuantumRegister(2))
    inner.cx(0, 1)
    a, b = QuantumRegister(2), QuantumRegister(2)
    qc = QuantumCircuit(a, b)
    qc.append(inner, b)
    t = qc.transformations[0]
    assert t.target == 3 and t.controls.controls == [0]
    with pytest.raises(AssertionError):
        qc.append(inner, QuantumRegister(3))


def test_c_append_and_mc_append():  # circuit.rs:463-511
    one = QuantumCircuit(QuantumRegister(1))
    one.p(0.5, 0)
    r0, r1 = QuantumRegister(2), QuantumRegister(1)
    qc = QuantumCircuit(r0, r1)
    qc.c_append(one, 0, r1)
    t = qc.transformations[-1]
    assert t.target == 2 and t.controls.kind == Controls.SINGLE and t.controls.controls == [0]
    qc.mc_append(one, [0, 1], r1)
    assert [x.controls.controls for x in qc.transformations[-2:]] == [[0], [1]]
    with pytest.raises(AssertionError):
        qc.c_append(one, 2, r1)
    with pytest.raises(ValueError):
        qc.mc_append(one, [2], r1)


def test_encode_ops_layout():
    qc = QuantumCircuit(QuantumRegister(4))
    qc.ccx(0, 1, 2)
    qc.swap(1, 3)
    qc.cu(0.1, 0.2, 0.3, 3, 0)
    arr, n = qc._encode()
    assert n == 3
    assert (arr[0].kind, arr[0].ctrl_kind, arr[0].ctrl_mask, arr[0].target) == (Gate.KIND_X, Controls.ONES, 0b11, 2)
    assert (arr[1].kind, arr[1].t0, arr[1].t1) == (Gate.KIND_SWAP, 1, 3)
    assert (arr[2].kind, arr[2].ctrl_mask, list(arr[2].p)) == (Gate.KIND_U, 0b1000, [0.1, 0.2, 0.3])


@pytest.mark.parametrize("name,count", [("iqft.qasm", 10), ("quantum_lstm.qasm", 24), ("test0.qasm", 8)])
def test_qasm_fixture_loads(name, count):  # openqasm.rs:180-303
    qc = sb.openqasm.load(QASM / name)
    assert len(qc.transformations) == count


def test_qasm_iqft_matches_hand_built():  # openqasm.rs:195-217 (angles are literally -pi, -pi/4, -pi/9: SURVEY Q5)
    qc = sb.openqasm.load(QASM / "iqft.qasm")
    got = [(t.gate, t.target, t.controls.controls) for t in qc.transformations]
    assert got[0] == (Gate.H, 0, [])
    assert got[1] == (Gate.P(-PI), 1, [0])
    assert got[2] == (Gate.P(-PI / 4), 2, [0])
    assert got[3] == (Gate.P(-PI / 9), 3, [0])


def test_qasm_test0_matches_hand_built():  # openqasm.rs:219-303
    qc = sb.openqasm.load(QASM / "test0.qasm")
    got = [(t.gate, t.target) for t in qc.transformations]
    assert got == [(Gate.H, 0), (Gate.X, 1), (Gate.Y, 2), (Gate.Z, 3), (Gate.RX(1.0), 4), (Gate.RY(2.0), 5),
                   (Gate.RZ(3.0), 6), (Gate.U(1.0, 2.0, 3.0), 7)]


def test_qasm_unsupported_gate_raises():  # openqasm.rs:163 todo!()
    with pytest.raises(NotImplementedError):
        sb.openqasm.loads("OPENQASM 2.0; qreg q[2]; ccz q[0],q[1];")


def test_qasm_angle_expressions():
    assert sb.openqasm.eval_angle("-pi/4") == -PI / 4
    assert sb.openqasm.eval_angle("2*pi/8 + 0.5") == 2 * PI / 8 + 0.5
    with pytest.raises(ValueError):
        sb.openqasm.eval_angle("__import__('os')")


def test_spynoza_facade_surface():  # spynoza/src/lib.rs:411-421
    from spinoza_b200 import spynoza as sp
    for name in ("get_samples", "show_table", "run", "qubit_expectation_value", "xyz_expectation_value", "PyState",
                 "QuantumRegister", "QuantumCircuit", "PyQuantumTransformation"):
        assert hasattr(sp, name)
    q = sp.QuantumRegister(3)
    qc = sp.QuantumCircuit(q)
    qc.h(0); qc.cp(0.5, 0, 1); qc.ccx(0, 1, 2)
    assert qc.num_qubits == 3 and qc.register_sizes == [3]
    t = qc.py_transformations
    assert (t[0].name, t[0].target, t[0].controls, t[0].arg) == ("h", 0, None, None)
    assert (t[1].name, t[1].controls, t[1].arg) == ("p", [0], (0.5, 0.0, 0.0))
    assert t[2].controls == [0, 1]
