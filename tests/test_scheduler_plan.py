"""CPU tests of the fusion scheduler (abi.cu: Fuser) through spz_plan_fusion -- pure host code, no GPU.

A plan is (execution order, pass number).  It must (1) schedule every op exactly once, (2) be a semantics-preserving
reordering (checked by running both orders through the independent dense NumPy statement), (3) respect the tile
capacity in every pass, and (4) be the identity order in EXACT / KEEP_ORDER mode.
"""
import math

import numpy as np
import pytest

import spinoza_b200 as sb
from spinoza_b200 import Controls, Gate, QuantumCircuit, QuantumRegister, QuantumTransformation, workloads
from tests import _dense as D

TILE_BITS, MAX_HIGH, L_MIN = 12, 6, 6
DIAG = {Gate.KIND_Z, Gate.KIND_P, Gate.KIND_RZ}
KINDS = [Gate.KIND_H, Gate.KIND_X, Gate.KIND_Y, Gate.KIND_Z, Gate.KIND_P, Gate.KIND_RX, Gate.KIND_RY, Gate.KIND_RZ, Gate.KIND_U]


def random_circuit(n, count, seed, **kw):
    rng = np.random.default_rng(seed)
    qc = QuantumCircuit(QuantumRegister(n), **kw)
    for _ in range(count):
        r = rng.random()
        if r < 0.06:
            qc.swap(int(rng.integers(n)), int(rng.integers(n)))
            continue
        kind = KINDS[int(rng.integers(len(KINDS)))]
        g = Gate(kind, tuple(float(x) for x in rng.random(3) * 2 * math.pi))
        t = int(rng.integers(n))
        if r < 0.5:
            qc.add(QuantumTransformation(g, t))
        else:
            k = int(rng.integers(1, 4))
            cs = [int(c) for c in rng.choice([q for q in range(n) if q != t], size=min(k, n - 1), replace=False)]
            qc.add(QuantumTransformation(g, t, Controls.single(cs[0]) if len(cs) == 1 else Controls.mixed(cs, set())))
    return qc


def run_dense_order(n, psi, trs, order):
    for i in order:
        t = trs[i]
        g = t.gate
        if g.kind == Gate.KIND_SWAP:
            psi = D.apply_swap(psi, n, g.t0, g.t1)
        else:
            cm, zm = t.controls.mask(), t.controls.zeros_mask()
            if t.controls.kind == t.controls.MIXED:
                cm, zm = cm & ~zm, 0  # mc_apply drops the zeros (gates.rs:298-311)
            elif t.controls.kind != t.controls.SIGNED:
                zm = 0
            psi = D.apply_matrix(psi, n, D.matrix(g.kind, g.params), t.target, cm, zm)
    return psi


def check_pass_capacity(trs, plan):
    by_pass = {}
    for idx, p in plan:
        by_pass.setdefault(p, []).append(trs[idx])
    for p, ops in by_pass.items():
        if len(ops) == 1:
            continue  # single ops go to the direct kernel
        targets = set()
        for t in ops:
            if t.gate.kind == Gate.KIND_SWAP:
                targets |= {t.gate.t0, t.gate.t1}
            elif t.gate.kind not in DIAG:
                targets.add(t.target)
        ok = any(sum(1 for q in targets if q >= TILE_BITS - h) <= h for h in range(MAX_HIGH + 1))
        assert ok, f"pass {p}: non-diagonal targets {sorted(targets)} do not fit a 12-bit tile with >= {L_MIN} low bits"


@pytest.mark.parametrize("select", ["0", "1"], ids=["first-come-tile", "chosen-tile"])
@pytest.mark.parametrize("n,count,seed", [(13, 120, 1), (14, 200, 2), (15, 200, 3), (16, 150, 4), (14, 300, 5)])
def test_plan_is_a_valid_semantics_preserving_reordering(n, count, seed, select, monkeypatch):
    # tile selection (abi.cu: choose_tile) is on by default only from 24 qubits up; force both variants at test sizes
    monkeypatch.setenv("SPZ_TILE_SELECT", select)
    qc = random_circuit(n, count, seed)
    trs = list(qc.transformations)
    plan, n_pass = qc.plan()
    order = [i for i, _ in plan]
    swaps_same = [i for i, t in enumerate(trs) if t.gate.kind == Gate.KIND_SWAP and t.gate.t0 == t.gate.t1]
    assert sorted(order + swaps_same) == list(range(len(trs)))            # everything scheduled exactly once
    passes = [p for _, p in plan]
    assert passes == sorted(passes) and n_pass == passes[-1] + 1           # passes are emitted in order
    assert n_pass < len(trs)
    check_pass_capacity(trs, plan)
    psi0 = D.random_state(n, seed)
    want = run_dense_order(n, psi0, trs, range(len(trs)))
    got = run_dense_order(n, psi0, trs, order)
    assert np.max(np.abs(got - want)) < 1e-12
    assert order != sorted(order)                                          # and it did reorder something


@pytest.mark.parametrize("kw", [dict(exact=True), dict(reorder=False)])
def test_exact_and_keep_order_modes_preserve_program_order(kw):
    qc = random_circuit(15, 200, 7, **kw)
    plan, _ = qc.plan()
    order = [i for i, _ in plan]
    assert order == sorted(order)
    check_pass_capacity(list(qc.transformations), plan)


def test_unfused_plan_is_one_pass_per_gate():
    qc = random_circuit(14, 50, 9, fuse=False)
    plan, n_pass = qc.plan()
    assert n_pass == len(plan) and [p for _, p in plan] == list(range(len(plan)))


def test_qft_and_layered_pass_counts():
    qc = QuantumCircuit(QuantumRegister(30)); qc.qft()
    assert qc.plan()[1] == 4                      # 465 gates, 4 HBM passes
    qc = QuantumCircuit(QuantumRegister(36)); qc.qft()
    assert qc.plan()[1] <= 6
    a = QuantumCircuit(QuantumRegister(30)); workloads.random_layered_circuit(a)
    b = QuantumCircuit(QuantumRegister(30), reorder=False); workloads.random_layered_circuit(b)
    pa, pb = a.plan()[1], b.plan()[1]
    assert pa <= 32 and pb >= 80                  # DAG scheduling: 30 passes instead of 88
    check_pass_capacity(list(a.transformations), a.plan()[0])


def test_diagonal_gates_never_open_a_pass():
    n = 20
    qc = QuantumCircuit(QuantumRegister(n))
    for layer in range(5):
        for q in range(n):
            qc.rz(0.1 * (q + 1), q)
            qc.cp(0.3, q, (q + 7) % n)
            qc.z(q)
    assert qc.plan()[1] == 1


def test_measurement_is_a_barrier():
    qc = QuantumCircuit(QuantumRegister(14))
    for q in range(14):
        qc.h(q)
    qc.measure(3)
    for q in range(14):
        qc.rx(0.2, q)
    plan, _ = qc.plan()
    order = [i for i, _ in plan]
    m = order.index(14)
    assert set(order[:m]) == set(range(14)) and set(order[m + 1:]) == set(range(15, 29))


def test_sequences_on_one_qubit_stay_in_order():
    # RZ RX RZ on each qubit (the QCBM layer, gates.rs:1512-1517): per-qubit order must be kept
    n = 16
    qc = QuantumCircuit(QuantumRegister(n))
    for q in range(n):
        qc.rz(1.0, q); qc.rx(1.0, q); qc.rz(1.0, q)
    plan, n_pass = qc.plan()
    pos = {i: k for k, (i, _) in enumerate(plan)}
    for q in range(n):
        assert pos[3 * q] < pos[3 * q + 1] < pos[3 * q + 2]
    assert n_pass == 2


@pytest.mark.parametrize("a,b", [(11, 13), (13, 11), (11, 12), (10, 13), (5, 13), (12, 13)])
def test_swap_between_a_boundary_qubit_and_a_high_qubit_fits_an_empty_tile(a, b):
    """Regression: SWAP(11, b >= 12) was rejected by an empty 12-bit tile when its operands were tried in the given order
    (qubit 11 is a low tile bit only while the tile has no high bit)."""
    n = 14
    for extra in (0, 1, 4):           # alone (in-order path), and inside a window that takes the DAG path
        qc = QuantumCircuit(QuantumRegister(n))
        qc.swap(a, b)
        for k in range(extra):
            qc.h(k)
            qc.swap(a, b)
        trs = list(qc.transformations)
        plan, n_pass = qc.plan()
        assert sorted(i for i, _ in plan) == list(range(len(trs))) and n_pass >= 1
        psi = D.random_state(n, a * 16 + b)
        got = run_dense_order(n, psi.copy(), trs, [i for i, _ in plan])
        want = run_dense_order(n, psi.copy(), trs, range(len(trs)))
        assert np.max(np.abs(got - want)) < 1e-12


def test_tile_selection_keeps_layered_circuits_in_few_passes():
    """The scheduler picks each pass's tile qubits by how many ops they admit (abi.cu: choose_tile).  BASELINE config 3 at 30
    qubits (890 gates): 88 passes with the order-preserving greedy, 30 when the first ready ops claimed the tile, 17 now."""
    qc = QuantumCircuit(QuantumRegister(30))
    workloads.random_layered_circuit(qc, depth=20, seed=42)
    plan, n_pass = qc.plan()
    assert len(plan) == len(qc.transformations) == 890
    assert n_pass <= 20
    check_pass_capacity(list(qc.transformations), plan)
