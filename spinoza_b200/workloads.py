"""The synthetic circuits of BASELINE.json's configs, as builders over `QuantumCircuit`.

* `qft`                    -- config 1 / 4: QFT-n := qc.iqft(&(0..n).rev()); qc.inverse()   (n + n(n-1)/2 gates)
* `random_layered_circuit` -- config 3: depth x [1q rotation on every qubit, entanglers on pairs (i, i+1), i = l mod 2,
                              alternating CNOT / CP(angle)], kinds and angles from splitmix64(seed)
* `qcbm`                   -- the reference's own circuit benchmark (benches/benchmark.rs:12-58, criterion "qcbm": n = 25, depth 9,
                              ring of CX pairs (i, i+1 mod n)): RX RZ on every qubit, entangler, depth x [RZ RX RZ, entangler], RZ RX.
                              The reference draws its angles from StdRng(42), which only Rust reproduces; here from splitmix64(42)
* `tiled_qasm`             -- config 5: a 4-qubit OpenQASM program repeated over disjoint 4-qubit blocks
The generators are deterministic and shared by bench.py and the parity tests, so "the gate list is the shared input".
"""
from __future__ import annotations

import math
from typing import List, Tuple

from . import Gate
from .circuit import Controls, QuantumCircuit, QuantumTransformation

_MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.x = seed & _MASK

    def next(self) -> int:
        self.x = (self.x + 0x9E3779B97F4A7C15) & _MASK
        z = self.x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def u01(self) -> float:
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)


def qft(qc: QuantumCircuit, n: int = None) -> int:
    qc.qft(n)
    n = qc.n_qubits if n is None else n
    return n + n * (n - 1) // 2


def random_layered_ops(n: int, depth: int = 20, seed: int = 42) -> List[Tuple]:
    """[(kind, target, control_or_None, angle)] -- kinds: 'rx','ry','rz','cx','cp'."""
    rng = SplitMix64(seed)
    ops = []
    for layer in range(depth):
        for q in range(n):
            kind = ("rx", "ry", "rz")[rng.next() % 3]
            ops.append((kind, q, None, rng.u01() * 2.0 * math.pi))
        for idx, i in enumerate(range(layer % 2, n - 1, 2)):
            if idx % 2 == 0:
                ops.append(("cx", i + 1, i, 0.0))
            else:
                ops.append(("cp", i + 1, i, rng.u01() * 2.0 * math.pi))
    return ops


def random_layered_circuit(qc: QuantumCircuit, depth: int = 20, seed: int = 42) -> int:
    ops = random_layered_ops(qc.n_qubits, depth, seed)
    for kind, t, c, ang in ops:
        if kind == "rx":
            qc.rx(ang, t)
        elif kind == "ry":
            qc.ry(ang, t)
        elif kind == "rz":
            qc.rz(ang, t)
        elif kind == "cx":
            qc.cx(c, t)
        else:
            qc.cp(ang, c, t)
    return len(ops)


def qcbm(qc: QuantumCircuit, depth: int = 9, seed: int = 42) -> int:
    """benches/benchmark.rs:12-58 build_circuit(nqubits, depth, pairs) with pairs = [(i, (i + 1) % n)] (:155)."""
    n = qc.n_qubits
    rng = SplitMix64(seed)
    pairs = [(i, (i + 1) % n) for i in range(n)]
    before = len(qc.transformations)

    def entangler():
        for a, b in pairs:
            if a != b:
                qc.cx(a, b)

    for k in range(n):                      # first_rotation
        qc.rx(rng.u01(), k)
        qc.rz(rng.u01(), k)
    entangler()
    for _ in range(depth):                  # mid_rotation + entangler
        for k in range(n):
            qc.rz(rng.u01(), k)
            qc.rx(rng.u01(), k)
            qc.rz(rng.u01(), k)
        entangler()
    for k in range(n):                      # last_rotation
        qc.rz(rng.u01(), k)
        qc.rx(rng.u01(), k)
    return len(qc.transformations) - before


def tiled_qasm(qc: QuantumCircuit, qasm_text: str, block: int = 4) -> int:
    """Append the (block-qubit) program once per disjoint block of `block` qubits of qc's register."""
    from . import openqasm
    before = len(qc.transformations)
    for b in range(qc.n_qubits // block):
        openqasm.loads(qasm_text, shift=b * block, into=qc)
    return len(qc.transformations) - before


def multi_controlled_layer(qc: QuantumCircuit) -> int:
    """config 5's extra gates: mc X / P / RX / RY with 2-3 controls, inside and outside the reference's safe domain
    (SURVEY.md 2.3 B2), through Controls::Mixed as `execute` requires (circuit.rs:588-596)."""
    n = qc.n_qubits
    before = len(qc.transformations)
    cases = [(Gate.X, [0, 1], 2), (Gate.P(0.7), [0, 1], 2), (Gate.RX(0.9), [1, 2], 0), (Gate.RY(1.1), [n - 1, 0], n // 2),
             (Gate.X, [n - 1, n - 2, 1], 3 % n), (Gate.P(2.1), [2, n - 1, n // 2], 0)]
    for g, cs, t in cases:
        cs = [c for c in dict.fromkeys(cs) if c != t and c < n]
        if len(cs) >= 2 and t < n:
            qc.add(QuantumTransformation(g, t, Controls.mixed(cs, set())))
    return len(qc.transformations) - before
