"""`spynoza`-compatible facade (mirror of /root/reference/spynoza/src/lib.rs) over the B200 engine.

    from spinoza_b200 import spynoza as sp
    q = sp.QuantumRegister(3); qc = sp.QuantumCircuit(q); qc.h(0); qc.cx(0, 1)
    state = sp.run(qc)                 # lib.rs:405-409
    print(sp.show_table(state)); print(state[0])          # (re, im), lib.rs:55-57
    sp.get_samples(state, 1000, 10000)                    # {basis index: count}, lib.rs:395-403

Differences, all forced by device residency: `state_vector` is a handle to the live device state (the
reference clones the whole state, lib.rs:205-211); `get_samples` draws `reservoir_size` exact inverse-CDF
samples (the reference's weighted reservoir converges to the same distribution; `num_tests` is accepted and
ignored).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

from . import (QuantumRegister, State, qubit_expectation_value, sample,  # noqa: F401
               xyz_expectation_value)
from .circuit import Controls
from .circuit import QuantumCircuit as _QuantumCircuit
from .circuit import QuantumTransformation


class PyState:
    """lib.rs:35-58"""

    def __init__(self, state: State):
        self.data = state

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, i: int):
        a = self.data.amp(i)
        return (a.real, a.imag)

    def __str__(self) -> str:
        re, im = self.data.download()
        return "".join(f"{a} + i{b}\n" for a, b in zip(re, im))


class PyQuantumTransformation:
    """lib.rs:91-136"""

    _NAMES = ["h", "m", "x", "y", "z", "p", "rx", "ry", "rz", "swap", "u", "unitary", "bit_flip_noise"]

    def __init__(self, tr: QuantumTransformation):
        self.target = tr.target
        self.controls = list(tr.controls.controls) or None
        self.name = self._NAMES[tr.gate.kind]
        p = list(tr.gate.params) + [0.0] * 3
        self.arg = tuple(p[:3]) if tr.gate.params else None

    def __str__(self) -> str:
        return f"name: {self.name}\ntarget: {self.target}\narg: {self.arg}\ncontrols: {self.controls}\n"


class QuantumCircuit(_QuantumCircuit):
    """lib.rs:171-373: same builder methods as the core circuit plus the read-only properties."""

    @property
    def num_qubits(self) -> int:
        return self.n_qubits

    @property
    def register_sizes(self) -> List[int]:
        return list(self.quantum_registers_info)

    @property
    def state_vector(self) -> PyState:
        return PyState(self.state)

    @property
    def py_transformations(self) -> List[PyQuantumTransformation]:
        return [PyQuantumTransformation(t) for t in self.transformations]


def run(qc: QuantumCircuit) -> PyState:
    """lib.rs:405-409"""
    qc.execute()
    return PyState(qc.state)


def get_samples(state, reservoir_size: int, num_tests: int = 0, seed: int = 0) -> Dict[int, int]:
    """lib.rs:395-403 -> histogram {outcome: count} with `reservoir_size` entries in total."""
    st = state.data if isinstance(state, PyState) else state
    out = sample(st, reservoir_size, seed=seed)
    hist: Dict[int, int] = {}
    for o in out.tolist():
        hist[o] = hist.get(o, 0) + 1
    return hist


def show_table(state, rows: int = 16) -> str:
    """utils.rs:50-88 `to_table`: outcome, amplitude, magnitude, phase, probability of the first 16 basis states."""
    st = state.data if isinstance(state, PyState) else state
    n = st.n
    cnt = min(rows, len(st))
    re, im = st.download(0, cnt)
    lines = [f"{'Outcome':>8} | {'Amplitude':>27} | {'Magnitude':>9} | {'Amplitude Bar':<16} | {'Probability':>11}"]
    for i in range(cnt):
        mag = math.hypot(re[i], im[i])
        bar = "#" * int(round(mag * 16))
        lines.append(f"{i:>3} = {format(i, f'0{n}b'):>{max(n, 2)}} | {re[i]:+.8f} {im[i]:+.8f}i | {mag:9.5f} | {bar:<16} | {mag * mag:11.6f}")
    return "\n".join(lines)
