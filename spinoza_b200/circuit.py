"""Circuit IR + builder: the caller of the hot path (mirror of /root/reference/spinoza/src/circuit.rs).

`QuantumCircuit.execute()` hands the whole transformation list to `spz_execute`, where the native
scheduler batches runs of gates into shared-memory tiles (SPZ_EXEC_FUSE) -- the reference applies one
gate per full pass (circuit.rs:553-599).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

EXEC_NO_FUSE = 0
EXEC_FUSE = 1   # fused tiles, runs of diagonal gates merged into phase accumulators (few-ulp rounding difference)
EXEC_EXACT = 2  # with EXEC_FUSE: every gate applied with the reference arithmetic (bit-identical to unfused)
EXEC_KEEP_ORDER = 4  # with EXEC_FUSE: pack passes in strict program order (no commutation-aware reordering)


class QuantumRegister:
    """circuit.rs:13-51"""

    def __init__(self, size: int):
        assert size > 0  # circuit.rs:29
        self.qubits: List[int] = list(range(size))

    def __getitem__(self, i: int) -> int:
        return self.qubits[i]

    def __len__(self) -> int:
        return len(self.qubits)

    def len(self) -> int:
        return len(self.qubits)

    def update_shift(self, shift: int):  # circuit.rs:42-44
        self.qubits = [q + shift for q in self.qubits]

    def get_shift(self) -> int:  # circuit.rs:48-50
        return self.qubits[0]


class Controls:
    """circuit.rs:55-109.  kind in {NONE, SINGLE, ONES, MIXED}; SIGNED is this engine's extension: `zeros` (a subset of
    `controls`) fire on 0 -- what `Mixed { zeros }` describes and the reference's mc_apply drops (gates.rs:298-311)."""

    NONE, SINGLE, ONES, MIXED, SIGNED = range(5)
    __slots__ = ("kind", "controls", "zeros")

    def __init__(self, kind: int, controls: Sequence[int] = (), zeros: Iterable[int] = ()):
        self.kind, self.controls, self.zeros = kind, list(controls), set(zeros)

    @staticmethod
    def none() -> "Controls":
        return Controls(Controls.NONE)

    @staticmethod
    def single(c: int) -> "Controls":
        return Controls(Controls.SINGLE, [c])

    @staticmethod
    def ones(cs: Sequence[int]) -> "Controls":
        return Controls(Controls.ONES, cs)

    @staticmethod
    def mixed(controls: Sequence[int], zeros: Iterable[int]) -> "Controls":
        return Controls(Controls.MIXED, controls, zeros)

    @staticmethod
    def signed(controls: Sequence[int], zeros: Iterable[int]) -> "Controls":
        """Extension: every qubit of `controls` is a control; those listed in `zeros` must be 0, the others 1."""
        zeros = set(zeros)
        if not zeros <= set(controls):
            raise ValueError("Controls.signed: zeros must be a subset of controls")
        return Controls(Controls.SIGNED, controls, zeros)

    @staticmethod
    def _from(controls: Sequence[int], zeros: Optional[set]) -> "Controls":  # circuit.rs:73-86
        if zeros is not None:
            return Controls.mixed(controls, zeros)
        if len(controls) == 0:
            return Controls.none()
        if len(controls) == 1:
            return Controls.single(controls[0])
        return Controls.ones(controls)

    def new_with_control(self, control: int, shift: int) -> "Controls":  # circuit.rs:97-108
        controls = [c + shift for c in self.controls]
        controls.append(control)
        if not self.zeros:
            return Controls._from(controls, None)
        return Controls._from(controls, {z + shift for z in self.zeros})

    def clone(self) -> "Controls":
        return Controls(self.kind, self.controls, self.zeros)

    def mask(self) -> int:
        m = 0
        for c in self.controls:
            m |= 1 << c
        return m

    def zeros_mask(self) -> int:
        m = 0
        for z in self.zeros:
            m |= 1 << z
        return m


class QuantumTransformation:
    """circuit.rs:113-120"""

    __slots__ = ("gate", "target", "controls")

    def __init__(self, gate, target: int, controls: Optional[Controls] = None):
        self.gate, self.target, self.controls = gate, target, controls or Controls.none()


class QuantumCircuit:
    """circuit.rs:168-601"""

    def __init__(self, *registers: QuantumRegister, device: int = 0, state=None, fuse: bool = True, exact: bool = False,
                 reorder: bool = True):
        bits = 0
        self.quantum_registers_info: List[int] = []
        for r in registers:  # circuit.rs:186-190
            r.update_shift(bits)
            self.quantum_registers_info.append(len(r))
            bits += len(r)
        self.transformations: List[QuantumTransformation] = []
        # The device state is allocated on first use so that circuits can be built (and inspected) on a
        # host without a GPU; executing one still requires the CUDA engine -- there is no CPU fallback.
        self._state = state
        self.n_qubits = state.n if state is not None else bits
        self.device = device
        # QubitTracker circuit.rs:122-164
        self._measured_qubits = 0
        self._measured_qubits_vals = 0
        self.fuse = fuse
        self.exact = exact
        self.reorder = reorder

    @property
    def state(self):
        if self._state is None:
            from . import State
            self._state = State(self.n_qubits, self.device)  # State::new(bits) circuit.rs:193
        return self._state

    @state.setter
    def state(self, value):
        self._state = value
        self.n_qubits = value.n

    @classmethod
    def from_state(cls, state, fuse: bool = True, exact: bool = False, reorder: bool = True) -> "QuantumCircuit":
        """The tests' `QuantumCircuit { state, transformations: Vec::new(), .. }` literal (circuit.rs:839-844)."""
        return cls(state=state, fuse=fuse, exact=exact, reorder=reorder)

    def get_statevector(self):  # circuit.rs:200-202
        return self.state

    def is_qubit_measured(self, q: int) -> bool:  # circuit.rs:141-143
        return ((self._measured_qubits >> q) & 1) == 1

    def get_qubit_measured_val(self, q: int) -> Optional[int]:  # circuit.rs:145-153
        return ((self._measured_qubits_vals >> q) & 1) if self.is_qubit_measured(q) else None

    def inverse(self):  # circuit.rs:206-211
        self.transformations.reverse()
        for qt in self.transformations:
            qt.gate = qt.gate.inverse()

    def add(self, tr: QuantumTransformation):  # circuit.rs:547-549
        self.transformations.append(tr)

    def _g(self, gate, target, controls=None):
        self.add(QuantumTransformation(gate, target, controls))

    # builder methods circuit.rs:213-434
    def measure(self, target: int):
        from . import Gate
        self._g(Gate.M, target)

    def swap(self, t0: int, t1: int):
        from . import Gate
        self._g(Gate.SWAP(t0, t1), 0)

    def x(self, target: int):
        from . import Gate
        self._g(Gate.X, target)

    def y(self, target: int):
        from . import Gate
        self._g(Gate.Y, target)

    def z(self, target: int):
        from . import Gate
        self._g(Gate.Z, target)

    def h(self, target: int):
        from . import Gate
        self._g(Gate.H, target)

    def p(self, angle: float, target: int):
        from . import Gate
        self._g(Gate.P(angle), target)

    def rx(self, angle: float, target: int):
        from . import Gate
        self._g(Gate.RX(angle), target)

    def ry(self, angle: float, target: int):
        from . import Gate
        self._g(Gate.RY(angle), target)

    def rz(self, angle: float, target: int):
        from . import Gate
        self._g(Gate.RZ(angle), target)

    def u(self, theta: float, phi: float, lam: float, target: int):
        from . import Gate
        self._g(Gate.U(theta, phi, lam), target)

    def cx(self, control: int, target: int):
        from . import Gate
        self._g(Gate.X, target, Controls.single(control))

    def ccx(self, control1: int, control2: int, target: int):
        from . import Gate
        self._g(Gate.X, target, Controls.ones([control1, control2]))

    def ch(self, control: int, target: int):
        from . import Gate
        self._g(Gate.H, target, Controls.single(control))

    def cy(self, control: int, target: int):
        from . import Gate
        self._g(Gate.Y, target, Controls.single(control))

    def cp(self, angle: float, control: int, target: int):
        from . import Gate
        self._g(Gate.P(angle), target, Controls.single(control))

    def crx(self, angle: float, control: int, target: int):
        from . import Gate
        self._g(Gate.RX(angle), target, Controls.single(control))

    def cry(self, angle: float, control: int, target: int):
        from . import Gate
        self._g(Gate.RY(angle), target, Controls.single(control))

    def crz(self, angle: float, control: int, target: int):
        from . import Gate
        self._g(Gate.RZ(angle), target, Controls.single(control))

    def cu(self, theta: float, phi: float, lam: float, control: int, target: int):
        from . import Gate
        self._g(Gate.U(theta, phi, lam), target, Controls.single(control))

    def bit_flip_noise(self, prob: float, target: int):
        from . import Gate
        self._g(Gate.BitFlipNoise(prob), target)

    def iqft(self, targets: Sequence[int]):  # circuit.rs:438-445
        from . import PI
        for j in reversed(range(len(targets))):
            self.h(targets[j])
            for k in reversed(range(j)):
                self.cp(-PI / 2.0 ** (j - k), targets[j], targets[k])

    def qft(self, n: Optional[int] = None):
        """QFT-n as BASELINE defines it: `qc.iqft(&(0..n).rev()); qc.inverse()`, appended."""
        n = self.n_qubits if n is None else n
        tmp = QuantumCircuit.__new__(QuantumCircuit)
        tmp.transformations = []
        QuantumCircuit.iqft(tmp, list(reversed(range(n))))
        QuantumCircuit.inverse(tmp)
        self.transformations.extend(tmp.transformations)

    def append(self, circuit: "QuantumCircuit", reg: QuantumRegister):  # circuit.rs:448-460
        assert len(reg) == sum(circuit.quantum_registers_info)
        for tr in circuit.transformations:
            # controls are cloned unshifted, as in the reference (circuit.rs:457)
            self.add(QuantumTransformation(tr.gate, reg.get_shift() + tr.target, tr.controls.clone()))

    def c_append(self, circuit: "QuantumCircuit", c: int, reg: QuantumRegister):  # circuit.rs:463-476
        assert not (reg.get_shift() <= c < reg.get_shift() + len(reg))
        for tr in circuit.transformations:
            self.add(QuantumTransformation(tr.gate, reg.get_shift() + tr.target,
                                           tr.controls.new_with_control(c, reg.get_shift())))

    def mc_append(self, circuit: "QuantumCircuit", controls: Sequence[int], reg: QuantumRegister):  # circuit.rs:479-511
        assert len(set(controls)) == len(controls)
        lo, hi = reg.get_shift(), reg.get_shift() + len(reg)
        for c in controls:
            if lo <= c < hi:
                raise ValueError(f"control {c} should not be in: Range(start: {lo} end: {hi})")
        for control in controls:
            for tr in circuit.transformations:
                self.add(QuantumTransformation(tr.gate, reg.get_shift() + tr.target,
                                               tr.controls.new_with_control(control, reg.get_shift())))

    def expanded(self) -> "QuantumCircuit":
        """The same circuit with every SIGNED control written as X on its zero-controls around the all-ones gate -- what
        `execute` schedules for fused passes and sharded registers (csrc/abi.cu: SPZ_CTRL_SIGNED).  `plan()` needs this form."""
        from . import Gate
        out = QuantumCircuit.__new__(QuantumCircuit)
        out.__dict__.update(self.__dict__)
        out.transformations = []
        for tr in self.transformations:
            if tr.controls.kind != Controls.SIGNED or not tr.controls.zeros:
                out.transformations.append(tr)
                continue
            zs = sorted(tr.controls.zeros)
            out.transformations += [QuantumTransformation(Gate.X, z, Controls.none()) for z in zs]
            out.transformations.append(QuantumTransformation(tr.gate, tr.target, Controls._from(tr.controls.controls, None)))
            out.transformations += [QuantumTransformation(Gate.X, z, Controls.none()) for z in zs]
        return out

    def _encode(self):
        from . import _Op
        n = len(self.transformations)
        arr = (_Op * max(n, 1))()
        for i, tr in enumerate(self.transformations):
            op = arr[i]
            g = tr.gate
            op.kind, op.target, op.t0, op.t1 = g.kind, tr.target, g.t0, g.t1
            for j, v in enumerate(g.params[:3]):
                op.p[j] = v
            op.ctrl_kind = tr.controls.kind
            op.ctrl_mask = tr.controls.mask()
            op.zeros_mask = tr.controls.zeros_mask()
        return arr, n

    def _flags(self) -> int:
        return (EXEC_FUSE | (EXEC_EXACT if self.exact else 0) | (0 if self.reorder else EXEC_KEEP_ORDER)) if self.fuse else EXEC_NO_FUSE

    def plan(self):
        """How `execute` would schedule the current list (no GPU needed): [(op index, pass number)] in execution order."""
        from . import _lib, _check
        arr, n = self._encode()
        order = (C.c_int32 * max(n, 1))()
        passes = (C.c_int32 * max(n, 1))()
        n_pass = C.c_int32()
        _check(_lib.spz_plan_fusion(self.n_qubits, arr, n, self._flags(), order, passes, C.byref(n_pass)))
        return [(order[i], passes[i]) for i in range(n) if order[i] >= 0], n_pass.value

    def execute(self):
        """circuit.rs:552-600.  Drains the transformation list (re-entrant on the same state)."""
        from . import _lib, _check
        arr, n = self._encode()
        m = C.c_uint64(self._measured_qubits)
        v = C.c_uint64(self._measured_qubits_vals)
        flags = self._flags()
        self.transformations = []  # drain(..): the list is consumed even if a gate "panics"
        _check(_lib.spz_execute(self.state._h, arr, n, flags, C.byref(m), C.byref(v)))
        self._measured_qubits, self._measured_qubits_vals = m.value, v.value
