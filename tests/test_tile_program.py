"""CPU tests of the fused-pass compiler (abi.cu: Fuser::compile) -- pure host code, no GPU.

`spz_debug_compile_pass` serialises the micro-program (TileInstr / TileGroup / TileTerm, engine.h) that spz_execute would
upload for one pass.  This file interprets that program in NumPy, following the execution model documented at the top of
csrc/kernels_tile.cu (LAYOUT / GATE / DIAG / RUN, phase accumulators F0..F4, per-tile group factors), and checks that the
whole plan -- every pass, in order -- equals the gate-by-gate dense statement of the circuit.  It covers what the planner
tests cannot: masks, register layouts, control splitting (outer / thread / register), SWAP lowering and the folding of
diagonal gates into phase terms.
"""
import ctypes as C
import math

import numpy as np
import pytest

import spinoza_b200 as sb
from spinoza_b200 import Gate, QuantumCircuit, QuantumRegister, workloads
from tests import _dense as D
from tests.test_scheduler_plan import random_circuit, run_dense_order

TI_LAYOUT, TI_GATE, TI_DIAG, TI_RUN = 0, 1, 2, 3

INSTR = np.dtype([
    ("op", "<i4"), ("kind", "<i4"), ("rpos", "<i4"), ("t_mask", "<u4"),
    ("reg_cmask", "<u4"), ("thr_cmask", "<u4"), ("t_where", "<i4"), ("outer_target", "<i4"),
    ("rbit", "<i4", 4),
    ("outer_cmask", "<u8"), ("const_hi", "<u4"), ("has_f0", "<i4"),
    ("s", "<f8", 7), ("f0", "<f8", 2), ("f1", "<f8", 2), ("pad", "<f8"),
])
GROUP = np.dtype([("thr", "<u4"), ("m", "<u4"), ("first", "<i4"), ("count", "<i4")])
TERM = np.dtype([("outer", "<u8"), ("thr", "<u4"), ("m", "<u4"), ("fr", "<f8"), ("fi", "<f8")])


def compile_pass(qc, pass_index):
    """-> ('tile', plan dict, instrs, groups, terms) | ('direct',) | None."""
    arr, n = qc._encode()
    buf = (C.c_char * (8 << 20))()
    used = C.c_int64()
    sb._check(sb._lib.spz_debug_compile_pass(qc.n_qubits, arr, n, qc._flags(), pass_index, buf, len(buf), C.byref(used)))
    raw = bytes(buf[: used.value])
    hdr = np.frombuffer(raw, dtype="<i4", count=16)
    if hdr[0] == 2:
        return None
    if hdr[0] == 1:
        return ("direct",)
    assert hdr[15] == INSTR.itemsize, "TileInstr layout changed: update the dtype in this test"
    ni, ng, nt = int(hdr[12]), int(hdr[13]), int(hdr[14])
    off = 64
    instrs = np.frombuffer(raw, dtype=INSTR, count=ni, offset=off); off += ni * INSTR.itemsize
    groups = np.frombuffer(raw, dtype=GROUP, count=ng, offset=off); off += ng * GROUP.itemsize
    terms = np.frombuffer(raw, dtype=TERM, count=nt, offset=off); off += nt * TERM.itemsize
    assert off == used.value
    plan = {"T": int(hdr[1]), "L": int(hdr[2]), "high": [int(h) for h in hdr[4:4 + int(hdr[3])]]}
    return ("tile", plan, instrs, groups, terms)


def gate_matrix_from_scalars(kind, s):
    """The 2x2 matrix the butterfly of gate_math.cuh applies, from the instruction's scalars."""
    r = math.sqrt(0.5)
    if kind == Gate.KIND_H:
        return np.array([[r, r], [r, -r]], dtype=complex)
    if kind == Gate.KIND_X:
        return np.array([[0, 1], [1, 0]], dtype=complex)
    if kind == Gate.KIND_Y:
        return np.array([[0, -1j], [1j, 0]], dtype=complex)
    if kind == Gate.KIND_RX:  # s = (cos, -sin)(theta/2): off-diagonal i * s[1]
        return np.array([[s[0], 1j * s[1]], [1j * s[1], s[0]]], dtype=complex)
    if kind == Gate.KIND_RY:  # s = (sin, cos)(theta/2)
        return np.array([[s[1], -s[0]], [s[0], s[1]]], dtype=complex)
    if kind == Gate.KIND_U:
        return np.array([[s[0], s[1] + 1j * s[2]], [s[3] + 1j * s[4], s[5] + 1j * s[6]]], dtype=complex)
    raise AssertionError(f"kind {kind} cannot be a TI_GATE")


class TileMachine:
    """All tiles at once: every per-thread quantity of k_tile becomes an array over the 2^n absolute indices."""

    def __init__(self, n, plan, groups, terms, exact, lazy=False, allow_rank_constants=False):
        self.n, self.T, self.L, self.high = n, plan["T"], plan["L"], plan["high"]
        self.exact = exact
        self.lazy = lazy  # k_tile2: a butterfly on register bit r flushes accumulator F_{r+1} only
        self.allow_rank_constants = allow_rank_constants  # sharded registers: diagonal target on a bit of the rank
        idx = np.arange(1 << n, dtype=np.uint64)
        self.qubit_of_bit = list(range(self.L)) + self.high
        assert len(self.qubit_of_bit) == self.T and len(set(self.qubit_of_bit)) == self.T
        assert all(h >= self.L for h in self.high) and self.high == sorted(self.high)
        j = np.zeros(1 << n, dtype=np.uint32)
        tile_abs = 0
        for b, q in enumerate(self.qubit_of_bit):
            j |= (((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(b)).astype(np.uint32)
            tile_abs |= 1 << q
        self.j = j
        self.base = idx & np.uint64(~tile_abs & ((1 << 64) - 1))
        # per-tile group factors (the cooperative reduction before the first barrier)
        self.gfac, self.gany = [], []
        for g in groups:
            f = np.ones(1 << n, dtype=complex)
            any_ = np.zeros(1 << n, dtype=bool)
            for t in terms[g["first"]: g["first"] + g["count"]]:
                assert t["thr"] == g["thr"] and t["m"] == g["m"]
                assert int(t["outer"]) & tile_abs == 0, "a term's outer mask must not contain tile qubits"
                hit = (self.base & t["outer"]) == t["outer"]
                f = np.where(hit, f * complex(t["fr"], t["fi"]), f)
                any_ |= hit
            self.gfac.append(f)
            self.gany.append(any_)
        self.groups = groups
        self.R = None
        self.F = [np.ones(1 << n, dtype=complex) for _ in range(5)]

    def flush(self, psi):
        psi = psi * self.F[0]
        for c in range(1, 5):
            psi = np.where((self.k >> (c - 1)) & 1 == 1, psi * self.F[c], psi)
        self.F = [np.ones(1 << self.n, dtype=complex) for _ in range(5)]
        return psi

    def layout(self, psi, rbit):
        if self.R is not None:
            psi = self.flush(psi)
        rbit = [int(b) for b in rbit]
        assert rbit == sorted(set(rbit)) and len(rbit) == 4 and 0 <= rbit[0] and rbit[3] < self.T, rbit
        self.R = rbit
        k = np.zeros_like(self.j)
        rmask = 0
        for i, b in enumerate(rbit):
            k |= ((self.j >> b) & 1) << i
            rmask |= 1 << b
        self.k = k
        self.tj = self.j & np.uint32(~rmask & 0xFFFFFFFF)
        self.rmask = rmask
        return psi

    def run(self, psi, ins):
        g = int(ins["rpos"])
        counts = [int(c) for c in ins["rbit"]] + [int(ins["reg_cmask"]), int(ins["thr_cmask"])]
        class_m = [0, 1, 2, 4, 8]
        for c in range(5):
            for _ in range(counts[c]):
                gd = self.groups[g]
                assert gd["m"] == class_m[c], "group filed under the wrong accumulator class"
                assert int(gd["thr"]) & self.rmask == 0, "thread mask overlaps the register bits"
                hit = self.gany[g] & ((self.tj & gd["thr"]) == gd["thr"])
                self.F[c] = np.where(hit, self.F[c] * self.gfac[g], self.F[c])
                g += 1
        for _ in range(counts[5]):
            gd = self.groups[g]
            m = int(gd["m"])
            assert bin(m).count("1") >= 2 and m < 16
            hit = self.gany[g] & ((self.tj & gd["thr"]) == gd["thr"]) & ((self.k & m) == m)
            psi = np.where(hit, psi * self.gfac[g], psi)
            g += 1
        self.next_group = g
        return psi

    def gate(self, psi, ins):
        rpos = int(ins["rpos"])
        if self.lazy:
            psi = np.where((self.k >> rpos) & 1 == 1, psi * self.F[rpos + 1], psi)
            self.F[rpos + 1] = np.ones(1 << self.n, dtype=complex)
        else:
            psi = self.flush(psi)
        q = self.qubit_of_bit[self.R[rpos]]
        idx = np.arange(1 << self.n, dtype=np.uint64)
        ocm = ins["outer_cmask"]
        assert int(ocm) & sum(1 << x for x in self.qubit_of_bit) == 0
        assert int(ins["thr_cmask"]) & self.rmask == 0
        k0 = self.k & np.uint32(~(1 << rpos) & 0xF)
        km_hit = ((np.uint32(ins["t_mask"]) >> k0) & 1) == 1
        sel0 = ((self.base & ocm) == ocm) & ((self.tj & ins["thr_cmask"]) == ins["thr_cmask"]) & km_hit & (((idx >> np.uint64(q)) & np.uint64(1)) == 0)
        s0 = idx[sel0]
        s1 = s0 | np.uint64(1 << q)
        m = gate_matrix_from_scalars(int(ins["kind"]), ins["s"])
        out = psi.copy()
        a, b = psi[s0], psi[s1]
        out[s0] = m[0, 0] * a + m[0, 1] * b
        out[s1] = m[1, 0] * a + m[1, 1] * b
        return out

    def diag(self, psi, ins):
        assert self.exact, "merged mode folds diagonal gates into TI_RUN; a TI_DIAG must not appear"
        kind, tw, s = int(ins["kind"]), int(ins["t_where"]), ins["s"]
        ocm = ins["outer_cmask"]
        ok = ((self.base & ocm) == ocm) & ((self.tj & ins["thr_cmask"]) == ins["thr_cmask"]) & ((self.k & ins["reg_cmask"]) == ins["reg_cmask"])
        if tw == 0 and ins["const_hi"] != 0:
            assert self.allow_rank_constants  # rank-bit targets only exist in sharded lowering
            hi = np.full(1 << self.n, ins["const_hi"] == 2)
        elif tw == 0:
            hi = ((self.base >> np.uint64(ins["outer_target"])) & np.uint64(1)) == 1
        elif tw == 1:
            hi = (self.tj & ins["t_mask"]) != 0
        else:
            hi = (self.k & ins["t_mask"]) != 0
        if kind == Gate.KIND_Z:
            f_hi, f_lo = -1.0, 1.0
        elif kind == Gate.KIND_P:
            f_hi, f_lo = complex(s[0], s[1]), 1.0
        else:
            assert kind == Gate.KIND_RZ
            f_hi, f_lo = complex(s[0], s[1]), complex(s[0], -s[1])
        return np.where(ok, psi * np.where(hi, f_hi, f_lo), psi)


def interpret(n, psi, compiled, exact, lazy=False, allow_rank_constants=False):
    _, plan, instrs, groups, terms = compiled
    assert plan["T"] == plan["L"] + len(plan["high"]) and plan["T"] <= 12 and len(plan["high"]) <= 8
    assert len(groups) <= 2048
    assert instrs[0]["op"] == TI_LAYOUT, "a program starts by choosing a register layout"
    tm = TileMachine(n, plan, groups, terms, exact, lazy, allow_rank_constants)
    seen_groups = 0
    for ins in instrs:
        op = int(ins["op"])
        if op == TI_LAYOUT:
            psi = tm.layout(psi, ins["rbit"])
        elif op == TI_RUN:
            assert not exact
            assert ins["rpos"] == seen_groups, "runs consume the group table front to back"
            psi = tm.run(psi, ins)
            seen_groups = tm.next_group
        elif op == TI_GATE:
            psi = tm.gate(psi, ins)
        else:
            assert op == TI_DIAG
            psi = tm.diag(psi, ins)
    assert seen_groups == len(groups)
    return tm.flush(psi)


def run_plan(qc, psi, lazy=False):
    """Execute the circuit the way spz_execute schedules it, interpreting every fused pass."""
    n = qc.n_qubits
    trs = list(qc.transformations)
    plan, n_pass = qc.plan()
    by_pass = {}
    for idx, p in plan:
        by_pass.setdefault(p, []).append(idx)
    n_tile = 0
    for p in range(n_pass):
        comp = compile_pass(qc, p)
        assert comp is not None
        if comp[0] == "direct":
            assert len(by_pass[p]) == 1
            psi = run_dense_order(n, psi, trs, by_pass[p])
        else:
            assert len(by_pass[p]) > 1
            psi = interpret(n, psi, comp, qc.exact, lazy)
            n_tile += 1
    assert compile_pass(qc, n_pass) is None
    return psi, n_tile


@pytest.mark.parametrize("select", ["0", "1"], ids=["first-come-tile", "chosen-tile"])
@pytest.mark.parametrize("lazy", [False, True], ids=["k_tile", "k_tile2-lazy-flush"])
@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("n,count,seed", [(13, 150, 11), (14, 200, 12), (15, 250, 13), (16, 120, 14)])
def test_compiled_passes_equal_the_dense_statement(n, count, seed, exact, lazy, select, monkeypatch):
    monkeypatch.setenv("SPZ_TILE_SELECT", select)
    qc = random_circuit(n, count, seed, exact=exact)
    trs = list(qc.transformations)
    psi0 = D.random_state(n, seed)
    got, n_tile = run_plan(qc, psi0.copy(), lazy)
    want = run_dense_order(n, psi0.copy(), trs, range(len(trs)))
    assert n_tile >= 1
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_qft_passes_fold_every_controlled_phase_into_runs():
    n = 16
    qc = QuantumCircuit(QuantumRegister(n))
    qc.qft()
    trs = list(qc.transformations)
    plan, n_pass = qc.plan()
    total_terms = 0
    for p in range(n_pass):
        comp = compile_pass(qc, p)
        if comp[0] != "tile":
            continue
        _, _, instrs, groups, terms = comp
        assert not np.any(instrs["op"] == TI_DIAG)
        total_terms += len(terms)
        assert len(groups) <= len(terms)
    n_cp = sum(1 for t in trs if t.gate.kind == Gate.KIND_P)
    assert total_terms == n_cp  # one phase term per controlled-phase gate (P has no f0 term)
    psi0 = D.random_state(n, 5)
    want = run_dense_order(n, psi0.copy(), trs, range(len(trs)))
    for lazy in (False, True):
        got, _ = run_plan(qc, psi0.copy(), lazy)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_layered_rotation_circuit_and_swaps():
    n = 15
    qc = QuantumCircuit(QuantumRegister(n))
    workloads.random_layered_circuit(qc, depth=6, seed=9)
    qc.swap(0, 14); qc.swap(3, 4); qc.swap(13, 12)
    trs = list(qc.transformations)
    psi0 = D.random_state(n, 2)
    got, n_tile = run_plan(qc, psi0.copy())
    want = run_dense_order(n, psi0.copy(), trs, range(len(trs)))
    assert n_tile >= 1
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_diagonal_gate_on_an_outer_qubit_becomes_an_outer_term():
    n = 16
    qc = QuantumCircuit(QuantumRegister(n))
    for t in range(4):
        qc.h(t)
    qc.rz(0.37, 15)          # target outside any tile the H's need
    qc.cp(0.91, 15, 14)      # control and target both outside
    qc.cp(0.5, 2, 15)        # register/thread control, outer target
    comp = compile_pass(qc, 0)
    assert comp[0] == "tile"
    _, plan, instrs, groups, terms = comp
    assert plan["high"] == []
    outers = sorted(int(t["outer"]) for t in terms)
    # RZ -> f0 term (no outer bit) + f1 term on bit 15; CP(15,14) -> bits 14|15; CP(2 -> 15) -> bit 15
    assert outers == [0, 1 << 15, 1 << 15, (1 << 14) | (1 << 15)]
    trs = list(qc.transformations)
    psi0 = D.random_state(n, 3)
    got, _ = run_plan(qc, psi0.copy())
    want = run_dense_order(n, psi0.copy(), trs, range(len(trs)))
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
