"""Sharded execute() replayed on the CPU -- pure host code on the library side, NumPy on this side, no GPU.

`spz_debug_compile_sharded` returns, for one rank, every step `spz_execute` would take on a register sharded over `world`
ranks: fused passes as tile micro-programs, single ops (with rank-constant diagonal ops already resolved), exchanges.  This
file replays all ranks in lockstep -- the interpreter of tests/test_tile_program.py for the programs, a bit swap of the
physical index for the exchanges -- and compares the un-permuted result with the dense gate-by-gate statement.

It covers what only the sharded path produces and no single-GPU test can reach: ops skipped on ranks whose global control
bit is 0, diagonal gates on rank bits folded into tile programs as constants (const_hi), windows closed by exchanges,
Belady victim choice from the execute look-ahead, and the scheduler (both ways of choosing a tile) inside those windows.
"""
import ctypes as C
import math

import numpy as np
import pytest

import spinoza_b200 as sb
from spinoza_b200 import Gate, QuantumCircuit, QuantumRegister, workloads
from tests import _dense as D
from tests.test_scheduler_plan import random_circuit, run_dense_order
from tests.test_tile_program import GROUP, INSTR, TERM, gate_matrix_from_scalars, interpret

DIAG = {Gate.KIND_Z, Gate.KIND_P, Gate.KIND_RZ}


def steps_of_rank(qc, world, rank):
    arr, n_ops = qc._encode()
    buf = (C.c_char * (32 << 20))()
    used = C.c_int64()
    sb._check(sb._lib.spz_debug_compile_sharded(qc.n_qubits, world, rank, arr, n_ops, qc._flags(), buf, len(buf), C.byref(used)))
    raw = bytes(buf[: used.value])
    n_steps = int(np.frombuffer(raw, dtype="<i8", count=1)[0])
    off = 8
    steps = []
    for _ in range(n_steps):
        h = np.frombuffer(raw, dtype="<i4", count=20, offset=off)
        off += 80
        kind = int(h[0])
        if kind == 0:
            assert h[15] == INSTR.itemsize
            ni, ng, nt = int(h[12]), int(h[13]), int(h[14])
            instrs = np.frombuffer(raw, dtype=INSTR, count=ni, offset=off); off += ni * INSTR.itemsize
            groups = np.frombuffer(raw, dtype=GROUP, count=ng, offset=off); off += ng * GROUP.itemsize
            terms = np.frombuffer(raw, dtype=TERM, count=nt, offset=off); off += nt * TERM.itemsize
            plan = {"T": int(h[1]), "L": int(h[2]), "high": [int(x) for x in h[4:4 + int(h[3])]]}
            steps.append(("tile", plan, instrs, groups, terms))
        elif kind == 1:
            cmask = int(np.frombuffer(raw, dtype="<u8", count=1, offset=off)[0])
            s = np.frombuffer(raw, dtype="<f8", count=7, offset=off + 8).copy()
            off += 64
            steps.append(("op", int(h[1]), int(h[2]), int(h[3]), int(h[4]), cmask, s))
        elif kind == 2:
            steps.append(("exchange", int(h[1]), int(h[2])))
        elif kind == 4:  # the basis state was re-placed before the first op: the new permutation
            steps.append(("relabel", tuple(int(x) for x in np.frombuffer(raw, dtype="<i4", count=64, offset=off))))
            off += 256
        else:
            raise AssertionError("measurement steps are not replayed here")
    perm = [int(x) for x in np.frombuffer(raw, dtype="<i4", count=64, offset=off)]
    assert off + 256 == used.value
    return steps, perm


def apply_single(n_local, psi, step):
    _, kind, target, t2, const_hi, cmask, s = step
    if kind == Gate.KIND_SWAP:
        return D.apply_swap(psi, n_local, target, t2)
    if kind in DIAG:
        if kind == Gate.KIND_Z:
            f_lo, f_hi = 1.0, -1.0
        elif kind == Gate.KIND_P:
            f_lo, f_hi = 1.0, complex(s[0], s[1])
        else:
            f_lo, f_hi = complex(s[0], -s[1]), complex(s[0], s[1])
        idx = np.arange(1 << n_local, dtype=np.int64)
        ctrl = (idx & cmask) == cmask
        if const_hi >= 0:                                   # target is a bit of the rank: one constant for the whole shard
            return np.where(ctrl, psi * (f_hi if const_hi else f_lo), psi)
        hi = ((idx >> target) & 1) == 1
        return np.where(ctrl, psi * np.where(hi, f_hi, f_lo), psi)
    assert const_hi < 0
    return D.apply_matrix(psi, n_local, gate_matrix_from_scalars(kind, s), target, cmask)


def replay(qc, world, psi0):
    """-> logical state vector after the sharded execution."""
    n = qc.n_qubits
    g = world.bit_length() - 1
    n_local = n - g
    per_rank = [steps_of_rank(qc, world, r) for r in range(world)]
    perms = [p for _, p in per_rank]
    assert all(p == perms[0] for p in perms), "the plan must evolve identically on every rank"
    shards = [psi0[r << n_local:(r + 1) << n_local].copy() for r in range(world)]
    cursors = [0] * world
    stats = {"tile": 0, "op": 0, "exchange": 0, "relabel": 0}
    if per_rank[0][0] and per_rank[0][0][0][0] == "relabel":
        # free placement of a basis state (dist_place_basis): every rank relabels the same way before its first op; the
        # physical layout of the (basis) state follows the new permutation
        first = [steps[0] for steps, _ in per_rank]
        assert all(f == first[0] for f in first)
        new_perm = first[0][1][:n]
        assert sorted(new_perm) == list(range(n))
        i = np.arange(1 << n, dtype=np.int64)
        p_idx = np.zeros_like(i)
        for q in range(n):
            p_idx |= ((i >> q) & 1) << new_perm[q]
        phys = np.zeros_like(psi0)
        phys[p_idx] = psi0
        assert np.count_nonzero(psi0) == 1, "only a basis state may be re-placed"
        shards = [phys[r << n_local:(r + 1) << n_local].copy() for r in range(world)]
        cursors = [1] * world
        stats["relabel"] = 1
    while True:
        pending = []
        for r in range(world):
            steps = per_rank[r][0]
            while cursors[r] < len(steps) and steps[cursors[r]][0] != "exchange":
                st = steps[cursors[r]]
                if st[0] == "tile":
                    shards[r] = interpret(n_local, shards[r], st, qc.exact, allow_rank_constants=True)
                    stats["tile"] += 1
                else:
                    shards[r] = apply_single(n_local, shards[r], st)
                    stats["op"] += 1
                cursors[r] += 1
            pending.append(steps[cursors[r]] if cursors[r] < len(steps) else None)
        if all(p is None for p in pending):
            break
        assert all(p is not None and p == pending[0] for p in pending), "every rank must reach the same exchange"
        _, gbit, lq = pending[0]
        phys = np.concatenate(shards)                        # physical index = rank << n_local | local index
        phys = D.apply_swap(phys, n, n_local + gbit, lq)      # the exchange swaps a rank bit with a local bit
        shards = [phys[r << n_local:(r + 1) << n_local].copy() for r in range(world)]
        stats["exchange"] += 1
        cursors = [c + 1 for c in cursors]
    phys = np.concatenate(shards)
    # logical amplitude at index i sits at the physical index whose bit perm[q] is bit q of i
    perm = perms[0][:n]
    i = np.arange(1 << n, dtype=np.int64)
    p_idx = np.zeros_like(i)
    for q in range(n):
        p_idx |= ((i >> q) & 1) << perm[q]
    return phys[p_idx], stats


def dense(qc, psi0):
    trs = list(qc.transformations)
    return run_dense_order(qc.n_qubits, psi0.copy(), trs, range(len(trs)))


def layered(n, depth, seed, **kw):
    qc = QuantumCircuit(QuantumRegister(n), **kw)
    workloads.random_layered_circuit(qc, depth=depth, seed=seed)
    return qc


def qft(n, **kw):
    qc = QuantumCircuit(QuantumRegister(n), **kw)
    qc.qft()
    return qc


@pytest.mark.parametrize("window", ["0", "1"], ids=["window-closes-at-exchange", "window-spans-exchanges"])
@pytest.mark.parametrize("select", ["0", "1"], ids=["first-come-tile", "chosen-tile"])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("name", ["qft", "layered", "random"])
def test_sharded_fused_execution_equals_the_dense_statement(name, exact, world, select, window, monkeypatch):
    monkeypatch.setenv("SPZ_TILE_SELECT", select)
    monkeypatch.setenv("SPZ_DIST_WINDOW", window)
    n = 14 + (world.bit_length() - 1) - 1          # 13 local qubits: real 12-bit tiles on every shard
    qc = {"qft": lambda: qft(n, exact=exact), "layered": lambda: layered(n, 8, 3, exact=exact),
          "random": lambda: random_circuit(n, 180, 17, exact=exact)}[name]()
    psi0 = D.random_state(n, 9)
    got, stats = replay(qc, world, psi0)
    np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)
    assert stats["tile"] >= world and stats["exchange"] >= 1, stats


def test_eight_ranks_qft_with_chosen_tiles(monkeypatch):
    monkeypatch.setenv("SPZ_TILE_SELECT", "1")
    n, world = 16, 8
    qc = qft(n)
    psi0 = D.random_state(n, 12)
    got, stats = replay(qc, world, psi0)
    np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)
    # 3 global qubits need their H, and the victims evicted for them (chosen by farthest next use) need theirs later:
    # between 3 and 6 exchanges; the Belady choice gets away with 4
    assert 3 <= stats["exchange"] <= 6, stats


@pytest.mark.parametrize("window", ["0", "1"])
@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("world", [2, 4])
def test_exchange_fused_with_its_gate_keeps_the_op_accounting(world, fuse, window, monkeypatch):
    monkeypatch.setenv("SPZ_DIST_WINDOW", window)
    """SPZ_DIST_FUSE_GATE=1: the gate that triggers an exchange is applied by the exchange kernel and must not be applied
    again (or dropped) by the scheduler.  The dry run reports it as "exchange, then that gate as a step of its own"."""
    monkeypatch.setenv("SPZ_DIST_FUSE_GATE", "1")
    n = 13 + world.bit_length() - 1
    qc = qft(n, fuse=fuse)
    workloads.random_layered_circuit(qc, depth=5, seed=2)
    psi0 = D.random_state(n, 13)
    got, stats = replay(qc, world, psi0)
    np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)
    assert stats["exchange"] >= 1
    # the exchanges asked for by an uncontrolled gate (QFT's H gates) are directly followed by that gate alone, on the bit the
    # exchange brought in; exchanges asked for by controlled gates (the CX of the layered part) stay plain
    steps, _ = steps_of_rank(qc, world, 0)
    fused = 0
    for i, st in enumerate(steps[:-1]):
        nxt = steps[i + 1]
        if st[0] == "exchange" and nxt[0] == "op" and nxt[2] == st[2] and nxt[5] == 0 and nxt[1] not in DIAG:
            fused += 1
    assert fused >= world.bit_length() - 1, (fused, [s[0] for s in steps])


def test_unfused_sharded_execution(monkeypatch):
    n, world = 10, 4
    qc = random_circuit(n, 120, 23, fuse=False)
    psi0 = D.random_state(n, 4)
    got, stats = replay(qc, world, psi0)
    np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)
    assert stats["tile"] == 0 and stats["op"] > 0


def test_world_of_one_is_the_unsharded_plan():
    n = 13
    qc = layered(n, 6, 5)
    psi0 = D.random_state(n, 6)
    got, stats = replay(qc, 1, psi0)
    np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)
    assert stats["exchange"] == 0


def test_diagonal_gates_on_rank_bits_never_exchange():
    """QFT's controlled phases touch global qubits long before their H: they must all be rank constants."""
    n, world = 15, 4
    qc = QuantumCircuit(QuantumRegister(n))
    for t in range(n):
        qc.rz(0.1 * (t + 1), t)
    for c in range(n - 2):
        qc.cp(0.3 + 0.01 * c, c, n - 1)       # target: a global qubit
        qc.cp(0.2, n - 2, c)                  # control: a global qubit
    psi0 = D.random_state(n, 8)
    got, stats = replay(qc, world, psi0)
    np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)
    assert stats["exchange"] == 0


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("world", [2, 4])
def test_signed_controls_on_a_sharded_register(world, fuse):
    """Negative controls (SPZ_CTRL_SIGNED) on local and on global qubits: lowered to X . gate . X, so a zero-control that lives
    in the rank bits costs exchanges but must give the dense statement with that control at 0."""
    n = 13 + world.bit_length() - 1
    rng = np.random.default_rng(5 + world)
    qc = random_circuit(n, 40, 91, fuse=fuse)
    kinds = [Gate.KIND_X, Gate.KIND_P, Gate.KIND_RX, Gate.KIND_RY, Gate.KIND_H, Gate.KIND_RZ]
    for i in range(12):
        t = int(rng.integers(n))
        others = [q for q in range(n) if q != t]
        cs = [int(c) for c in rng.choice(others, size=int(rng.integers(1, 4)), replace=False)]
        if i % 3 == 0 and t != n - 1:
            cs = list({*cs, n - 1})           # a global qubit among the controls ...
        zs = {cs[-1]} | {c for c in cs if rng.random() < 0.4}   # ... and among the zeros
        qc.add(sb.QuantumTransformation(Gate(kinds[i % len(kinds)], (0.3 + 0.1 * i, 0.0, 0.0)), t, sb.Controls.signed(cs, zs)))
    psi0 = D.random_state(n, 14)
    got, stats = replay(qc, world, psi0)
    np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)


@pytest.mark.parametrize("select", ["0", "1"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_basis_state_is_placed_by_looking_ahead(world, select, monkeypatch):
    """Free placement (dist_place_basis): a register that is still a basis state gets its qubit permutation from the op list --
    the g qubits whose first non-diagonal use comes latest become the rank bits.  QFT then needs g exchanges, not g + 1."""
    monkeypatch.setenv("SPZ_TILE_SELECT", select)
    g = world.bit_length() - 1
    n = 13 + g
    x = 0x9E3779B97F4A7C15 % (1 << n)
    psi0 = np.zeros(1 << n, dtype=complex)
    psi0[x] = 1.0
    qc = qft(n)
    want = dense(qc, psi0)
    _, plain = replay(qc, world, psi0)                     # no hint: the qubits stay where they are
    assert plain["relabel"] == 0 and plain["exchange"] == g + 1
    monkeypatch.setenv("SPZ_DEBUG_BASIS", str(x))
    got, stats = replay(qc, world, psi0)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    assert stats["relabel"] == 1 and stats["exchange"] == g, stats
    steps, perm = steps_of_rank(qc, world, 0)
    assert sorted(steps[0][1][q] for q in range(g)) == list(range(n - g, n))   # QFT uses qubits 0 .. g-1 last: they start global
    # switched off, or a circuit whose global qubits are already the last ones used: nothing is relabelled
    monkeypatch.setenv("SPZ_DIST_PLACE", "0")
    _, off = replay(qc, world, psi0)
    assert off["relabel"] == 0 and off["exchange"] == g + 1
    monkeypatch.delenv("SPZ_DIST_PLACE")
    qc2 = QuantumCircuit(QuantumRegister(n))
    for t in range(n):
        qc2.h(t)                                           # ascending: the top qubits come last anyway
    got2, st2 = replay(qc2, world, psi0)
    np.testing.assert_allclose(got2, dense(qc2, psi0), rtol=0, atol=1e-12)
    assert st2["relabel"] == 0


@pytest.mark.parametrize("world", [2, 4])
def test_placement_with_layered_and_random_circuits(world, monkeypatch):
    g = world.bit_length() - 1
    n = 13 + g
    x = 12345 % (1 << n)
    psi0 = np.zeros(1 << n, dtype=complex)
    psi0[x] = 1.0
    monkeypatch.setenv("SPZ_DEBUG_BASIS", str(x))
    for qc in (layered(n, 6, 9), random_circuit(n, 150, 5), random_circuit(n, 60, 6, fuse=False)):
        got, stats = replay(qc, world, psi0)
        np.testing.assert_allclose(got, dense(qc, psi0), rtol=0, atol=1e-12)
