"""Parity of the opt-in second-generation tile kernel (csrc/kernels_tile2.cu, SPZ_TILE_V2=1) against k_tile and the oracle.

OPT-IN: k_tile2 was written after round 1's GPU budget was spent and has not run on hardware yet, so these tests only run
with SPZ_TEST_TILE_V2=1 (they must not be able to turn the default GPU suite red).  First thing to run in round 2:

    SPZ_TEST_TILE_V2=1 timeout 600 python -m pytest tests/test_gpu_tile_v2.py -x -q -m gpu
    SPZ_TILE_V2=1 python tools/profile_qft.py 30            # against the 79-85 ms of k_tile

What they assert: exact mode is bit-identical to k_tile (and therefore to the unfused path and the oracle); merged mode stays
within 1e-12 of the oracle (lazy flushing changes the order of the phase multiplications, not the mathematics --
tests/test_tile_program.py checks that statement on the CPU).
"""
import os

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit, workloads
from tests.test_gpu_parity import oracle_ops_from, to_gpu
from tests.test_tile_cpu_emulation import reference_cells_circuit

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SPZ_TEST_TILE_V2") != "1", reason="opt-in: SPZ_TEST_TILE_V2=1")]


class tile_v2:
    """Environment for one execute(): the library reads SPZ_TILE_V2 / SPZ_TILE_V2_DIRECT at every launch."""

    def __init__(self, on, direct=None, lmin=None):
        self.new = {"SPZ_TILE_V2": "1" if on else "0"}
        if direct is not None:
            self.new["SPZ_TILE_V2_DIRECT"] = str(direct)
        if lmin is not None:
            self.new["SPZ_TILE_LMIN"] = str(lmin)

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in ("SPZ_TILE_V2", "SPZ_TILE_V2_DIRECT", "SPZ_TILE_LMIN")}
        os.environ.pop("SPZ_TILE_V2_DIRECT", None)
        os.environ.pop("SPZ_TILE_LMIN", None)
        os.environ.update(self.new)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run(init, build, v2, direct=None, lmin=None, **kw):
    st = to_gpu(init)
    qc = QuantumCircuit.from_state(st, **kw)
    build(qc)
    ops = oracle_ops_from(qc)
    with tile_v2(v2, direct, lmin):
        qc.execute()
        st.sync()
    return st.download(), ops


def builders(n):
    def qft(qc):
        qc.qft()

    def layered(qc):
        workloads.random_layered_circuit(qc, depth=8, seed=7)

    def rand(qc):
        src = reference_cells_circuit(n, 300, 21)  # only gate x control cells the oracle (= the reference) supports
        for t in src.transformations:
            qc.add(t)

    def high_first_low_last(qc):  # first layout all >= 4 (direct load), last layout contains low bits (staged store)
        for t in (11, 10, 9, 8):
            qc.h(t)
        qc.cp(0.3, 11, 2)
        for t in (0, 1, 2, 3):
            qc.ry(0.1 * (t + 1), t)

    def low_first_high_last(qc):
        for t in (0, 1, 2, 3):
            qc.rx(0.2 * (t + 1), t)
        qc.cp(0.7, 1, 9)
        for t in (n - 1, n - 2, n - 3, n - 4):
            qc.h(t)
        qc.cx(n - 1, n - 2)

    return {"qft": qft, "layered": layered, "random": rand, "high_low": high_first_low_last, "low_high": low_first_high_last}


@pytest.mark.parametrize("n", [12, 13, 16, 20, 22])
@pytest.mark.parametrize("name", ["qft", "layered", "random", "high_low", "low_high"])
def test_exact_mode_is_bit_identical_to_k_tile(n, name):
    init = orc.gen_random_state(n, 100 + n)
    build = builders(n)[name]
    (r1, i1), ops = run(init, build, False, fuse=True, exact=True)
    (r2, i2), _ = run(init, build, True, fuse=True, exact=True)
    assert np.array_equal(r1, r2) and np.array_equal(i1, i2)
    cpu = init.clone()
    orc.execute(cpu, ops)
    assert np.array_equal(r2, cpu.reals) and np.array_equal(i2, cpu.imags)


@pytest.mark.parametrize("n", [12, 13, 16, 20, 22])
@pytest.mark.parametrize("name", ["qft", "layered", "random", "high_low", "low_high"])
def test_merged_mode_within_tolerance_of_the_oracle(n, name):
    init = orc.gen_random_state(n, 200 + n)
    build = builders(n)[name]
    (r1, i1), ops = run(init, build, False, fuse=True)
    (r2, i2), _ = run(init, build, True, fuse=True)
    cpu = init.clone()
    orc.execute(cpu, ops)
    for r, i in ((r1, i1), (r2, i2)):
        assert np.max(np.abs(r - cpu.reals)) <= 1e-12 and np.max(np.abs(i - cpu.imags)) <= 1e-12


@pytest.mark.parametrize("level", [0, 2, 3])
@pytest.mark.parametrize("name", ["qft", "layered", "low_high", "high_low"])
def test_every_transfer_level_is_bit_identical_in_exact_mode(name, level):
    n = 18
    init = orc.gen_random_state(n, 300 + level)
    build = builders(n)[name]
    (r1, i1), _ = run(init, build, False, fuse=True, exact=True)
    (r2, i2), _ = run(init, build, True, direct=level, fuse=True, exact=True)
    assert np.array_equal(r1, r2) and np.array_equal(i1, i2)


@pytest.mark.parametrize("v2", [False, True], ids=["k_tile", "k_tile2"])
@pytest.mark.parametrize("lmin", [4, 5])
@pytest.mark.parametrize("name", ["qft", "layered"])
def test_shorter_tile_segments_are_bit_identical_in_exact_mode(name, lmin, v2):
    """SPZ_TILE_LMIN: passes with 7 / 8 arbitrary high qubits (never run on hardware in round 1, where the maximum was 6)."""
    n = 20
    init = orc.gen_random_state(n, 400 + lmin)
    build = builders(n)[name]
    (r1, i1), _ = run(init, build, False, fuse=True, exact=True)
    (r2, i2), _ = run(init, build, v2, lmin=lmin, fuse=True, exact=True)
    assert np.array_equal(r1, r2) and np.array_equal(i1, i2)


def test_v2_is_actually_selected_and_counts_one_launch_per_pass():
    n = 20
    init = orc.gen_random_state(n, 1)
    st = to_gpu(init)
    qc = QuantumCircuit.from_state(st, fuse=True)
    qc.qft()
    _, n_pass = qc.plan()
    before = sb.launch_count()
    with tile_v2(True):
        qc.execute()
        st.sync()
    assert sb.launch_count() - before == n_pass


# ---- exchange-spanning windows on sharded registers (SPZ_DIST_WINDOW=1, also opt-in until run on hardware) ----------------

@pytest.mark.parametrize("select", ["0", "1"])
@pytest.mark.parametrize("n,world", [(15, 2), (16, 4), (17, 8)])
def test_windows_spanning_exchanges_match_the_oracle(n, world, select, monkeypatch):
    from spinoza_b200.distributed import DistState
    from tests.test_gpu_dist import gather, run_group, upload_shards
    monkeypatch.setenv("SPZ_DIST_WINDOW", "1")
    monkeypatch.setenv("SPZ_TILE_SELECT", select)
    init = orc.gen_random_state(n, 46)
    states = DistState.create_local_group(n, world)
    upload_shards(states, init)
    box = {}

    def body(rank, s):
        q = QuantumCircuit.from_state(s, fuse=True)
        workloads.random_layered_circuit(q, depth=12, seed=42)
        q.qft()
        if rank == 0:
            box["ops"] = oracle_ops_from(q)
        q.execute()
        s.sync()
        box[rank] = s.stats()["exchanges"]
    run_group(states, body)
    re, im = gather(states)
    cpu = init.clone()
    orc.execute(cpu, box["ops"])
    assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12
    assert len({box[r] for r in range(world)}) == 1   # every rank ran the same exchanges


# ---- exchange fused with the gate that asked for it (SPZ_DIST_FUSE_GATE=1, kernels_xgate.cuh; opt-in) --------------------

@pytest.mark.parametrize("n,world", [(12, 2), (14, 4)])
def test_fused_exchange_gate_is_bit_identical_gate_by_gate(n, world, monkeypatch):
    """The 1-qubit sweep of the bench on shards: every non-diagonal gate on a global qubit takes the fused kernel."""
    from spinoza_b200.distributed import DistState
    from tests.test_gpu_dist import gather, run_group, upload_shards
    monkeypatch.setenv("SPZ_DIST_FUSE_GATE", "1")
    monkeypatch.setenv("SPZ_XG_CTAS", "8")   # several shards share one GPU here: every CTA of every shard must be resident
    init = orc.gen_random_state(n, 47)
    states = DistState.create_local_group(n, world)
    upload_shards(states, init)
    seq = [(orc.H, ()), (orc.RX, (1.0,)), (orc.RY, (0.4,)), (orc.X, ()), (orc.Y, ()), (orc.U, (0.1, 0.2, 0.3)), (orc.RZ, (1.0,))]
    cpu = init.clone()
    for kind, p in seq:
        for t in range(n):
            orc.apply(kind, cpu, t, p)

    def body(rank, s):
        for kind, p in seq:
            for t in range(n):
                sb.apply(sb.Gate(kind, p), s, t)
        s.sync()
    run_group(states, body)
    re, im = gather(states)
    assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags)


@pytest.mark.parametrize("window", ["0", "1"])
@pytest.mark.parametrize("n,world", [(15, 2), (16, 4)])
def test_fused_exchange_gate_inside_execute(n, world, window, monkeypatch):
    from spinoza_b200.distributed import DistState
    from tests.test_gpu_dist import gather, run_group, upload_shards
    monkeypatch.setenv("SPZ_DIST_FUSE_GATE", "1")
    monkeypatch.setenv("SPZ_DIST_WINDOW", window)
    monkeypatch.setenv("SPZ_XG_CTAS", "8")
    init = orc.gen_random_state(n, 48)
    states = DistState.create_local_group(n, world)
    upload_shards(states, init)
    box = {}

    def body(rank, s):
        q = QuantumCircuit.from_state(s, fuse=True)
        q.qft()
        workloads.random_layered_circuit(q, depth=6, seed=42)
        if rank == 0:
            box["ops"] = oracle_ops_from(q)
        q.execute()
        s.sync()
    run_group(states, body)
    re, im = gather(states)
    cpu = init.clone()
    orc.execute(cpu, box["ops"])
    assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12


# ---- clone of a sharded register (spz_dist_copy_from; new in the last hours of round 1, so opt-in like the rest) ---------

def test_clone_of_a_sharded_register_is_independent_and_keeps_the_permutation():
    from spinoza_b200.distributed import DistState
    from tests.test_gpu_dist import gather, run_group, upload_shards
    n, world = 14, 4
    init = orc.gen_random_state(n, 49)
    states = DistState.create_local_group(n, world)
    upload_shards(states, init)

    def first(rank, s):
        for t in (n - 1, n - 2, 3):          # two global targets: the permutation is no longer the identity
            sb.apply(sb.Gate.H, s, t)
        s.sync()
    run_group(states, first)
    clones = DistState.clone_local_group(states)
    assert clones[0].perm() == states[0].perm() != list(range(n))

    def second(rank, s):
        sb.apply(sb.Gate.X, s, 0)            # only the originals move on
        s.sync()
    run_group(states, second)
    cpu = init.clone()
    for t in (n - 1, n - 2, 3):
        orc.apply(orc.H, cpu, t)
    re, im = gather(clones)
    assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags)
    orc.apply(orc.X, cpu, 0)
    re, im = gather(states)
    assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags)
