"""Parity of the fused tile kernels on hardware: k_tile3 (csrc/kernels_tile3.cu, TMA, every merged-mode pass on full 12-bit
tiles) against k_tile (SPZ_TILE_V3=0) and the oracle; exact mode (k_tile) bit-identical to the oracle under every tile shape.

Merged mode is held to 1e-12 absolute against the oracle (the documented fused-gate rounding difference: phase runs are
multiplied as tables, H / RX / RY are rescaled and contracted); tests/test_tile_cpu_emulation.py checks the same kernel
bodies on the CPU.
"""
import os

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit, workloads
from tests.test_gpu_parity import oracle_ops_from, to_gpu
from tests.test_tile_cpu_emulation import reference_cells_circuit

pytestmark = pytest.mark.gpu


class tile_env:
    """Environment for one execute(): the library reads SPZ_TILE_V3 / SPZ_TILE_LMIN at every launch."""

    KEYS = ("SPZ_TILE_V3", "SPZ_TILE_LMIN")

    def __init__(self, v3=True, lmin=None):
        self.new = {"SPZ_TILE_V3": "1" if v3 else "0"}
        if lmin is not None:
            self.new["SPZ_TILE_LMIN"] = str(lmin)

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.KEYS}
        for k in self.KEYS:
            os.environ.pop(k, None)
        os.environ.update(self.new)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run(init, build, v3=True, lmin=None, **kw):
    st = to_gpu(init)
    qc = QuantumCircuit.from_state(st, **kw)
    build(qc)
    ops = oracle_ops_from(qc)
    with tile_env(v3, lmin):
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

    def high_first_low_last(qc):
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

    def near_pi(qc):  # rotations whose cosine is tiny keep their matrix; a long run of rescaled ones must stay in range
        for t in range(min(n, 12)):
            qc.rx(np.pi - 1e-9 * (t + 1), t)
            qc.ry(np.pi + 1e-7 * (t + 1), t)
        for rep in range(30):
            for t in range(min(n, 12)):
                qc.rx(np.pi - 0.02, t)

    return {"qft": qft, "layered": layered, "random": rand, "high_low": high_first_low_last, "low_high": low_first_high_last,
            "near_pi": near_pi}


@pytest.mark.parametrize("n", [12, 13, 16, 20, 22])
@pytest.mark.parametrize("name", ["qft", "layered", "random", "high_low", "low_high", "near_pi"])
def test_merged_mode_within_tolerance_of_the_oracle(n, name):
    init = orc.gen_random_state(n, 200 + n)
    build = builders(n)[name]
    (r1, i1), ops = run(init, build, v3=False, fuse=True)
    (r3, i3), _ = run(init, build, v3=True, fuse=True)
    cpu = init.clone()
    orc.execute(cpu, ops)
    for r, i in ((r1, i1), (r3, i3)):
        assert np.max(np.abs(r - cpu.reals)) <= 1e-12 and np.max(np.abs(i - cpu.imags)) <= 1e-12


@pytest.mark.parametrize("n", [12, 16, 20])
@pytest.mark.parametrize("name", ["qft", "layered", "random", "high_low", "low_high"])
def test_exact_mode_is_bit_identical_to_the_oracle(n, name):
    init = orc.gen_random_state(n, 100 + n)
    build = builders(n)[name]
    (r2, i2), ops = run(init, build, fuse=True, exact=True)
    cpu = init.clone()
    orc.execute(cpu, ops)
    assert np.array_equal(r2, cpu.reals) and np.array_equal(i2, cpu.imags)


@pytest.mark.parametrize("lmin", [4, 5])
@pytest.mark.parametrize("name", ["qft", "layered"])
def test_shorter_tile_segments(name, lmin):
    """SPZ_TILE_LMIN: passes with 7 / 8 arbitrary high qubits, i.e. TMA boxes of 2 / 1 rows of 128 bytes."""
    n = 20
    init = orc.gen_random_state(n, 400 + lmin)
    build = builders(n)[name]
    (r1, i1), ops = run(init, build, fuse=True, exact=True)              # k_tile, default segments
    (r2, i2), _ = run(init, build, lmin=lmin, fuse=True, exact=True)     # k_tile, short segments
    assert np.array_equal(r1, r2) and np.array_equal(i1, i2)
    (r3, i3), _ = run(init, build, lmin=lmin, fuse=True)                 # k_tile3, short segments
    assert np.max(np.abs(r3 - r1)) <= 1e-12 and np.max(np.abs(i3 - i1)) <= 1e-12


def test_one_launch_per_pass_and_k_tile3_is_the_kernel():
    n = 20
    init = orc.gen_random_state(n, 1)
    st = to_gpu(init)
    qc = QuantumCircuit.from_state(st, fuse=True)
    qc.qft()
    _, n_pass = qc.plan()
    before = sb.launch_count()
    qc.execute()
    st.sync()
    assert sb.launch_count() - before == n_pass


def test_qft_of_a_basis_state_closed_form_n26():
    """Above the 24-qubit threshold of the tile-choosing scheduler: high tile qubits, per-tile phase constants."""
    n = 26
    st = sb.State(n)
    x = 0x2545F491 % (1 << n)
    st.set_basis(x)
    qc = QuantumCircuit.from_state(st, fuse=True)
    qc.qft()
    qc.execute()
    st.sync()
    re, im = st.download()
    k = np.arange(1 << 16, dtype=np.int64)
    rev = np.zeros_like(k)
    for b in range(n):
        rev |= ((k >> b) & 1) << (n - 1 - b)
    want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ((x * rev) % (1 << n)) / (1 << n))
    got = re[: 1 << 16] + 1j * im[: 1 << 16]
    assert np.max(np.abs(got - want)) <= 1e-12
    assert abs(sb.norm2(st) - 1.0) <= 1e-10
