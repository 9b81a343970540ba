"""BASELINE.json configs 1, 3 and 5 as parity cases at sizes the oracle finishes in seconds (config 2 is the bench
sweep, config 4 the sharded QFT in test_gpu_dist.py).  Same generated gate list on both sides."""
import math
from pathlib import Path

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import Gate, QuantumCircuit, workloads
from spinoza_b200.distributed import DistState
from tests.test_gpu_dist import gather, run_group, upload_shards
from tests.test_gpu_parity import oracle_ops_from, to_gpu

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("n", [12, 16, 20, 22])
def test_config3_random_layered_circuit(n):
    init = orc.gen_random_state(n, 42)
    fused, exact, plain = to_gpu(init), to_gpu(init), to_gpu(init)
    qc = QuantumCircuit.from_state(fused, fuse=True)
    n_ops = workloads.random_layered_circuit(qc, depth=20, seed=42)
    assert n_ops == len(workloads.random_layered_ops(n, 20, 42)) > 20 * n
    ops = oracle_ops_from(qc)
    qc.execute()
    for st, kw in ((exact, dict(fuse=True, exact=True)), (plain, dict(fuse=False))):
        q2 = QuantumCircuit.from_state(st, **kw)
        workloads.random_layered_circuit(q2, depth=20, seed=42)
        q2.execute()
    cpu = init.clone()
    orc.execute(cpu, ops)
    pr, pi = plain.download(); er, ei = exact.download(); fr, fi = fused.download()
    assert np.array_equal(pr, cpu.reals) and np.array_equal(pi, cpu.imags)          # unfused == oracle, bit for bit
    assert np.array_equal(er, pr) and np.array_equal(ei, pi)                        # fused-exact == unfused
    assert np.max(np.abs(fr - cpu.reals)) <= 1e-12 and np.max(np.abs(fi - cpu.imags)) <= 1e-12
    assert abs(sb.norm2(fused) - 1.0) < 1e-10


@pytest.mark.parametrize("n", [13, 20])
def test_qcbm_circuit_benchmark_fused_exact_and_unfused(n):
    """The reference's own circuit benchmark (benches/benchmark.rs:12-58 `qcbm`: RX RZ | ring of CX | depth x [RZ RX RZ | ring] |
    RZ RX, depth 9) through execute: unfused == oracle bit for bit, fused-exact == unfused, merged within 1e-12."""
    init = orc.State(n)  # the bench starts from |0..0> (QuantumCircuit::new)
    states = {k: to_gpu(init) for k in ("fused", "exact", "plain")}
    kws = {"fused": dict(fuse=True), "exact": dict(fuse=True, exact=True), "plain": dict(fuse=False)}
    ops = None
    for k, st in states.items():
        qc = QuantumCircuit.from_state(st, **kws[k])
        assert workloads.qcbm(qc, depth=9, seed=42) == 41 * n
        ops = oracle_ops_from(qc)
        qc.execute()
    cpu = init.clone()
    orc.execute(cpu, ops)
    pr, pi = states["plain"].download(); er, ei = states["exact"].download(); fr, fi = states["fused"].download()
    assert np.array_equal(pr, cpu.reals) and np.array_equal(pi, cpu.imags)
    assert np.array_equal(er, pr) and np.array_equal(ei, pi)
    assert np.max(np.abs(fr - cpu.reals)) <= 1e-12 and np.max(np.abs(fi - cpu.imags)) <= 1e-12
    assert abs(sb.norm2(states["fused"]) - 1.0) < 1e-10


@pytest.mark.parametrize("n", [16, 20, 22])
def test_config3_with_chosen_tiles(n, monkeypatch):
    """From 24 qubits up the scheduler chooses each pass's tile qubits by how many ops they admit (abi.cu: choose_tile); force
    that mode at sizes the oracle can check.  Merged mode within 1e-12 of the oracle, and the plan must really be shorter."""
    init = orc.gen_random_state(n, 44)
    passes = {}
    for select in ("0", "1"):
        monkeypatch.setenv("SPZ_TILE_SELECT", select)
        gpu = to_gpu(init)
        qc = QuantumCircuit.from_state(gpu, fuse=True)
        workloads.random_layered_circuit(qc, depth=20, seed=42)
        ops = oracle_ops_from(qc)
        passes[select] = qc.plan()[1]
        before = sb.launch_count()
        qc.execute()
        gpu.sync()
        assert sb.launch_count() - before == passes[select]
        cpu = init.clone()
        orc.execute(cpu, ops)
        re, im = gpu.download()
        assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12
    assert passes["1"] <= passes["0"]


@pytest.mark.parametrize("n,world", [(14, 2), (16, 4)])
def test_config3_two_shards_equal_one_shard(n, world):
    init = orc.gen_random_state(n, 43)
    one = to_gpu(init)
    qc = QuantumCircuit.from_state(one, fuse=True, exact=True)
    workloads.random_layered_circuit(qc, depth=20, seed=42)
    qc.execute()
    states = DistState.create_local_group(n, world)
    upload_shards(states, init)

    def body(rank, s):
        q = QuantumCircuit.from_state(s, fuse=True, exact=True)
        workloads.random_layered_circuit(q, depth=20, seed=42)
        q.execute()
        s.sync()
    run_group(states, body)
    re, im = gather(states)
    r1, i1 = one.download()
    assert np.array_equal(re, r1) and np.array_equal(im, i1)


@pytest.mark.parametrize("select", ["0", "1"])
@pytest.mark.parametrize("n,world", [(15, 2), (16, 4)])
def test_config3_sharded_dag_schedule_matches_the_oracle(n, world, select, monkeypatch):
    """Merged (DAG-scheduled) execution on shards, with both ways of choosing a pass's tile: the windows between exchanges
    hold rank-constant diagonal ops (const_hi) next to local ones, which only the sharded path produces."""
    monkeypatch.setenv("SPZ_TILE_SELECT", select)
    init = orc.gen_random_state(n, 45)
    states = DistState.create_local_group(n, world)
    upload_shards(states, init)
    ops_box = {}

    def body(rank, s):
        q = QuantumCircuit.from_state(s, fuse=True)
        workloads.random_layered_circuit(q, depth=12, seed=42)
        if rank == 0:
            ops_box["ops"] = oracle_ops_from(q)
        q.execute()
        s.sync()
    run_group(states, body)
    re, im = gather(states)
    cpu = init.clone()
    orc.execute(cpu, ops_box["ops"])
    assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12


@pytest.mark.parametrize("n", [8, 12, 16, 20])
def test_config5_tiled_qasm_mc_gates_and_sampling(n):
    for name in ("quantum_lstm.qasm", "iqft.qasm"):
        text = (GOLDEN / name).read_text()
        init = orc.gen_random_state(n, 5 + n)
        gpu = to_gpu(init)
        qc = QuantumCircuit.from_state(gpu, fuse=True)
        n_q = workloads.tiled_qasm(qc, text)
        n_mc = workloads.multi_controlled_layer(qc)
        assert n_q == (n // 4) * (24 if "lstm" in name else 10) and n_mc >= 4
        ops = oracle_ops_from(qc)
        qc.execute()
        cpu = init.clone()
        orc.execute(cpu, ops)
        re, im = gpu.download()
        assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12
        # sampling: seed 42; compare with the oracle's CDF walk and with the exact distribution
        shots = 1 << 14
        u = orc.uniforms(42, shots)
        got = sb.sample(gpu, shots, u01=u)
        want = orc.sample_cdf(cpu, u)
        assert np.count_nonzero(got != want) <= 2
        if n <= 12:
            p = cpu.reals ** 2 + cpu.imags ** 2
            counts = np.bincount(got, minlength=1 << n)
            keep = p * shots > 5
            chi2 = np.sum((counts[keep] - shots * p[keep]) ** 2 / (shots * p[keep]))
            dof = int(np.count_nonzero(keep))
            assert chi2 < dof + 6 * math.sqrt(2 * dof)


def test_config5_at_scale_properties():
    """n = 28 (the 32-qubit run of config 5 needs 69 GB; same code path): norm, and sampling of a known state."""
    n = 28
    s = sb.State(n)
    qc = QuantumCircuit.from_state(s, fuse=True)
    workloads.tiled_qasm(qc, (GOLDEN / "quantum_lstm.qasm").read_text())
    workloads.multi_controlled_layer(qc)
    qc.execute()
    assert abs(sb.norm2(s) - 1.0) < 1e-10
    out = sb.sample(s, 1 << 16, seed=42)
    assert out.min() >= 0 and out.max() < (1 << n)
    # each 4-qubit block is an independent product factor: block 0's marginal must match a 4-qubit run
    small = sb.State(4)
    q4 = sb.openqasm.loads((GOLDEN / "quantum_lstm.qasm").read_text(), fuse=False)
    q4.state = small
    q4.execute()
    # marginal of block 6 (untouched by the mc layer: qubits 24..27 are controls at most) from the samples
    p_small = np.abs(small.amps()) ** 2
    blk = (out >> 8) & 0xF   # block 2: qubits 8..11 are never targets or controls of the mc layer
    freq = np.bincount(blk, minlength=16) / len(out)
    assert np.max(np.abs(freq - p_small)) < 0.01


def test_config1_qft20_matches_cpu_reference():
    n = 20
    s = sb.State(n)
    qc = QuantumCircuit.from_state(s, fuse=True)
    assert workloads.qft(qc) == 210
    ops = oracle_ops_from(qc)
    qc.execute()
    cpu = orc.State(n)
    orc.set_threads(orc.max_threads())
    try:
        orc.execute(cpu, ops)
    finally:
        orc.set_threads(1)
    re, im = s.download()
    assert np.max(np.abs(re - cpu.reals)) <= 1e-12 and np.max(np.abs(im - cpu.imags)) <= 1e-12
    assert np.max(np.abs(re - 2.0 ** (-n / 2))) < 1e-12  # QFT|0..0> is the uniform superposition
