"""GPU tests of the sharded register (csrc/dist.cu).

* `local group`: 2/4/8 shards inside this process on one GPU, one host thread per shard -- the same kernels,
  epoch flags and exchange protocol as one-process-per-GPU, minus IPC.  Runs on a 1-GPU box.
* `torchrun`: 2 real processes on 2 GPUs with CUDA IPC over NVLink (skipped when fewer than 2 GPUs).
Parity: the gathered, un-permuted state must equal the unsharded single-GPU engine bit for bit (same
arithmetic per amplitude) and the oracle within 1e-12.
"""
import math
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import Gate, QuantumCircuit, workloads
from spinoza_b200.distributed import DistState, unpermute
from tests.test_gpu_parity import GATES, G, build_circuit, oracle_ops_from, random_ops, run_dense

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def run_group(states, fn):
    """Run fn(rank, state) on one thread per shard (calls block on device-side handshakes)."""
    with ThreadPoolExecutor(max_workers=len(states)) as ex:
        futs = [ex.submit(fn, r, s) for r, s in enumerate(states)]
        return [f.result(timeout=300) for f in futs]


def upload_shards(states, cpu):
    n_local = states[0].n_local
    for r, s in enumerate(states):
        s.upload(cpu.reals[r << n_local:(r + 1) << n_local], cpu.imags[r << n_local:(r + 1) << n_local])


def gather(states):
    parts = [s.download() for s in states]
    re = np.concatenate([p[0] for p in parts]); im = np.concatenate([p[1] for p in parts])
    perms = [s.perm() for s in states]
    assert all(p == perms[0] for p in perms)
    return unpermute(re, im, perms[0])


@pytest.mark.parametrize("n,world", [(6, 2), (9, 4), (12, 8), (14, 2)])
def test_direct_gates_sharded_vs_single_gpu(n, world):
    cpu = orc.gen_random_state(n, 11 * n + world)
    ref = sb.State.from_arrays(cpu.reals, cpu.imags)
    states = DistState.create_local_group(n, world)
    upload_shards(states, cpu)
    rng = np.random.default_rng(n + world)
    seq = []
    for _ in range(60):
        kind, p = GATES[int(rng.integers(len(GATES)))]
        t = int(rng.integers(n))
        c = int(rng.integers(n - 1)); c += c >= t
        seq.append((kind, p, t, c, rng.random() < 0.5))
    seq += [(orc.H, (), n - 1, 0, False), (orc.RX, (1.0,), n - 1, 0, False), (orc.X, (), n - 2, n - 1, True)]

    def body(rank, s):
        for kind, p, t, c, ctl in seq:
            if ctl and kind != orc.Z:
                sb.c_apply(G(kind, p), s, c, t)
            else:
                sb.apply(G(kind, p), s, t)
        s.sync()
    run_group(states, body)
    for kind, p, t, c, ctl in seq:
        if ctl and kind != orc.Z:
            sb.c_apply(G(kind, p), ref, c, t)
        else:
            sb.apply(G(kind, p), ref, t)
    re, im = gather(states)
    rre, rim = ref.download()
    assert np.array_equal(re, rre) and np.array_equal(im, rim)
    assert states[0].stats()["exchanges"] > 0


@pytest.mark.parametrize("n,world", [(7, 2), (10, 4)])
def test_signed_controls_sharded_vs_single_gpu(n, world):
    """spz_mc_apply_signed on a sharded register (X on the zero-controls around the all-ones gate, wherever those qubits live)
    against the single-GPU one-launch form: bit for bit."""
    cpu = orc.gen_random_state(n, 5 * n + world)
    ref = sb.State.from_arrays(cpu.reals, cpu.imags)
    states = DistState.create_local_group(n, world)
    upload_shards(states, cpu)
    rng = np.random.default_rng(n * world)
    seq = []
    for i in range(30):
        kind, p = GATES[int(rng.integers(len(GATES)))]
        t = int(rng.integers(n))
        others = [q for q in range(n) if q != t]
        cs = [int(c) for c in rng.choice(others, size=int(rng.integers(1, 4)), replace=False)]
        if i % 4 == 0 and t != n - 1 and n - 1 not in cs:
            cs.append(n - 1)                      # a global qubit as a zero-control
        zs = [cs[-1]] + [c for c in cs[:-1] if rng.random() < 0.4]
        seq.append((kind, p, [c for c in cs if c not in zs], zs, t))

    def body(rank, s):
        for kind, p, ones, zs, t in seq:
            sb.mc_apply_signed(G(kind, p), s, ones, zs, t)
        s.sync()
    run_group(states, body)
    for kind, p, ones, zs, t in seq:
        sb.mc_apply_signed(G(kind, p), ref, ones, zs, t)
    re, im = gather(states)
    rre, rim = ref.download()
    assert np.array_equal(re, rre) and np.array_equal(im, rim)


@pytest.mark.parametrize("n,world,fuse", [(8, 2, True), (10, 4, True), (10, 4, False), (13, 8, True), (16, 4, True)])
def test_execute_sharded_vs_dense(n, world, fuse):
    ops = random_ops(n, 150, seed=n * 7 + world, with_swap=True)
    cpu = orc.gen_random_state(n, 3 * n + world)
    states = DistState.create_local_group(n, world)
    upload_shards(states, cpu)

    def body(rank, s):
        build_circuit(n, ops, s, fuse=fuse).execute()
        s.sync()
    run_group(states, body)
    re, im = gather(states)
    want = run_dense(n, cpu.amps(), ops)
    assert np.max(np.abs((re + 1j * im) - want)) < 1e-12
    ref = sb.State.from_arrays(cpu.reals, cpu.imags)
    build_circuit(n, ops, ref, fuse=False).execute()
    rre, rim = ref.download()
    if not fuse:
        assert np.array_equal(re, rre) and np.array_equal(im, rim)
    else:
        assert np.max(np.abs(re - rre)) <= 1e-12 and np.max(np.abs(im - rim)) <= 1e-12


@pytest.mark.parametrize("n,world", [(10, 8), (14, 4)])
def test_qft_sharded_closed_form_and_reductions(n, world):
    x = 0x9E3779B97F4A7C15 % (1 << n)
    states = DistState.create_local_group(n, world)

    def body(rank, s):
        s.set_basis(x)
        qc = QuantumCircuit.from_state(s, fuse=True)
        qc.qft()
        qc.execute()
        ex = s.stats()["exchanges"]
        out = {"ex": ex, "norm": sb.norm2(s), "p0": [sb.prob0(s, t) for t in range(n)],
               "x": sb.xyz_expectation_value("x", s, [0, n - 1]), "z": sb.xyz_expectation_value("z", s, [n - 1]),
               "qev": sb.qubit_expectation_value(s, n - 2)}
        s.sync()
        return out
    res = run_group(states, body)
    assert all(r == res[0] for r in res)  # every rank sees the bitwise-identical reduction
    re, im = gather(states)
    k = np.arange(1 << n)
    rev = np.zeros_like(k)
    for b in range(n):
        rev |= ((k >> b) & 1) << (n - 1 - b)
    want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ((x * rev) % (1 << n)) / (1 << n))
    assert np.max(np.abs((re + 1j * im) - want)) < 1e-12
    assert abs(res[0]["norm"] - 1.0) < 1e-10
    cpu = orc.State(n, re, im)
    for t in range(n):
        assert abs(res[0]["p0"][t] - orc.prob0(cpu, t)) < 1e-12
    assert np.max(np.abs(np.array(res[0]["x"]) - orc.xyz_expectation_value("x", cpu, [0, n - 1]))) < 1e-12
    assert abs(res[0]["z"][0] - orc.xyz_expectation_value("z", cpu, [n - 1])[0]) < 1e-12
    # the register was a basis state when execute() started: its qubits were placed by looking ahead (dist_place_basis), QFT's
    # last g qubits start as the rank bits and only they are ever exchanged (the revolving door alone needs g + 1)
    assert res[0]["ex"] == int(math.log2(world))


@pytest.mark.parametrize("n,world", [(13, 2), (15, 8)])
def test_placement_is_only_used_while_the_register_is_a_basis_state(n, world):
    """dist_place_basis: free relabelling needs a basis state.  A gate before execute(), an upload, or a state that merely
    started as a basis state must leave the permutation alone -- and every variant must give the same amplitudes."""
    g = int(math.log2(world))
    x = 0x5DEECE66D % (1 << n)
    want = None
    for variant in ("placed", "gate-first", "uploaded", "switched-off"):
        if variant == "switched-off":
            os.environ["SPZ_DIST_PLACE"] = "0"
        try:
            states = DistState.create_local_group(n, world)
            if variant == "uploaded":
                cpu = orc.State(n)
                cpu.reals[0] = 0.0
                cpu.reals[x] = 1.0
                upload_shards(states, cpu)        # the same amplitudes, but the engine was not told it is a basis state

            def body(rank, s):
                if variant != "uploaded":
                    s.set_basis(x)
                if variant == "gate-first":
                    sb.apply(Gate.Z, s, 0)        # any gate ends the freedom (Z leaves |x> alone up to a sign)
                qc = QuantumCircuit.from_state(s, fuse=True)
                qc.qft()
                qc.execute()
                s.sync()
                return s.stats()["exchanges"], s.perm()
            res = run_group(states, body)
            assert all(r == res[0] for r in res)
            re, im = gather(states)
            got = re + 1j * im
            if variant == "gate-first" and (x & 1):
                got = -got
            if want is None:
                k = np.arange(1 << n)
                rev = np.zeros_like(k)
                for b in range(n):
                    rev |= ((k >> b) & 1) << (n - 1 - b)
                want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ((x * rev) % (1 << n)) / (1 << n))
            assert np.max(np.abs(got - want)) < 1e-12, variant
            assert res[0][0] == (g if variant == "placed" else g + 1), (variant, res[0][0])
        finally:
            os.environ.pop("SPZ_DIST_PLACE", None)


@pytest.mark.parametrize("n,world", [(6, 4), (13, 2), (15, 8)])
def test_z_expectation_of_every_qubit_sharded(n, world):
    """xyz_expectation_value('z', all qubits) on a sharded register: one read pass per shard for the local qubits, the rank's
    bit for the global ones -- also after exchanges have permuted the qubits."""
    cpu = orc.gen_random_state(n, 23 + n)
    states = DistState.create_local_group(n, world)
    upload_shards(states, cpu)

    def body(rank, s):
        first = sb.xyz_expectation_value("z", s, list(range(n)))
        sb.apply(Gate.H, s, n - 1)          # an exchange: qubit n-1 becomes local, another one global
        sb.apply(Gate.RX(0.4), s, 1)
        second = sb.xyz_expectation_value("z", s, list(reversed(range(n))))
        s.sync()
        return first, second
    res = run_group(states, body)
    assert all(r == res[0] for r in res)
    assert np.max(np.abs(np.array(res[0][0]) - orc.xyz_expectation_value("z", cpu, list(range(n))))) < 1e-12
    orc.apply(orc.H, cpu, n - 1)
    orc.apply(orc.RX, cpu, 1, (0.4,))
    assert np.max(np.abs(np.array(res[0][1]) - orc.xyz_expectation_value("z", cpu, list(reversed(range(n)))))) < 1e-12
    assert states[0].stats()["exchanges"] >= 1


@pytest.mark.parametrize("n,world", [(7, 4), (11, 2)])
def test_measure_sharded(n, world):
    cpu = orc.gen_random_state(n, 19 + n)
    states = DistState.create_local_group(n, world)
    upload_shards(states, cpu)
    plan = [(n - 1, 1, True), (0, 0, True), (n - 2, 1, False), (1, 1, True)]

    def body(rank, s):
        bits = [sb.measure_qubit(s, t, reset, v) for t, v, reset in plan]
        s.sync()
        return bits
    res = run_group(states, body)
    for t, v, reset in plan:
        orc.measure_qubit(cpu, t, reset, v)
    assert all(r == [v for _, v, _ in plan] for r in res)
    re, im = gather(states)
    assert np.max(np.abs(re - cpu.reals)) < 1e-12 and np.max(np.abs(im - cpu.imags)) < 1e-12


def test_init_random_sharded_matches_single_gpu():
    n, world = 12, 4
    states = DistState.create_local_group(n, world)
    run_group(states, lambda r, s: (s.init_random(42), s.sync()))
    re, im = gather(states)
    ref = sb.State(n); ref.init_random(42)
    rre, rim = ref.download()
    assert np.max(np.abs(re - rre)) < 1e-14 and np.max(np.abs(im - rim)) < 1e-14


def test_world_one_dist_state_behaves_like_state():
    from spinoza_b200.distributed import DistEnv
    n = 8
    cpu = orc.gen_random_state(n, 2)
    s = DistState(n, DistEnv(0, 1, 0))
    s.upload(cpu.reals, cpu.imags)
    sb.apply(Gate.H, s, n - 1); orc.apply(orc.H, cpu, n - 1)
    sb.c_apply(Gate.P(0.3), s, 0, n - 1); orc.c_apply(orc.P, cpu, 0, n - 1, (0.3,))
    re, im = s.download()
    assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags)
    assert abs(sb.prob0(s, 3) - orc.prob0(cpu, 3)) < 1e-12


@pytest.mark.skipif(sb.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_ipc_nvlink():
    env = dict(os.environ)
    env["PYTHONPATH"] = str(ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "tests" / "_dist_gpu_worker.py")],
                       capture_output=True, text=True, env=env, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "GPU_DIST_OK" in r.stdout


@pytest.mark.parametrize("n,world", [(8, 4), (12, 2)])
def test_sample_sharded(n, world):
    cpu = orc.gen_random_state(n, 77 + n)
    states = DistState.create_local_group(n, world)
    upload_shards(states, cpu)
    # move a qubit around first so that the permutation is not the identity
    run_group(states, lambda r, s: (sb.apply(Gate.H, s, n - 1), sb.apply(Gate.H, s, n - 1), s.sync()))
    shots = 1 << 15
    u = orc.uniforms(7, shots)
    outs = run_group(states, lambda r, s: sb.sample(s, shots, u01=u))
    merged = np.max(np.stack(outs), axis=0)
    owners = np.sum(np.stack(outs) >= 0, axis=0)
    assert np.all(owners == 1) and merged.min() >= 0 and merged.max() < (1 << n)   # every shot answered exactly once
    re, im = gather(states)
    p = re ** 2 + im ** 2
    counts = np.bincount(merged, minlength=1 << n)
    keep = p * shots > 5
    chi2 = np.sum((counts[keep] - shots * p[keep]) ** 2 / (shots * p[keep]))
    dof = int(np.count_nonzero(keep))
    assert chi2 < dof + 6 * math.sqrt(2 * dof)
    # basis state: exact (core.rs:272-291)
    run_group(states, lambda r, s: (s.set_basis(5), s.sync()))
    outs = run_group(states, lambda r, s: sb.sample(s, 64, seed=1))
    assert np.all(np.max(np.stack(outs), axis=0) == 5)


# ---- windows that span exchanges (the default; SPZ_DIST_WINDOW=0 closes the window at every exchange) ----------------------

@pytest.mark.parametrize("window", ["0", "1"])
@pytest.mark.parametrize("select", ["0", "1"])
@pytest.mark.parametrize("n,world", [(15, 2), (16, 4), (17, 8)])
def test_windows_spanning_exchanges_match_the_oracle(n, world, select, window, monkeypatch):
    monkeypatch.setenv("SPZ_DIST_WINDOW", window)
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
