"""The one sharded-register feature that is off by default in the product: the exchange fused with the gate that asked for it
(SPZ_DIST_FUSE_GATE=1, kernels_xgate.cuh).  It is correct on hardware -- these tests pass on a B200, and bench.py's sharded
parity extra passes with it on two GPUs over NVLink -- but slower than exchange + gate: its peer traffic is loads only and
reaches 385 GB/s per direction against 660 GB/s for the load + store exchange (profiles/round2_summary.md).  The kernel ships
in the library, so its tests run with the rest of the GPU suite (SPZ_TEST_DIST_OPTIN=0 skips them).  Local groups: every
shard on the one visible GPU, plain device pointers instead of IPC mappings."""
import os

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit, workloads
from tests.test_gpu_parity import oracle_ops_from

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SPZ_TEST_DIST_OPTIN") == "0", reason="SPZ_TEST_DIST_OPTIN=0")]


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
