"""Sharding logic on CPU: the planner that the multi-GPU layer executes (dist_plan.h via the C ABI) must turn
any circuit into per-rank actions whose effect equals the unsharded circuit -- checked in-process for 2/4/8
logical ranks and with two real processes exchanging half shards over gloo."""
import math
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests import _dense as D
from tests.dist_emulator import ACT_EXCHANGE, Plan, make_op, run_sharded_inprocess

ROOT = Path(__file__).resolve().parent.parent
KINDS = [D.H, D.X, D.Y, D.Z, D.P, D.RX, D.RY, D.RZ, D.U]


def random_ops(n, count, seed):
    rng = np.random.default_rng(seed)
    ops, dense = [], []
    for _ in range(count):
        r = rng.random()
        if r < 0.08:
            a, b = int(rng.integers(n)), int(rng.integers(n))
            ops.append(make_op(9, t0=a, t1=b)); dense.append(("s", a, b))
            continue
        kind = KINDS[int(rng.integers(len(KINDS)))]
        p = tuple(float(x) for x in rng.random(3) * 2 * math.pi)
        t = int(rng.integers(n))
        cm = 0
        if r > 0.5:
            k = int(rng.integers(1, min(3, n - 1) + 1))
            for c in rng.choice([q for q in range(n) if q != t], size=k, replace=False):
                cm |= 1 << int(c)
        ops.append(make_op(kind, t, p, cm)); dense.append(("g", kind, p, t, cm))
    return ops, dense


def run_dense(n, psi, dense):
    for o in dense:
        psi = D.apply_swap(psi, n, o[1], o[2]) if o[0] == "s" else D.apply_matrix(psi, n, D.matrix(o[1], o[2]), o[3], o[4])
    return psi


@pytest.mark.parametrize("n,world", [(4, 2), (6, 2), (6, 4), (7, 8), (10, 4), (12, 8)])
def test_sharded_equals_unsharded(n, world):
    psi0 = D.random_state(n, 10 * n + world)
    ops, dense = random_ops(n, 120, seed=n * 31 + world)
    got, n_ex, perm = run_sharded_inprocess(n, world, psi0, ops)
    want = run_dense(n, psi0, dense)
    assert np.max(np.abs(got - want)) < 1e-12
    assert n_ex > 0 and sorted(perm) == list(range(n))


def test_diagonal_gates_never_communicate():
    n, world = 8, 4
    plan = Plan(n, world)
    for kind in (D.Z, D.P, D.RZ):
        for t in range(n):
            for c in range(n):
                if c == t:
                    continue
                for r in range(world):
                    acts = plan.lower(r, make_op(kind, t, (0.3,), 1 << c))
                    assert all(a.type != ACT_EXCHANGE for a in acts)
    assert plan.perm() == list(range(n))


def test_qft_exchange_count_and_lookahead_victims():
    # QFT-n on 8 ranks (g = 3): each qubit gets exactly one non-diagonal gate (its H); afterwards it is only ever
    # a CP control/target (diagonal, free).  So a qubit that has had its H is the ideal victim for the next
    # global H ("revolving door"): g exchanges for the g global qubits + 1 to bring the first victim back = g + 1.
    n, world = 12, 8
    plan = Plan(n, world)
    ex = 0
    for j in range(n):
        for k in range(j):
            acts = plan.lower(0, make_op(D.P, n - 1 - k, (math.pi / 2 ** (j - k),), 1 << (n - 1 - j)))
            ex += sum(a.type == ACT_EXCHANGE for a in acts)
        acts = plan.lower(0, make_op(D.H, n - 1 - j))
        ex += sum(a.type == ACT_EXCHANGE for a in acts)
    assert ex == 4


def test_swap_is_a_relabel():
    plan = Plan(6, 4)
    assert plan.lower(1, make_op(9, t0=0, t1=5)) == []
    perm = plan.perm()
    assert perm[0] == 5 and perm[5] == 0


def test_two_process_gloo_exchange():
    """world_size 2, real processes: half shards travel through torch.distributed (gloo) send/recv."""
    script = ROOT / "tests" / "_dist_gloo_worker.py"
    env = dict(os.environ)
    env["PYTHONPATH"] = str(ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_DIST_OK" in r.stdout
