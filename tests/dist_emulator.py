"""CPU emulation of the sharded execution: the REAL planner (spz_dist_plan_* in libspinoza_b200.so, pure host
code) drives NumPy shards.  Local gates use the independent dense statement in tests/_dense.py; the exchange is
either an in-process swap (any world size) or gloo send/recv between real processes (world_size 2 test)."""
import ctypes as C

import numpy as np

import spinoza_b200 as sb
from tests import _dense as D

ACT_SKIP, ACT_LOCAL_GATE, ACT_DIAG_CONST, ACT_EXCHANGE = 0, 1, 2, 3


class Plan:
    def __init__(self, n, world):
        self.n, self.world = n, world
        self.h = C.c_void_p()
        sb._check(sb._lib.spz_dist_plan_create(n, world, C.byref(self.h)))

    def __del__(self):
        sb._lib.spz_dist_plan_destroy(self.h)

    def lower(self, rank, op):
        out = (sb._DistAction * 8)()
        cnt = C.c_int()
        sb._check(sb._lib.spz_dist_plan_lower(self.h, rank, C.byref(op), out, 8, C.byref(cnt)))
        return [out[i] for i in range(cnt.value)]

    def perm(self):
        out = (C.c_int32 * self.n)()
        sb._check(sb._lib.spz_dist_plan_perm(self.h, out))
        return list(out)


def make_op(kind, target=0, params=(), ctrl_mask=0, t0=0, t1=0):
    op = sb._Op()
    op.kind, op.target, op.t0, op.t1 = kind, target, t0, t1
    for i, v in enumerate(params[:3]):
        op.p[i] = v
    nc = bin(ctrl_mask).count("1")
    op.ctrl_kind = 0 if nc == 0 else 1 if nc == 1 else 3
    op.ctrl_mask = ctrl_mask
    return op


def apply_local_action(psi, n_local, a):
    """psi: complex shard.  Returns the new shard for LOCAL_GATE / DIAG_CONST actions."""
    p = tuple(a.p)
    if a.type == ACT_LOCAL_GATE:
        return D.apply_matrix(psi, n_local, D.matrix(a.kind, p), a.target, a.cmask)
    if a.type == ACT_DIAG_CONST:
        m = D.matrix(a.kind, p)
        f = m[1, 1] if a.hi else m[0, 0]
        idx = np.arange(1 << n_local)
        sel = (idx & a.cmask) == a.cmask
        out = psi.copy()
        out[sel] *= f
        return out
    return psi


def exchange_halves(lo_shard, hi_shard, lq):
    """In place: low rank's [lq = 1] half <-> high rank's [lq = 0] half."""
    idx = np.arange(len(lo_shard))
    one = idx[(idx >> lq) & 1 == 1]
    zero = one ^ (1 << lq)
    tmp = lo_shard[one].copy()
    lo_shard[one] = hi_shard[zero]
    hi_shard[zero] = tmp


def run_sharded_inprocess(n, world, psi0, ops):
    """ops: list of spz_op.  Returns the final state in logical order."""
    g = world.bit_length() - 1
    n_local = n - g
    shards = [psi0[r << n_local:(r + 1) << n_local].copy() for r in range(world)]
    plans = [Plan(n, world) for _ in range(world)]
    n_exchanges = 0
    for op in ops:
        acts = [plans[r].lower(r, op) for r in range(world)]
        assert len({len(a) for a in acts}) == 1, "ranks disagree on the number of actions"
        for step in range(len(acts[0])):
            types = {acts[r][step].type == ACT_EXCHANGE for r in range(world)}
            assert len(types) == 1, "ranks disagree on where the exchange is"
            if acts[0][step].type == ACT_EXCHANGE:
                n_exchanges += 1
                for r in range(world):
                    a = acts[r][step]
                    assert a.partner == r ^ (1 << a.gbit)
                    if not (r >> a.gbit) & 1:
                        exchange_halves(shards[r], shards[a.partner], a.lq)
            else:
                for r in range(world):
                    shards[r] = apply_local_action(shards[r], n_local, acts[r][step])
    perms = [p.perm() for p in plans]
    assert all(p == perms[0] for p in perms), "permutation diverged between ranks"
    phys = np.concatenate(shards)
    re, im = sb.distributed.unpermute(phys.real, phys.imag, perms[0])
    return re + 1j * im, n_exchanges, perms[0]
