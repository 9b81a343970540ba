"""Worker for tests/test_dist_plan.py::test_two_process_gloo_exchange (run under torchrun, 2 ranks, CPU only)."""
import math
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as td

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from tests import _dense as D  # noqa: E402
from tests.dist_emulator import ACT_EXCHANGE, Plan, apply_local_action  # noqa: E402
from tests.test_dist_plan import random_ops, run_dense  # noqa: E402
import spinoza_b200 as sb  # noqa: E402


def main():
    td.init_process_group(backend="gloo")
    rank, world = td.get_rank(), td.get_world_size()
    n = 9
    n_local = n - 1
    psi0 = D.random_state(n, 77)
    ops, dense = random_ops(n, 150, seed=5)
    shard = psi0[rank << n_local:(rank + 1) << n_local].copy()
    plan = Plan(n, world)
    n_ex = 0
    for op in ops:
        for a in plan.lower(rank, op):
            if a.type == ACT_EXCHANGE:
                n_ex += 1
                my_bit = (rank >> a.gbit) & 1
                idx = np.arange(len(shard))
                mine = idx[((idx >> a.lq) & 1) == (0 if my_bit else 1)]  # low rank gives its lq=1 half
                send = torch.from_numpy(np.ascontiguousarray(shard[mine]).view(np.float64).copy())
                recv = torch.empty_like(send)
                reqs = [td.isend(send, a.partner), td.irecv(recv, a.partner)]
                for r in reqs:
                    r.wait()
                shard[mine] = recv.numpy().view(np.complex128)
            else:
                shard = apply_local_action(shard, n_local, a)
    parts = [None] * world
    td.all_gather_object(parts, shard)
    if rank == 0:
        phys = np.concatenate(parts)
        re, im = sb.distributed.unpermute(phys.real, phys.imag, plan.perm())
        want = run_dense(n, psi0, dense)
        err = float(np.max(np.abs((re + 1j * im) - want)))
        assert err < 1e-12, err
        assert n_ex > 0
        print(f"GLOO_DIST_OK exchanges={n_ex} err={err:.2e}")
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
