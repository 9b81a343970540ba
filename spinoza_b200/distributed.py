"""Multi-GPU host plumbing: one process per GPU (torchrun), amplitudes sharded by the top log2(world) index bits.

`torch.distributed` (gloo) is used only to all-gather the CUDA IPC handles of each rank's shard, for
barriers and for max-over-ranks timing.  The data path is NOT a torch / NCCL collective: global-qubit gates
are pairwise half-shard exchanges done by our own kernels through IPC-mapped peer pointers over NVLink
(csrc/dist.cu), and scalar reductions go through the same peer-mapped control blocks.

Every rank must issue the same sequence of calls on its `DistState` (SPMD), exactly like the single-GPU API:
`sb.apply(Gate.H, state, t)`, `QuantumCircuit.from_state(state).execute()`, `sb.prob0(state, t)`, ...
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np

from . import State, _check, _lib, _vp, _dp

IPC_BLOB_BYTES = 256


class DistEnv:
    def __init__(self, rank: int, world: int, local_rank: int, pg=None):
        self.rank, self.world, self.local_rank, self._pg = rank, world, local_rank, pg

    def barrier(self):
        if self.world > 1:
            import torch.distributed as td
            td.barrier()

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        if self.world == 1:
            return [b]
        import torch.distributed as td
        out = [None] * self.world
        td.all_gather_object(out, b)
        return out

    def max_float(self, x: float) -> float:
        if self.world == 1:
            return x
        import torch
        import torch.distributed as td
        t = torch.tensor([x], dtype=torch.float64)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t[0])

    def gather_arrays(self, a: np.ndarray) -> Optional[List[np.ndarray]]:
        """Gather one array per rank onto rank 0 (tests / small registers only)."""
        if self.world == 1:
            return [a]
        import torch.distributed as td
        out = [None] * self.world if self.rank == 0 else None
        td.gather_object(a, out, dst=0)
        return out

    def shutdown(self):
        if self.world > 1:
            import torch.distributed as td
            if td.is_initialized():
                td.barrier()
                td.destroy_process_group()


def init_from_env(backend: str = "gloo") -> DistEnv:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1:
        import torch.distributed as td
        if not td.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            td.init_process_group(backend=backend, rank=rank, world_size=world)
    return DistEnv(rank, world, local_rank)


class DistState(State):
    """One shard of an n-qubit register spread over `env.world` GPUs.  `len(state)` is the shard length;
    `state.n` is the total qubit count."""

    def __init__(self, n: int, env: DistEnv, device: Optional[int] = None, _connect: bool = True):
        self.env = env
        dev = env.local_rank if device is None else device
        h = _vp()
        _check(_lib.spz_dist_create(int(n), env.rank, env.world, dev, C.byref(h)))
        super().__init__(n, dev, _handle=h)
        self.n_local = _lib.spz_dist_local_qubits(self._h)
        self._local_group = not _connect
        if _connect:
            blob = C.create_string_buffer(IPC_BLOB_BYTES)
            _check(_lib.spz_dist_export(self._h, blob))
            blobs = env.all_gather_bytes(blob.raw)
            _check(_lib.spz_dist_connect(self._h, b"".join(blobs)))
            env.barrier()

    @staticmethod
    def create_local_group(n: int, world: int, devices: Optional[List[int]] = None) -> List["DistState"]:
        """All `world` shards inside THIS process (drive each from its own host thread): shard r lives on
        devices[r] (default: all on device 0).  Peers are plain device pointers instead of IPC mappings; the
        kernels, flags and exchange protocol are the same as in the one-process-per-GPU deployment."""
        devices = devices or [0] * world
        states = [DistState(n, DistEnv(r, world, devices[r]), device=devices[r], _connect=False) for r in range(world)]
        arr = (_vp * world)(*[s._h for s in states])
        _check(_lib.spz_dist_connect_local(arr, world))
        return states

    def perm(self) -> List[int]:
        """logical qubit -> physical index bit (bits >= n_local are the rank's bits)."""
        out = (C.c_int32 * self.n)()
        _check(_lib.spz_dist_perm(self._h, out))
        return list(out)

    def stats(self) -> dict:
        out = (C.c_double * 4)()
        _check(_lib.spz_dist_stats(self._h, out))
        ex, sent, ms = out[0], out[1], out[2]
        return {"exchanges": int(ex), "overlapped": int(out[3]), "bytes_sent": sent, "exchange_ms": ms,
                "nvlink_GBps_per_direction": (sent / (ms * 1e-3) / 1e9) if ms > 0 else None}

    def clone(self) -> "DistState":
        """#[derive(Clone)] (core.rs:18) for a sharded register.  COLLECTIVE: every rank calls it at the same point (the new
        register exchanges IPC handles like the first one did).  Shards created with create_local_group are cloned as a
        group with clone_local_group."""
        if self._local_group:
            raise RuntimeError("shards of a local group are cloned together: DistState.clone_local_group(states)")
        new = DistState(self.n, self.env, device=self.device)
        _check(_lib.spz_dist_copy_from(new._h, self._h))
        return new

    @staticmethod
    def clone_local_group(states: List["DistState"]) -> List["DistState"]:
        new = DistState.create_local_group(states[0].n, len(states), [s.device for s in states])
        for d, s in zip(new, states):
            _check(_lib.spz_dist_copy_from(d._h, s._h))
        return new

    def sample(self, shots: int, seed: int = 0, u01: Optional[np.ndarray] = None) -> np.ndarray:
        """Exact sampling of the whole sharded register: every rank passes the same uniforms, the owning rank answers
        each shot, and the answers are combined with a max all-reduce (gloo)."""
        from . import sample as _sample
        local = _sample(self, shots, seed=seed, u01=u01)
        if self.env.world == 1:
            return local
        import torch
        import torch.distributed as td
        t = torch.from_numpy(local.copy())
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return t.numpy()

    def gather_logical(self):
        """Rank 0: the full state in LOGICAL index order (re, im); other ranks: None.  Small registers only."""
        re, im = self.download()
        perm = self.perm()
        parts = self.env.gather_arrays(np.stack([re, im]))
        if self.env.rank != 0:
            return None
        phys = np.concatenate(parts, axis=1)  # physical index = rank * 2^n_local + local index
        idx = np.arange(1 << self.n, dtype=np.int64)
        p = np.zeros_like(idx)
        for q in range(self.n):
            p |= ((idx >> q) & 1) << perm[q]
        return phys[0][p], phys[1][p]


def unpermute(phys_re: np.ndarray, phys_im: np.ndarray, perm: List[int]):
    """Map a full physical-order state to logical order given perm[logical] = physical bit."""
    n = len(perm)
    idx = np.arange(1 << n, dtype=np.int64)
    p = np.zeros_like(idx)
    for q in range(n):
        p |= ((idx >> q) & 1) << perm[q]
    return phys_re[p], phys_im[p]
