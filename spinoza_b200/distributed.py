"""Multi-GPU host plumbing: one process per GPU (torchrun or any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK),
amplitudes sharded by the top log2(world) index bits.

Nothing here imports torch.  The host-side control plane -- all-gathering the CUDA IPC handles of each rank's shard, barriers,
max-over-ranks of a timing -- is the library's own file rendezvous (`spz_rdv_*`, csrc/rendezvous.cu: small files in
/dev/shm), the same four C functions a C++ or Rust caller would use.  The data path is no collective library either:
global-qubit gates are pairwise half-shard exchanges done by our own kernels through IPC-mapped peer pointers over NVLink
(csrc/dist.cu), and scalar reductions go through the same peer-mapped control blocks.

Every rank must issue the same sequence of calls on its `DistState` (SPMD), exactly like the single-GPU API:
`sb.apply(Gate.H, state, t)`, `QuantumCircuit.from_state(state).execute()`, `sb.prob0(state, t)`, ...
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from typing import List, Optional

import numpy as np

from . import State, _check, _lib, _vp, _dp

IPC_BLOB_BYTES = 256


class DistEnv:
    """Rank / world of this process and the rendezvous between the processes of the node."""

    def __init__(self, rank: int, world: int, local_rank: int, rdv_dir: Optional[str] = None, _open: bool = True):
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self._rdv = None
        if world > 1 and _open:
            h = _vp()
            _check(_lib.spz_rdv_open(rdv_dir.encode() if rdv_dir else None, rank, world, C.byref(h)))
            self._rdv = h

    def barrier(self):
        if self._rdv is not None:
            _check(_lib.spz_rdv_barrier(self._rdv))

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        """Equal-length blobs, one per rank, in rank order."""
        if self._rdv is None:
            return [b]
        out = C.create_string_buffer(len(b) * self.world)
        _check(_lib.spz_rdv_allgather(self._rdv, b, len(b), out))
        return [out.raw[i * len(b):(i + 1) * len(b)] for i in range(self.world)]

    def max_float(self, x: float) -> float:
        if self._rdv is None:
            return x
        return max(struct.unpack("<d", p)[0] for p in self.all_gather_bytes(struct.pack("<d", float(x))))

    def all_gather_array(self, a: np.ndarray) -> List[np.ndarray]:
        """Same shape and dtype on every rank."""
        a = np.ascontiguousarray(a)
        return [np.frombuffer(p, dtype=a.dtype).reshape(a.shape) for p in self.all_gather_bytes(a.tobytes())]

    def gather_arrays(self, a: np.ndarray) -> Optional[List[np.ndarray]]:
        """One array per rank onto rank 0 (tests / small registers only; every rank takes part)."""
        parts = self.all_gather_array(a)
        return parts if self.rank == 0 else None

    def shutdown(self):
        if self._rdv is not None:
            _check(_lib.spz_rdv_close(self._rdv))
            self._rdv = None


def init_from_env(backend: str = "files", rdv_dir: Optional[str] = None) -> DistEnv:
    """RANK / WORLD_SIZE / LOCAL_RANK as torchrun sets them.  `backend` is kept for callers of the first version (which
    rendezvoused through torch.distributed): there is one backend now, the library's own."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    return DistEnv(rank, world, local_rank, rdv_dir or os.environ.get("SPZ_RDV_DIR"))


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
            if env._rdv is not None:
                _check(_lib.spz_dist_connect_rdv(self._h, env._rdv))  # export + all-gather + connect + barrier, in C
            else:
                blob = C.create_string_buffer(IPC_BLOB_BYTES)
                _check(_lib.spz_dist_export(self._h, blob))
                _check(_lib.spz_dist_connect(self._h, blob.raw))

    @staticmethod
    def create_local_group(n: int, world: int, devices: Optional[List[int]] = None) -> List["DistState"]:
        """All `world` shards inside THIS process (drive each from its own host thread): shard r lives on
        devices[r] (default: all on device 0).  Peers are plain device pointers instead of IPC mappings; the
        kernels, flags and exchange protocol are the same as in the one-process-per-GPU deployment."""
        devices = devices or [0] * world
        states = [DistState(n, DistEnv(r, world, devices[r], _open=False), device=devices[r], _connect=False) for r in range(world)]
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
        each shot (the others report -1), and the answers are combined with an element-wise max over ranks."""
        from . import sample as _sample
        local = _sample(self, shots, seed=seed, u01=u01)
        if self.env.world == 1 or self.env._rdv is None:
            return local
        return np.maximum.reduce(self.env.all_gather_array(local))

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
