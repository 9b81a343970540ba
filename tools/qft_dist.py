"""BASELINE config 4: QFT-n sharded over the GPUs of one box (torchrun, one process per GPU).

    torchrun --nproc-per-node 8 tools/qft_dist.py --local-qubits 33      # QFT-36, 2^33 amplitudes (137 GB) per GPU

Checks (no CPU state needed): closed form QFT|x>[k] = 2^(-n/2) exp(2 pi i x rev(k) / 2^n) on a sample of every shard,
norm, and QFT followed by IQFT returning the start state.  Prints one JSON line on rank 0.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spinoza_b200 as sb  # noqa: E402
from spinoza_b200 import QuantumCircuit  # noqa: E402
from spinoza_b200.distributed import DistState, init_from_env  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local-qubits", type=int, default=30)
    ap.add_argument("--sample", type=int, default=1 << 16)
    ap.add_argument("--unfused", action="store_true")
    args = ap.parse_args()
    env = init_from_env()
    g = env.world.bit_length() - 1
    n = args.local_qubits + g
    t0 = time.time()
    s = DistState(n, env)
    x = 0x9E3779B97F4A7C15 % (1 << n)
    s.set_basis(x)
    s.sync(); env.barrier()
    alloc_s = time.time() - t0
    qc = QuantumCircuit.from_state(s, fuse=not args.unfused)
    qc.qft()
    n_gates = len(qc.transformations)
    l0 = sb.launch_count()
    s.sync(); env.barrier()
    s.timer_start()
    qc.execute()
    ms = env.max_float(s.timer_stop())
    launches = sb.launch_count() - l0
    stats = s.stats()
    nrm = sb.norm2(s)
    # closed form on the first `sample` local amplitudes of this shard
    perm = s.perm()
    cnt = min(args.sample, len(s))
    re, im = s.download(0, cnt)
    phys = (np.uint64(env.rank) << np.uint64(s.n_local)) + np.arange(cnt, dtype=np.uint64)
    # logical index k of physical index p: bit perm[q] of p is bit q of k; rev(k) bit (n-1-q) = bit q of k
    rev = np.zeros(cnt, dtype=np.uint64)
    for q in range(n):
        rev |= ((phys >> np.uint64(perm[q])) & np.uint64(1)) << np.uint64(n - 1 - q)
    mod = (1 << n) - 1
    # (x * rev) mod 2^n with 64-bit wraparound is exact for n <= 64
    ph = ((np.uint64(x) * rev) & np.uint64(mod)).astype(np.float64) / float(1 << n)
    want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ph)
    err = float(np.max(np.abs((re + 1j * im) - want)))
    err = env.max_float(err)
    # round trip
    qc.iqft(list(reversed(range(n))))
    s.timer_start()
    qc.execute()
    ms_inv = env.max_float(s.timer_stop())
    nrm2 = sb.norm2(s)
    perm = s.perm()
    px = 0
    for q in range(n):
        if (x >> q) & 1:
            px |= 1 << perm[q]
    owner, local = px >> s.n_local, px & ((1 << s.n_local) - 1)
    back = abs(s.amp(local) - 1.0) if owner == env.rank else 0.0
    back = env.max_float(back)
    stats2 = s.stats()
    if env.rank == 0:
        print(json.dumps({
            "workload": f"QFT-{n} on {env.world} GPU(s), 2^{s.n_local} amplitudes per GPU, {'unfused' if args.unfused else 'fused'}",
            "gates": n_gates, "seconds": ms * 1e-3, "sec_per_gate": ms * 1e-3 / n_gates, "launches": int(launches),
            "exchanges": stats["exchanges"], "overlapped_exchanges": stats["overlapped"], "nvlink_GBps_per_direction": stats["nvlink_GBps_per_direction"],
            "exchange_ms_total": stats["exchange_ms"], "max_abs_err_vs_closed_form": err, "norm2": nrm,
            "iqft_seconds": ms_inv * 1e-3, "roundtrip_abs_err_at_x": back, "norm2_after_roundtrip": nrm2,
            "exchanges_incl_roundtrip": stats2["exchanges"], "alloc_seconds": alloc_s,
        }), flush=True)
    env.shutdown()


if __name__ == "__main__":
    main()
