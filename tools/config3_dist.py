"""BASELINE config 3 on a sharded register: random layered circuit (depth 20) on the GPUs of one box.

    torchrun --nproc-per-node 2 tools/config3_dist.py --local-qubits 32          # 33 qubits on 2 GPUs
    SPZ_DIST_WINDOW=1 torchrun --nproc-per-node 2 tools/config3_dist.py ...      # windows that span exchanges (opt-in)

Checks without a CPU state: norm, and circuit followed by its inverse returning the start state on a sample.  Prints one JSON
line on rank 0 with seconds, launches (passes) and exchange statistics, tagged with the scheduler switches in effect.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spinoza_b200 as sb  # noqa: E402
from spinoza_b200 import QuantumCircuit, workloads  # noqa: E402
from spinoza_b200.distributed import DistState, init_from_env  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local-qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--sample", type=int, default=1 << 16)
    args = ap.parse_args()
    env = init_from_env()
    g = env.world.bit_length() - 1
    n = args.local_qubits + g
    s = DistState(n, env)
    s.init_random(42)
    cnt = min(args.sample, len(s))
    re0, im0 = s.download(0, cnt)
    s.sync(); env.barrier()
    qc = QuantumCircuit.from_state(s, fuse=True)
    workloads.random_layered_circuit(qc, depth=args.depth, seed=42)
    n_gates = len(qc.transformations)
    inverse = QuantumCircuit.from_state(s, fuse=True)
    workloads.random_layered_circuit(inverse, depth=args.depth, seed=42)
    inverse.inverse()
    l0 = sb.launch_count()
    st0 = s.stats()
    s.timer_start()
    qc.execute()
    ms = env.max_float(s.timer_stop())
    launches = sb.launch_count() - l0
    st1 = s.stats()
    nrm = sb.norm2(s)
    inverse.execute()
    s.sync(); env.barrier()
    # the inverse ends with the permutation it ends with: compare through the logical view only if it is the identity again
    perm = s.perm()
    back = None
    if perm == list(range(n)):
        re1, im1 = s.download(0, cnt)
        back = env.max_float(float(np.max(np.abs((re1 - re0) + 1j * (im1 - im0)))))
    nrm_inv = sb.norm2(s)  # a collective: every rank
    if env.rank == 0:
        print(json.dumps({
            "workload": f"random layered circuit, {n} qubits, depth {args.depth}, {n_gates} gates, {env.world} GPU(s), "
                        f"2^{s.n_local} amplitudes per GPU",
            "seconds": ms * 1e-3, "sec_per_gate": ms * 1e-3 / n_gates, "launches": int(launches),
            "exchanges": st1["exchanges"] - st0["exchanges"], "exchange_ms": st1["exchange_ms"] - st0["exchange_ms"],
            "norm2": nrm, "norm2_after_inverse": nrm_inv, "roundtrip_max_abs_err_first_sample": back,
            "switches": {k: os.environ.get(k) for k in ("SPZ_DIST_WINDOW", "SPZ_TILE_SELECT", "SPZ_TILE_V3", "SPZ_TILE_LMIN", "SPZ_NO_OVERLAP")},
        }))
    env.shutdown()


if __name__ == "__main__":
    main()
