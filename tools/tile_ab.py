"""A/B timing of the fused tile kernels on one GPU: k_tile3 (default: TMA, lowered program, rescaled butterflies) against k_tile
(SPZ_TILE_V3=0), with the scheduler knobs, on QFT-n and on the random layered circuit of BASELINE config 3.

    python tools/tile_ab.py [n=30] [reps=5] > gpurun_out/tile_ab.json

For each variant: correctness first (max |amp - k_tile amp| on QFT of a basis state, and the closed form), then the median of
`reps` CUDA-event timings per workload.  Prints one JSON object; nothing here is a bench value (bench.py is).
"""
import json
import math
import os
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
VARIANTS = [("k_tile", {"SPZ_TILE_V3": "0"}),
            ("k_tile3", {}),
            # the round-1 scheduler (first ready ops claim the tile)
            ("k_tile3/first-come-tile", {"SPZ_TILE_SELECT": "0"}),
            # shorter tile segments: more arbitrary high qubits per pass, smaller TMA boxes
            ("k_tile3/lmin5", {"SPZ_TILE_LMIN": "5"}), ("k_tile3/lmin4", {"SPZ_TILE_LMIN": "4"}),
            # groups of fewer than k ops as k-1 roofline passes instead of one tile pass
            ("k_tile3/minops3", {"SPZ_TILE_MIN_OPS": "3"})]
if len(sys.argv) > 3:
    VARIANTS = [v for v in VARIANTS if v[0] in sys.argv[3].split(",")]


def set_env(env):
    for k in ("SPZ_TILE_V3", "SPZ_TILE_LMIN", "SPZ_TILE_SELECT", "SPZ_TILE_MIN_OPS"):
        os.environ.pop(k, None)
    os.environ.update(env)


def timed(state, build):
    qc = QuantumCircuit.from_state(state, fuse=True)
    build(qc)
    state.timer_start()
    qc.execute()
    return state.timer_stop()


def qft(qc):
    qc.qft()


def layered(qc):
    workloads.random_layered_circuit(qc, depth=20, seed=42)


out = {"n": n, "reps": reps, "device": sb.device_name() if hasattr(sb, "device_name") else "", "variants": {}}
x = 0x9E3779B97F4A7C15 % (1 << n)
probe = np.array([0, 1, 2, 3, 12345 % (1 << n), (1 << n) - 1], dtype=np.int64)


def rev(k):
    return int(format(int(k), f"0{n}b")[::-1], 2)


closed = np.array([2.0 ** (-n / 2) * np.exp(2j * math.pi * ((x * rev(k)) % (1 << n)) / (1 << n)) for k in probe])
ref_amps = None
for name, env in VARIANTS:
    set_env(env)
    s = sb.State(n)
    s.set_basis(x)
    timed(s, qft)                                     # warm-up + correctness state
    amps = np.array([s.amp(int(k)) for k in probe]) if hasattr(s, "amp") else None
    rec = {"env": env}
    if amps is not None:
        rec["qft_closed_form_err"] = float(np.max(np.abs(amps - closed)))
        if ref_amps is None:
            ref_amps = amps
        rec["qft_vs_k_tile"] = float(np.max(np.abs(amps - ref_amps)))
    rec["norm2_after_qft"] = sb.norm2(s)
    for wname, build in (("qft", qft), ("layered_d20", layered)):
        s.init_random(42) if hasattr(s, "init_random") else None
        timed(s, build)
        ts = [timed(s, build) for _ in range(reps)]
        rec[wname + "_ms"] = {"median": statistics.median(ts), "min": min(ts), "max": max(ts)}
    rec["launches"] = sb.launch_count()
    out["variants"][name] = rec
    del s
print(json.dumps(out, indent=1))
