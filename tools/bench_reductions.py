"""Times the measurement / probability path at n qubits: prob0, norm2, <X>,<Y>,<Z>, measure_qubit (prob0 + collapse),
sampling.  Effective GB/s = algorithmic bytes / time (bytes listed per row).   python tools/bench_reductions.py [n]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spinoza_b200 as sb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
s = sb.State(n)
s.init_random(42)
N = 1 << n
rows = {}


def timed(name, fn, nbytes, reps=5):
    fn()
    s.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    s.sync()
    dt = (time.perf_counter() - t0) / reps
    rows[name] = {"ms": dt * 1e3, "algorithmic_bytes": nbytes, "GBps": nbytes / dt / 1e9}


for t in (0, n // 2, n - 1):
    timed(f"prob0(t={t})", lambda t=t: sb.prob0(s, t), 8 * N)           # re+im of the target-bit-0 half
timed("norm2", lambda: sb.norm2(s), 16 * N)
for obs in "xyz":
    timed(f"expect_{obs}(t={n // 2})", lambda obs=obs: sb.xyz_expectation_value(obs, s, [n // 2]), 16 * N)
timed("qubit_expectation_value", lambda: sb.qubit_expectation_value(s, 3), 8 * N)


def meas():
    sb.measure_qubit(s, 5, False, 0)


s.init_random(42)
timed("measure_qubit(prob0 + collapse)", meas, 8 * N + 32 * N, reps=3)
s.init_random(42)
u = np.random.default_rng(42).random(1 << 20)
timed("sample(2^20 shots)", lambda: sb.sample(s, len(u), u01=u), 2 * 16 * N, reps=2)
print(json.dumps({"n": n, "rows": rows}))
