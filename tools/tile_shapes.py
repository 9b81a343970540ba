"""How fast does a fused pass MOVE its tile?  Minimal programs (one H per listed qubit, one pass each) over tile shapes: the
low 12 qubits only, a run of six high qubits at several positions, scattered high qubits.  Prints ms per pass and the
effective GB/s (32 * 2^n bytes per pass); nothing here is a bench value.

    python tools/tile_shapes.py [n=30] [reps=5]
"""
import json
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
shapes = {
    "low 0..5 (one 32 KB box per array)": [0, 1, 2, 3, 4, 5],
    "low 6..11": [6, 7, 8, 9, 10, 11],
    "high 12..17": list(range(12, 18)),
    "high 18..23": list(range(18, 24)),
    f"high {n-6}..{n-1}": list(range(n - 6, n)),
    f"high {n-3}..{n-1} (L=9)": list(range(n - 3, n)),
    f"high {n-1} (L=11)": [n - 1],
    "two runs 14..16, 24..26": [14, 15, 16, 24, 25, 26],
    "scattered 13,16,19,22,25,28": [13, 16, 19, 22, 25, 28],
}
s = sb.State(n)
s.init_random(1)
out = {"n": n, "rows": {}}
for name, qs in shapes.items():
    ts = []
    for r in range(reps + 1):
        qc = QuantumCircuit.from_state(s, fuse=True)
        for q in qs:
            qc.h(q)
        passes = qc.plan()[1]
        l0 = sb.launch_count()
        s.timer_start()
        qc.execute()
        ms = s.timer_stop()
        if r:
            ts.append(ms)
    med = statistics.median(ts)
    out["rows"][name] = {"ms": med, "min_ms": min(ts), "passes": passes, "launches": sb.launch_count() - l0,
                         "GBps": 32.0 * (1 << n) / (med * 1e-3) / 1e9}
    print(f"{name:42s} {med:7.2f} ms  ({min(ts):6.2f} min)  passes {passes}  {out['rows'][name]['GBps']:7.0f} GB/s", file=sys.stderr)
print(json.dumps(out))
