"""One sampling call for ncu launch lists: python tools/profile_sample.py [n=30] [uniform|random|basis] [shots=2^20]"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
kind = sys.argv[2] if len(sys.argv) > 2 else "random"
shots = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
s = sb.State(n)
if kind == "random":
    s.init_random(42)
elif kind == "uniform":
    qc = QuantumCircuit.from_state(s, fuse=True)
    for q in range(n):
        qc.h(q)
    qc.execute()
else:
    s.set_basis(12345)
s.sync()
u = np.random.default_rng(1).random(shots)
for rep in range(3):
    t0 = time.perf_counter()
    idx = sb.sample(s, shots, u01=u)
    t1 = time.perf_counter()
    print(f"sample {kind} n={n} shots={shots}: {1e3 * (t1 - t0):.2f} ms wall (uniforms given), distinct outcomes {len(np.unique(idx))}")
t0 = time.perf_counter()
idx = sb.sample(s, shots, seed=42)
print(f"with host-generated uniforms: {1e3 * (time.perf_counter() - t0):.2f} ms")
