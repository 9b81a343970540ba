"""Runs one fused QFT-n (and optionally the random layered circuit) for ncu captures of the tile kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
s = sb.State(n)
s.set_basis(12345 % (1 << n))
for rep in range(2):
    qc = QuantumCircuit.from_state(s, fuse=True)
    qc.qft()
    s.timer_start()
    qc.execute()
    print(f"fused QFT-{n}: {s.timer_stop():.2f} ms, launches so far {sb.launch_count()}")
