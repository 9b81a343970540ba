"""One fused run of the config-3 circuit for ncu captures.  python tools/profile_config3.py [n] [depth]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
s = sb.State(n)
s.init_random(42)
qc = QuantumCircuit.from_state(s, fuse=True)
g = workloads.random_layered_circuit(qc, depth=depth, seed=42)
s.timer_start(); qc.execute(); print(f"{g} gates fused: {s.timer_stop():.1f} ms, launches {sb.launch_count()}")
