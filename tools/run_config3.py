"""BASELINE config 3: random layered circuit (1q rotations + CNOT / CP entanglers, depth 20) on one GPU.
    python tools/run_config3.py 30 [depth]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spinoza_b200 as sb  # noqa: E402
from spinoza_b200 import QuantumCircuit, workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
s = sb.State(n)
out = {"workload": f"random layered circuit, n={n}, depth={depth}, seed 42"}
for label, kw in (("fused", dict(fuse=True)), ("fused_keep_order", dict(fuse=True, reorder=False)),
                  ("fused_exact", dict(fuse=True, exact=True)), ("unfused", dict(fuse=False))):
    s.init_random(42)
    qc = QuantumCircuit.from_state(s, **kw)
    gates = workloads.random_layered_circuit(qc, depth=depth, seed=42)
    l0 = sb.launch_count()
    s.sync()
    s.timer_start()
    qc.execute()
    ms = s.timer_stop()
    out[label] = {"seconds": ms * 1e-3, "sec_per_gate": ms * 1e-3 / gates, "launches": int(sb.launch_count() - l0),
                  "effective_GBps": gates * 32.0 * (1 << n) / (ms * 1e-3) / 1e9, "norm2": sb.norm2(s)}
out["gates"] = gates
print(json.dumps(out))
