"""BASELINE config 5 at its stated size: qasm/quantum_lstm.qasm and qasm/iqft.qasm (4-qubit programs, tests/golden/) tiled over
disjoint 4-qubit blocks of an n-qubit register (default 32: 69 GB), the multi-controlled layer (mc X / P / RX / RY with 2-3
controls inside and outside the reference's safe domain, gates.rs:290-320) and 2^20-shot sampling, all timed.

    python tools/config5.py [n=32] [shots=2^20] > gpurun_out/config5_n32.json

Checks that need no CPU copy of the state: norm; every 4-qubit block the mc layer does not touch is an independent product
factor, so its marginal over the samples must match a 4-qubit run of the same program (max |freq - p| reported, with the
3-sigma statistical bound next to it); sampled indices in range.  Prints one JSON object.
"""
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import spinoza_b200 as sb  # noqa: E402
from spinoza_b200 import QuantumCircuit, openqasm, workloads  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
shots = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
out = {"qubits": n, "shots": shots, "state_GB": 16 * (1 << n) / 1e9, "programs": {}}
touched = {0, 1, 2, 3, n - 1, n - 2, n // 2}          # qubits the mc layer uses (workloads.multi_controlled_layer)
clean_blocks = [b for b in range(n // 4) if not (set(range(4 * b, 4 * b + 4)) & touched)]
for name in ("quantum_lstm", "iqft"):
    text = (GOLDEN / f"{name}.qasm").read_text()
    s = sb.State(n)
    rec = {}
    for label, fuse in (("fused", True), ("unfused", False)):
        s.reset_zero() if hasattr(s, "reset_zero") else s.set_basis(0)
        qc = QuantumCircuit.from_state(s, fuse=fuse)
        gates = workloads.tiled_qasm(qc, text)
        gates += workloads.multi_controlled_layer(qc)
        passes = qc.plan()[1] if fuse else gates
        l0 = sb.launch_count()
        s.timer_start()
        qc.execute()
        ms = s.timer_stop()
        rec[label] = {"seconds": ms * 1e-3, "gates": gates, "sec_per_gate": ms * 1e-3 / gates, "passes": passes,
                      "launches": int(sb.launch_count() - l0), "norm2": sb.norm2(s)}
    fused_state_norm = rec["unfused"]["norm2"]
    # sampling (the state left by the unfused run; the fused one is the same state within 1e-12)
    s.timer_start()
    idx = sb.sample(s, shots, seed=42)
    ms = s.timer_stop()
    rec["sample"] = {"shots": shots, "ms": ms, "in_range": bool(idx.min() >= 0 and idx.max() < (1 << n))}
    small = sb.State(4)
    q4 = openqasm.loads(text, fuse=False)
    q4.state = small
    q4.execute()
    p_small = np.abs(small.amps()) ** 2
    worst = 0.0
    for b in clean_blocks:
        freq = np.bincount((idx >> (4 * b)) & 0xF, minlength=16) / len(idx)
        worst = max(worst, float(np.max(np.abs(freq - p_small))))
    rec["marginals"] = {"blocks_checked": len(clean_blocks), "max_abs_freq_minus_p": worst,
                        "three_sigma": 3.0 * math.sqrt(0.25 / shots)}
    out["programs"][name] = rec
    del s
print(json.dumps(out))
