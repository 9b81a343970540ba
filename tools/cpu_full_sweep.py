"""How good is the weighted target sample of bench.py's reference arm?  Times the FULL 1-qubit sweep (H / RX / RZ on every
target) with the oracle port on this machine's host cores, then the weighted sample, and prints both rates.

    python tools/cpu_full_sweep.py [n=30] > gpurun_out/cpu_full_sweep.json
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import oracle as orc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
threads = orc.max_threads()
n = bench.pick_cpu_n(n)
full, t_full, passes_full, _ = bench.cpu_sweep_weighted(n, threads, full=True)
samp, t_samp, passes_samp, est = bench.cpu_sweep_weighted(n, threads)
print(json.dumps({"qubits": n, "threads": threads, "full_sweep": {"GBps": full, "seconds": t_full, "passes": passes_full},
                  "weighted_sample": {"GBps": samp, "seconds_measured": t_samp, "passes": passes_samp, "seconds_estimated_full": est,
                                      "targets": bench.cpu_targets(n, threads)},
                  "ratio_sample_over_full": samp / full}))
