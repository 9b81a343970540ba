"""DRAM traffic of the headline kernel, measured (not quoted): one ncu pass over a few single-gate launches at n qubits.

    python tools/measure_traffic.py [n=30]      ->  profiles/round2_traffic_n<n>.json (bench.py reads it for roofline.traffic)

Runs `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` on this same script in worker mode (H, RX,
RZ on a middle and on a low target), and records per kernel launch: bytes read + written, against the algorithmic 32 * 2^n
(RZ included: rz_apply multiplies every amplitude, gates.rs:930-947).  Times under ncu are not bench values.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if len(sys.argv) > 2 and sys.argv[2] == "worker":
    import spinoza_b200 as sb
    from spinoza_b200 import Gate
    n = int(sys.argv[1])
    s = sb.State(n)
    s.init_random(42)
    s.sync()
    for g in (Gate.H, Gate.RX(1.0), Gate.RZ(1.0)):
        for t in (n // 2, 1):
            sb.apply(g, s, t)
    s.sync()
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
       "-k", "regex:k_pair", "--csv", sys.executable, __file__, str(n), "worker"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    rec = launches.setdefault(r[ix["ID"]], {"kernel": r[ix["Kernel Name"]].split("(")[0]})
    v, unit = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
    rec[r[ix["Metric Name"]]] = v * scale
names = ["H@mid", "H@1", "RX@mid", "RX@1", "RZ@mid", "RZ@1"]
res = {"qubits": n, "algorithmic_bytes_full_pass": 32.0 * (1 << n), "launches": []}
for (k, rec), nm in zip(sorted(launches.items(), key=lambda kv: int(kv[0])), names):
    alg = 32.0 * (1 << n)
    tr = rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
    res["launches"].append({"gate": nm, "kernel": rec["kernel"], "dram_read": rec.get("dram__bytes_read.sum"), "dram_write": rec.get("dram__bytes_write.sum"),
                            "traffic": tr, "algorithmic": alg, "traffic_over_algorithmic": tr / alg, "ms_under_ncu": rec.get("gpu__time_duration.sum")})
full = [l["traffic"] for l in res["launches"]]
res["traffic_full_pass_mean"] = sum(full) / len(full) if full else None
res["source"] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (tools/measure_traffic.py)"
path = ROOT / "profiles" / f"round2_traffic_n{n}.json"
path.write_text(json.dumps(res, indent=1) + "\n")
print(json.dumps(res))
