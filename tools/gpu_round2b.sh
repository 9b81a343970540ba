#!/usr/bin/env bash
# 1-GPU call: the whole GPU suite, the reductions table, config 5 at 32 qubits, measured DRAM traffic, the bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/gpu_suite.log
timeout 300 python tools/bench_reductions.py 30 > gpurun_out/reductions_n30.json 2> gpurun_out/reductions.err; echo "reductions rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/reductions_n30.json')); print({k: round(v['ms'],2) for k,v in d['rows'].items()})"
timeout 900 python tools/config5.py 32 > gpurun_out/config5_n32.json 2> gpurun_out/config5_n32.err; echo "config 5 at 32 qubits rc=$?"; tail -c 1500 gpurun_out/config5_n32.json; echo; tail -3 gpurun_out/config5_n32.err
timeout 300 python tools/measure_traffic.py 30 > gpurun_out/traffic.log 2>&1; echo "traffic rc=$?"; tail -c 600 gpurun_out/traffic.log; cp profiles/round2_traffic_n30.json gpurun_out/ 2>/dev/null
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e")}); print(d["roofline"])
print("qft", d.get("qft")); print("config3", d.get("config3")); print("sweep", {k: v for k, v in d.get("sweep", {}).items() if k != "per_target_GBps"})
PY
