#!/usr/bin/env bash
# BASELINE configs 3 and 5 at their stated sizes on ONE GPU (config 3 on two GPUs: tools/gpu_multi.sh), the reference arm's
# full-sweep check, and the 1-GPU bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/config3_dist.py --local-qubits 33 > gpurun_out/config3_n33_1gpu.json 2> gpurun_out/config3_n33_1gpu.err; echo "config 3 at 33 qubits rc=$?"; tail -c 600 gpurun_out/config3_n33_1gpu.json; echo
timeout 900 python tools/config5.py 32 > gpurun_out/config5_n32.json 2> gpurun_out/config5_n32.err; echo "config 5 at 32 qubits rc=$?"; tail -c 1800 gpurun_out/config5_n32.json; echo; tail -3 gpurun_out/config5_n32.err
timeout 600 python tools/cpu_full_sweep.py 30 > gpurun_out/cpu_full_sweep.json 2> gpurun_out/cpu_full_sweep.err; echo "cpu full sweep rc=$?"; cat gpurun_out/cpu_full_sweep.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "cpu_baseline")})
print("qft", d.get("qft")); print("config3", d.get("config3")); print("sweep", {k: v for k, v in d.get("sweep", {}).items() if k != "per_target_GBps"})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"; tail -c 1200 gpurun_out/bench_ref.json
