#!/usr/bin/env bash
# Quick GPU check of the fused tile kernel: parity tests, QFT-30 / config-3 timing, per-launch ncu counters.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py -x -q -m gpu > gpurun_out/tile_tests.log 2>&1
echo "tile parity tests: rc=$?"; tail -2 gpurun_out/tile_tests.log
timeout 600 python tools/tile_ab.py 30 5 k_tile3 > gpurun_out/tile_ab_n30.json 2> gpurun_out/tile_ab.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/tile_ab_n30.json"))
for k, v in d["variants"].items():
    print(f"{k:24s} qft {v['qft_ms']['median']:8.2f} ms  layered {v['layered_d20_ms']['median']:9.2f} ms  err {v.get('qft_closed_form_err', float('nan')):.2e}")
PY
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_tile -c 4 --csv --log-file gpurun_out/q_qft30.csv python tools/profile_qft.py 30 > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_tile -c 17 --csv --log-file gpurun_out/q_config3.csv python tools/profile_config3.py 30 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/q_qft30.csv", "gpurun_out/q_config3.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    data = {}
    for r in rows[1:]:
        data.setdefault(r[idx["ID"]], {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    tot = 0
    for k, m in data.items():
        w = 2 ** 21
        tot += m["gpu__time_duration.sum"] / 1e6
        print(f, k, f"ms {m['gpu__time_duration.sum']/1e6:6.2f} inst/tilewarp {m['smsp__inst_executed.sum']/w:7.0f} fp64 {m['smsp__inst_executed_pipe_fp64.sum']/w:6.0f} issue% {m['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f} bankconf {m['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']/1e6:7.1f}M dram {(m['dram__bytes_read.sum']+m['dram__bytes_write.sum'])/1e9:5.1f} GB")
    print("total ms under ncu", tot)
PY
SPZ_TILE_PROF=1 python tools/profile_qft.py 30 2>&1 | tail -5
