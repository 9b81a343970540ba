#!/usr/bin/env bash
# GPU call for the fused tile kernels: parity of k_tile3 (csrc/kernels_tile3.cu, TMA) on hardware, A/B timing against k_tile,
# ncu per-launch instruction counts, one full ncu capture of the four QFT-30 passes with source counters.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# Outputs (all under gpurun_out/): tile_tests.log, tile_ab_n30.json, k_tile_v{0,1}_qft30.csv, k_tile3_qft30_full.ncu-rep.
# Every step runs under its own timeout so a hang cannot hold the box.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke: rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_gpu_tile.py -x -q -m gpu > gpurun_out/tile_tests.log 2>&1
echo "tile parity tests: rc=$?"; tail -5 gpurun_out/tile_tests.log
timeout 600 python tools/tile_ab.py 30 5 ${TILE_AB_VARIANTS:-} > gpurun_out/tile_ab_n30.json 2> gpurun_out/tile_ab.err
echo "A/B timing: rc=$?"; tail -3 gpurun_out/tile_ab.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/tile_ab_n30.json"))
    for k, v in d["variants"].items():
        print(f"{k:24s} qft {v['qft_ms']['median']:8.2f} ms  layered {v['layered_d20_ms']['median']:9.2f} ms  "
              f"err {v.get('qft_closed_form_err', float('nan')):.2e}")
except Exception as e:
    print("no A/B result:", e)
PY
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_pipe_fp64.sum,lts__t_bytes.sum
for v in ${NCU_VARIANTS:-1 0}; do
  SPZ_TILE_V3=$v timeout 600 ncu --metrics $M --clock-control none -k regex:k_tile --csv --log-file gpurun_out/k_tile_v${v}_qft30.csv python tools/profile_qft.py 30 > gpurun_out/ncu_v${v}.log 2>&1
  echo "ncu QFT-30 SPZ_TILE_V3=$v: rc=$?"
  SPZ_TILE_V3=$v timeout 600 ncu --metrics $M --clock-control none -k regex:k_tile -c 20 --csv --log-file gpurun_out/k_tile_v${v}_config3.csv python tools/profile_config3.py 30 > gpurun_out/ncu_c3_v${v}.log 2>&1
  echo "ncu config3 SPZ_TILE_V3=$v: rc=$?"
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_tile3 -c 4 -o gpurun_out/k_tile3_qft30_full python tools/profile_qft.py 30 > gpurun_out/ncu_full.log 2>&1
echo "ncu full: rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
