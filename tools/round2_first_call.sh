#!/usr/bin/env bash
# First GPU call of round 2: validate and measure the opt-in tile kernel k_tile2 (csrc/kernels_tile2.cu), which was written
# and checked on the CPU emulation only (tests/test_tile_cpu_emulation.py) after round 1's GPU budget was spent.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'            # ~15 min
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/round2_first_call.sh --full-suite' # + the whole GPU suite under k_tile2
#
# Outputs (all under gpurun_out/):
#   v2_tests.log            parity of k_tile2 against k_tile and the oracle, the tile-segment knob, and exchange-spanning
#                           windows on sharded registers (tests/test_gpu_tile_v2.py: everything that is opt-in)
#   tile_ab_n30.json        QFT-30 and config-3 timings, k_tile vs k_tile2 at every direct-transfer level
#   k_tile_v{0,1}_qft30.csv ncu per-launch duration / instructions / issue utilisation of the four QFT-30 passes
# Every step runs under its own timeout so a hang cannot hold the box.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SPZ_TEST_TILE_V2=1
timeout 900 python -m pytest tests/test_gpu_tile_v2.py -x -q -m gpu > gpurun_out/v2_tests.log 2>&1
echo "k_tile2 parity tests: rc=$?"; tail -3 gpurun_out/v2_tests.log
timeout 600 python tools/tile_ab.py 30 5 > gpurun_out/tile_ab_n30.json 2> gpurun_out/tile_ab.err
echo "A/B timing: rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/tile_ab_n30.json"))
    for k, v in d["variants"].items():
        print(f"{k:18s} qft {v['qft_ms']['median']:8.2f} ms  layered {v['layered_d20_ms']['median']:9.2f} ms  "
              f"err {v.get('qft_closed_form_err', float('nan')):.2e}")
except Exception as e:
    print("no A/B result:", e)
PY
if [ "${1:-}" = "--full-suite" ]; then
  # the whole GPU suite with k_tile2 as the fused kernel (tests/conftest.py pins the tuning switches unless told otherwise)
  SPZ_TEST_KEEP_ENV=1 SPZ_TILE_V2=1 timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite_under_v2.log 2>&1
  echo "GPU suite under SPZ_TILE_V2=1: rc=$?"; tail -3 gpurun_out/gpu_suite_under_v2.log
fi
for v in 0 1; do
  SPZ_TILE_V2=$v timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:k_tile --csv --log-file gpurun_out/k_tile_v${v}_qft30.csv python tools/profile_qft.py 30 > gpurun_out/ncu_v${v}.log 2>&1
  echo "ncu SPZ_TILE_V2=$v: rc=$?"
done
