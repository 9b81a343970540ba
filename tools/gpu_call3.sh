set -u
mkdir -p gpurun_out
TILE_AB_VARIANTS="k_tile3" NCU_VARIANTS=1 bash tools/round2_first_call.sh
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite.log 2>&1; echo "full GPU suite rc=$?"; tail -5 gpurun_out/gpu_suite.log
timeout 300 python tools/bench_reductions.py 30 > gpurun_out/reductions_n30.json 2>&1; echo "reductions rc=$?"; tail -c 1500 gpurun_out/reductions_n30.json
