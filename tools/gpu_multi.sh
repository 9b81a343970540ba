#!/usr/bin/env bash
# Multi-GPU call: parity of the sharded engine on real NVLink (IPC), the opt-in features, and their timing.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_multi.sh 2'
set -u
cd "$(dirname "$0")/.."
N="${1:-2}"
LQ="${2:-30}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/dist_tests_N$N.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/dist_tests_N$N.log
SPZ_TEST_DIST_OPTIN=1 timeout 900 python -m pytest tests/test_gpu_dist_optin.py -x -q -m gpu > gpurun_out/dist_optin_N$N.log 2>&1; echo "opt-in dist tests rc=$?"; tail -3 gpurun_out/dist_optin_N$N.log
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
for v in "default:" "fusegate:SPZ_DIST_FUSE_GATE=1" "window:SPZ_DIST_WINDOW=1" "both:SPZ_DIST_FUSE_GATE=1 SPZ_DIST_WINDOW=1"; do
  name="${v%%:*}"; envs="${v#*:}"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus "$N" --steps 3 --warmup 3 --no-cpu --no-e2e --no-northstar --qubits "$LQ" > "gpurun_out/bench_N${N}_${name}.json" 2> "gpurun_out/bench_N${N}_${name}.err"
  echo "bench $name rc=$?"; python - "gpurun_out/bench_N${N}_${name}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("  value", round(d["value"]), "GB/s; ms/step", round(d["ms_per_step"], 2), "; nvlink GB/s/dir", nv.get("GBps_per_direction_per_gpu"), "; exchanges", nv.get("exchanges_total"),
          "; qft fused s", d.get("qft", {}).get("fused", {}).get("seconds"), "; parity", d.get("parity"))
except Exception as e:
    print("  no line:", e)
PY
done
for w in 0 1; do
  SPZ_DIST_WINDOW=$w timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      tools/config3_dist.py --local-qubits "$LQ" > "gpurun_out/config3_dist_N${N}_window${w}.json" 2> "gpurun_out/config3_dist_N${N}_window${w}.err"
  echo "config3_dist window=$w rc=$?"; tail -c 700 "gpurun_out/config3_dist_N${N}_window${w}.json"; echo
done
