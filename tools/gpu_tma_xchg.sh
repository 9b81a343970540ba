#!/usr/bin/env bash
# 2-GPU A/B of the bulk-copy exchange kernel (SPZ_XCHG_TMA=1) against the load/store one: parity first, then the sweep.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SPZ_RDV_TIMEOUT_MS=60000
SPZ_XCHG_TMA=1 timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/dist_tests_tma.log 2>&1; echo "dist tests (TMA exchange) rc=$?"; tail -3 gpurun_out/dist_tests_tma.log
for v in ${SPZ_TMA_VARIANTS:-"lsu:" "tma32:SPZ_XCHG_TMA=1" "tma64:SPZ_XCHG_TMA=1 SPZ_XCHG_TMA_CTAS=64" "tma16:SPZ_XCHG_TMA=1 SPZ_XCHG_TMA_CTAS=16"}; do
  name="${v%%:*}"; envs="${v#*:}"
  env $envs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --no-northstar --no-extras > "gpurun_out/bench_N2_${name}.json" 2> "gpurun_out/bench_N2_${name}.err"
  echo "bench $name rc=$?"; python - "gpurun_out/bench_N2_${name}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("  value", round(d["value"]), "GB/s; ms/step", round(d["ms_per_step"], 2), "; nvlink GB/s/dir", round(nv.get("GBps_per_direction_per_gpu", 0), 1), "; exchanges", nv.get("exchanges_total"))
except Exception as e:
    print("  no line:", e)
PY
done
