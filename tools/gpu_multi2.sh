#!/usr/bin/env bash
# 2-GPU call: parity on real NVLink, then the chunked overlapped exchange at several chunk counts / CTA counts, then config 3 at
# 2^32 amplitudes per GPU.      /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_multi2.sh 2'
set -u
cd "$(dirname "$0")/.."
N="${1:-2}"
LQ="${2:-30}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/dist_tests_N$N.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/dist_tests_N$N.log
for v in "k4:" "k1:SPZ_XCHG_CHUNKS=1" "k2:SPZ_XCHG_CHUNKS=2" "k8:SPZ_XCHG_CHUNKS=8" "k4c64:SPZ_XCHG_CTAS=64" "k4c96:SPZ_XCHG_CTAS=96" "noov:SPZ_NO_OVERLAP=1"; do
  name="${v%%:*}"; envs="${v#*:}"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus "$N" --steps 3 --warmup 3 --no-cpu --no-e2e --no-northstar --qubits "$LQ" > "gpurun_out/bench_N${N}_${name}.json" 2> "gpurun_out/bench_N${N}_${name}.err"
  echo "bench $name rc=$?"; python - "gpurun_out/bench_N${N}_${name}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("  value", round(d["value"]), "GB/s; ms/step", round(d["ms_per_step"], 2), "; nvlink GB/s/dir", round(nv.get("GBps_per_direction_per_gpu", 0), 1), "; exchanges", nv.get("exchanges_total"),
          "; qft fused s", round(d.get("qft", {}).get("fused", {}).get("seconds", 0), 4), "; parity", d.get("parity", {}).get("sharded_vs_oracle", {}).get("max_abs_err"), d.get("parity", {}).get("qft_closed_form_max_abs_err"))
except Exception as e:
    print("  no line:", e)
PY
done
if [ "${3:-}" != "skip3" ]; then
SPZ_DIST_WINDOW=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    tools/config3_dist.py --local-qubits 32 > "gpurun_out/config3_dist_N${N}_q33.json" 2> "gpurun_out/config3_dist_N${N}_q33.err"
echo "config3_dist 33q rc=$?"; tail -c 900 "gpurun_out/config3_dist_N${N}_q33.json"; echo; tail -5 "gpurun_out/config3_dist_N${N}_q33.err"
fi
