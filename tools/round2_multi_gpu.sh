#!/usr/bin/env bash
# Round 2, multi-GPU call: what the exchange-spanning windows (SPZ_DIST_WINDOW=1) are worth on real NVLink.
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/round2_multi_gpu.sh 2'
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1800 -- 'bash tools/round2_multi_gpu.sh 8'
#
# Outputs under gpurun_out/: config3_dist_N{n}_window{0,1}.json (BASELINE config 3, 30 local qubits per GPU) and
# qft_dist_N{n}_window{0,1}.json (sharded QFT).  Every step under its own timeout; rendezvous on 127.0.0.1.
set -u
cd "$(dirname "$0")/.."
N="${1:-2}"
LQ="${2:-30}"
mkdir -p gpurun_out
run() { # name, window, script, args...
  local name="$1" window="$2"; shift 2
  SPZ_DIST_WINDOW="$window" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 200)) "$@" > "gpurun_out/${name}_N${N}_window${window}.json" 2> "gpurun_out/${name}_N${N}_window${window}.err"
  echo "$name N=$N window=$window rc=$?"; tail -c 600 "gpurun_out/${name}_N${N}_window${window}.json"; echo
}
for w in 0 1; do
  run config3_dist "$w" tools/config3_dist.py --local-qubits "$LQ"
  run qft_dist "$w" tools/qft_dist.py --local-qubits "$LQ"
done
# exchange fused with the gate that asked for it (kernels_xgate.cuh): QFT and the bench's 1-qubit sweep
export SPZ_DIST_FUSE_GATE=1
run qft_dist_fusegate 0 tools/qft_dist.py --local-qubits "$LQ"
run qft_dist_fusegate 1 tools/qft_dist.py --local-qubits "$LQ"          # both: windows spanning exchanges + fused gates
run config3_dist_fusegate 1 tools/config3_dist.py --local-qubits "$LQ"
SPZ_DIST_FUSE_GATE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus "$N" --steps 2 --warmup 3 --no-cpu > "gpurun_out/bench_N${N}_fusegate.json" 2> "gpurun_out/bench_N${N}_fusegate.err"
echo "bench with fused exchange+gate rc=$?"; tail -c 400 "gpurun_out/bench_N${N}_fusegate.json"; echo
