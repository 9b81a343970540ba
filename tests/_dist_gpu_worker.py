"""Worker for tests/test_gpu_dist.py::test_two_gpus_ipc_nvlink (torchrun, one process per GPU, CUDA IPC)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import oracle as orc  # noqa: E402
import spinoza_b200 as sb  # noqa: E402
from spinoza_b200 import QuantumCircuit  # noqa: E402
from spinoza_b200.distributed import DistState, init_from_env  # noqa: E402
from tests.test_gpu_parity import build_circuit, random_ops, run_dense  # noqa: E402


def main():
    env = init_from_env()
    n = 16
    n_local = n - (env.world.bit_length() - 1)
    cpu = orc.gen_random_state(n, 5)
    s = DistState(n, env)
    s.upload(cpu.reals[env.rank << n_local:(env.rank + 1) << n_local], cpu.imags[env.rank << n_local:(env.rank + 1) << n_local])
    ops = random_ops(n, 200, seed=99)
    build_circuit(n, ops, s, fuse=True).execute()
    nrm = sb.norm2(s)
    out = s.gather_logical()
    st = s.stats()
    if env.rank == 0:
        want = run_dense(n, cpu.amps(), ops)
        err = float(np.max(np.abs((out[0] + 1j * out[1]) - want)))
        assert err < 1e-12, err
        assert abs(nrm - 1.0) < 1e-10 and st["exchanges"] > 0
    # bigger shard: exchange bandwidth + QFT round trip
    big = DistState(27, env)
    big.set_basis(12345)
    qc = QuantumCircuit.from_state(big, fuse=True)
    qc.qft(); qc.iqft(list(reversed(range(27))))
    qc.execute()
    nrm = sb.norm2(big)
    stats = big.stats()
    if env.rank == 0:
        assert abs(nrm - 1.0) < 1e-10
        print(f"GPU_DIST_OK err={err:.2e} exchanges={stats['exchanges']} nvlink_GBps={stats['nvlink_GBps_per_direction']}")
    env.shutdown()


if __name__ == "__main__":
    main()
