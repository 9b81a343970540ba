"""CPU fuzzing of the fused-execution path, no GPU: random circuits x every scheduler / kernel knob, compared with the dense
gate-by-gate statement.

    python tools/fuzz_cpu.py [iterations=200] [seed=1]

Three engines, chosen at random per case: the NumPy interpreter of the micro-program (tests/test_tile_program.py), the
kernels' own code on the CPU emulation (tests/emu/, k_tile or k_tile3), and the lockstep replay
of a register sharded over 2 / 4 / 8 ranks (tests/test_dist_fused_cpu.py).  Knobs drawn per case:
exact / merged, SPZ_TILE_SELECT, SPZ_TILE_LMIN, lazy flush.  Round-1 record: 400 + 300 + 160 cases, no failure
plus 120 sharded ones (the SWAP(11, high) bug fixed in abi.cu: Fuser::fits was found by the unit tests of the tile selection,
not by fuzzing).
"""
import ctypes as C
import math
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

from spinoza_b200 import Controls, Gate, QuantumCircuit, QuantumRegister, QuantumTransformation
from tests import _dense as D
from tests.test_scheduler_plan import KINDS, random_circuit, run_dense_order
from tests.test_tile_cpu_emulation import CUDA_INC, EMU_DIR, run_emulated
from tests.test_dist_fused_cpu import replay
from tests.test_tile_program import run_plan


def swap_heavy(n, count, seed, **kw):
    r = np.random.default_rng(seed)
    qc = QuantumCircuit(QuantumRegister(n), **kw)
    for _ in range(count):
        if r.random() < 0.4:
            qc.swap(int(r.integers(n)), int(r.integers(n)))
            continue
        g = Gate(KINDS[int(r.integers(len(KINDS)))], tuple(float(v) for v in r.random(3) * 2 * math.pi))
        t = int(r.integers(n))
        others = [q for q in range(n) if q != t]
        k = int(r.integers(0, min(6, len(others)) + 1))
        if k == 0:
            qc.add(QuantumTransformation(g, t))
        else:
            cs = [int(c) for c in r.choice(others, size=k, replace=False)]
            qc.add(QuantumTransformation(g, t, Controls.single(cs[0]) if k == 1 else Controls.mixed(cs, set())))
    return qc


def emu_lib():
    out = EMU_DIR / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libtile_emu.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-pthread", f"-I{CUDA_INC}", "-include",
                    str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "tile_emu.cpp"), "-o", str(lib)], check=True, cwd=ROOT)
    h = C.CDLL(str(lib))
    h.emu_tile3_run.restype = C.c_int
    h.emu_tile3_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong, C.POINTER(C.c_int)]
    h.emu_tile1_run.restype = C.c_int
    h.emu_tile1_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong, C.c_int, C.c_int]
    return h


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    h = emu_lib()
    fails = 0
    for it in range(iters):
        n = int(rng.integers(4, 16))
        count = int(rng.integers(1, 220))
        seed = int(rng.integers(1 << 30))
        exact = bool(rng.integers(2))
        os.environ["SPZ_TILE_SELECT"] = str(int(rng.integers(2)))
        os.environ["SPZ_TILE_LMIN"] = str(int(rng.integers(4, 7)))
        os.environ["SPZ_DIST_WINDOW"] = str(int(rng.integers(2)))  # sharded engine only: windows that span exchanges
        os.environ["SPZ_DIST_FUSE_GATE"] = str(int(rng.integers(2)))  # sharded engine only: op accounting of fused exchanges
        gen = swap_heavy if rng.random() < 0.3 else random_circuit
        x = rng.random()
        engine = "emu" if x < 0.35 and n <= 14 else "sharded" if x > 0.75 else "numpy"
        desc = dict(it=it, n=n, count=count, seed=seed, exact=exact, gen=gen.__name__, engine=engine,
                    select=os.environ["SPZ_TILE_SELECT"], lmin=os.environ["SPZ_TILE_LMIN"], window=os.environ["SPZ_DIST_WINDOW"], fuse_gate=os.environ["SPZ_DIST_FUSE_GATE"])
        try:
            qc = gen(n, count, seed, exact=exact)
            trs = list(qc.transformations)
            psi0 = D.random_state(n, seed % 997)
            want = run_dense_order(n, psi0.copy(), trs, range(len(trs)))
            if engine == "numpy":
                got, _ = run_plan(qc, psi0.copy(), lazy=bool(rng.integers(2)))
            elif engine == "sharded":
                world = int(2 ** rng.integers(1, max(2, min(4, n - 3))))  # at least 4 local qubits per rank
                desc["world"] = world
                os.environ.pop("SPZ_DEBUG_BASIS", None)
                if rng.random() < 0.5:  # a register that is still a basis state: its qubits are placed by look-ahead first
                    bx = int(rng.integers(1 << n))
                    psi0 = np.zeros(1 << n, dtype=complex)
                    psi0[bx] = 1.0
                    want = run_dense_order(n, psi0.copy(), trs, range(len(trs)))
                    os.environ["SPZ_DEBUG_BASIS"] = str(bx)
                    desc["basis"] = bx
                try:
                    got, st_ = replay(qc, world, psi0)
                    desc["relabel"] = st_["relabel"]
                finally:
                    os.environ.pop("SPZ_DEBUG_BASIS", None)
            else:
                kernel = 3 if n >= 12 and not exact and rng.random() < 0.7 else 1
                re, im = np.ascontiguousarray(psi0.real), np.ascontiguousarray(psi0.imag)
                run_emulated(h, qc, re, im, {}, kernel=kernel)
                got = re + 1j * im
                desc["kernel"] = kernel
            err = float(np.max(np.abs(got - want)))
            if err > 1e-11:
                fails += 1
                print("MISMATCH", err, desc)
        except Exception as e:  # noqa: BLE001 -- a fuzzer reports everything
            fails += 1
            print("EXCEPTION", repr(e)[:300], desc)
    print(f"{iters} cases, {fails} failures")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
