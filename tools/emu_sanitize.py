"""AddressSanitizer / UBSan over the kernels' own code on the CPU emulation (tests/emu/): out-of-bounds shared-memory or
state accesses and undefined behaviour in the index arithmetic, which neither the parity tests nor ThreadSanitizer see.

    python tools/emu_sanitize.py            # ~1 minute, no GPU; exits non-zero on any report

Tile kernels (k_tile with both decode variants, k_tile3 through its host lowering; exact and merged programs, short tile segments) run through the
stand-alone driver of tests/emu/tile_emu.cpp; the one-gate-per-pass kernels through a sanitised shared object loaded into a
Python started with libasan preloaded.
"""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GXX = "/usr/bin/g++"
COMMON = ["-O1", "-g", "-std=c++17", "-ffp-contract=off", "-w", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
          "-I/usr/local/cuda/include", "-include", str(ROOT / "tests/emu/cuda_cpu_shim.h"), "-x", "c++"]


def tile(tmp):
    import numpy as np
    from spinoza_b200 import QuantumCircuit, QuantumRegister, workloads
    from tests import _dense as D
    from tests.test_scheduler_plan import random_circuit
    from tests.test_tile_cpu_emulation import raw_pass, reference_cells_circuit
    exe = tmp / "tile_emu_asan"
    subprocess.run([GXX, *COMMON, "-DSPZ_EMU_MAIN", "-pthread", str(ROOT / "tests/emu/tile_emu.cpp"), "-o", str(exe)], check=True, cwd=ROOT)
    n = 13
    cases = []
    qc = QuantumCircuit(QuantumRegister(n)); qc.qft(); cases.append(("qft", qc, 0))
    cases.append(("random", random_circuit(n, 160, 31), 0))
    cases.append(("exact", reference_cells_circuit(n, 60, 41, exact=True), 1))
    os.environ["SPZ_TILE_LMIN"] = "4"
    qc = QuantumCircuit(QuantumRegister(n)); workloads.random_layered_circuit(qc, depth=6, seed=3); cases.append(("lmin4", qc, 0))
    bad = runs = 0
    for name, qc, exact in cases:
        psi = D.random_state(n, 1)
        _, n_pass = qc.plan()
        for p in range(n_pass):
            blob = raw_pass(qc, p)
            if int(np.frombuffer(blob, dtype="<i4", count=1)[0]) != 0:
                continue
            (tmp / "blob.bin").write_bytes(blob)
            for kernel, opts in ((1, (0, 1)), (3, (0,))):
                if kernel == 3 and exact:
                    continue  # k_tile3 runs merged mode only
                for opt in opts:
                    np.concatenate([psi.real, psi.imag]).tofile(tmp / "state.bin")
                    r = subprocess.run([str(exe), str(kernel), str(n), str(exact), str(tmp / "state.bin"), str(tmp / "blob.bin"), str(opt)],
                                       capture_output=True, text=True, timeout=300)
                    runs += 1
                    if r.returncode not in (0, 71) or "ERROR" in r.stderr or "runtime error" in r.stderr:  # 71: not eligible for k_tile3
                        bad += 1
                        print(name, "pass", p, "kernel", kernel, "option", opt, "rc", r.returncode, r.stderr[:1500])
    os.environ.pop("SPZ_TILE_LMIN", None)
    print(f"tile kernels: {runs} sanitised runs, {bad} reports")
    return bad


DIRECT_DRIVER = r'''
import ctypes as C, itertools, sys
import numpy as np
h = C.CDLL(sys.argv[1])
h.emu_apply.restype = C.c_int
h.emu_apply.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_ulonglong, C.c_int]
h.emu_swap.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
h.emu_apply_signed.restype = C.c_int
h.emu_apply_signed.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_ulonglong, C.c_ulonglong, C.c_int]
rng = np.random.default_rng(0)
cnt = 0
for n in range(1, 12):
    for kind in (0, 2, 3, 4, 5, 6, 7, 8, 10):
        for t in range(n):
            for trial in range(3):
                others = [q for q in range(n) if q != t]
                k = int(rng.integers(0, min(4, len(others)) + 1))
                cm = sum(1 << int(c) for c in rng.choice(others, size=k, replace=False)) if k else 0
                re = np.ascontiguousarray(rng.random(1 << n)); im = np.ascontiguousarray(rng.random(1 << n))
                assert h.emu_apply(n, re.ctypes.data, im.ctypes.data, kind, (C.c_double * 3)(0.3, 0.5, 0.7), cm, t) == 0
                cnt += 1
                if cm:  # the same mask with some controls negative (spz_mc_apply_signed)
                    neg = cm & int(rng.integers(1, 1 << n))
                    assert h.emu_apply_signed(n, re.ctypes.data, im.ctypes.data, kind, (C.c_double * 3)(0.3, 0.5, 0.7), cm, neg, t) == 0
                    cnt += 1
    for a, b in itertools.product(range(n), range(n)):
        re = np.ascontiguousarray(rng.random(1 << n)); im = np.ascontiguousarray(rng.random(1 << n))
        assert h.emu_swap(n, re.ctypes.data, im.ctypes.data, a, b) == 0
print("direct kernels:", cnt, "sanitised gate applications, 0 reports")
'''


ZALL_DRIVER = r'''
import ctypes as C, sys
import numpy as np
h = C.CDLL(sys.argv[1])
h.emu_z_all.restype = C.c_int
h.emu_z_all.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
rng = np.random.default_rng(1)
cnt = 0
for n in range(2, 19):
    for grid in (1, 2, 5, 296):
        re = np.ascontiguousarray(rng.random(1 << n)); im = np.ascontiguousarray(rng.random(1 << n))
        out = np.zeros(n + 1)
        assert h.emu_z_all(n, re.ctypes.data, im.ctypes.data, grid, 256, out.ctypes.data) == 0
        assert abs(out[0] - (re @ re + im @ im)) < 1e-9 * (1 << n)
        cnt += 1
print("all-qubit <Z> pass:", cnt, "sanitised runs, 0 reports")
'''


def zall(tmp):
    lib = tmp / "libzall_emu_asan.so"
    subprocess.run([GXX, *COMMON, "-shared", "-fPIC", str(ROOT / "tests/emu/zall_emu.cpp"), "-o", str(lib)], check=True, cwd=ROOT)
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-c", ZALL_DRIVER, str(lib)], env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout.strip() or r.stderr[-1500:])
    return 0 if r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr else 1


XY_DRIVER = r'''
import ctypes as C, sys
import numpy as np
h = C.CDLL(sys.argv[1])
h.emu_xy_tile.restype = C.c_int
h.emu_xy_tile.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p]
rng = np.random.default_rng(2)
cnt = 0
cases = [(7, 7, []), (8, 8, []), (11, 11, []), (12, 12, []), (13, 12, []), (13, 6, [6]), (14, 6, [9, 13]), (15, 6, [6, 7, 8, 9, 10, 11]),
         (16, 6, [12, 13, 14, 15]), (15, 6, [8, 10, 12, 13, 14])]
for n, L, high in cases:
    # exact-size buffers: any read past the state is an ASan report
    re = np.ascontiguousarray(rng.random(1 << n)); im = np.ascontiguousarray(rng.random(1 << n))
    hi = np.array(high + [0] * (8 - len(high)), dtype=np.int32)
    for obs in (0, 1):
        for grid, threads in ((1, 32), (3, 64)):
            out = np.zeros(12)
            tmask = (1 << (L + len(high))) - 1
            assert h.emu_xy_tile(n, re.ctypes.data, im.ctypes.data, L, len(high), hi.ctypes.data, tmask, obs, grid, threads, out.ctypes.data) == 0
            cnt += 1
print("batched <X>/<Y> tiles:", cnt, "sanitised runs, 0 reports")
'''


def xyall(tmp):
    lib = tmp / "libxyall_emu_asan.so"
    subprocess.run([GXX, *COMMON, "-shared", "-fPIC", "-pthread", str(ROOT / "tests/emu/xyall_emu.cpp"), "-o", str(lib)], check=True, cwd=ROOT)
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-c", XY_DRIVER, str(lib)], env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout.strip() or r.stderr[-1500:])
    return 0 if r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr else 1


def direct(tmp):
    lib = tmp / "libdirect_emu_asan.so"
    subprocess.run([GXX, *COMMON, "-shared", "-fPIC", str(ROOT / "tests/emu/direct_emu.cpp"), "-o", str(lib)], check=True, cwd=ROOT)
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-c", DIRECT_DRIVER, str(lib)], env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout.strip() or r.stderr[-1500:])
    return 0 if r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr else 1


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        tmp = Path(d)
        sys.exit(1 if tile(tmp) + direct(tmp) + zall(tmp) + xyall(tmp) else 0)
