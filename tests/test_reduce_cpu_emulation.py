"""The all-qubit <Z> pass (csrc/kernels_zall.cuh: one read of the state serves xyz_expectation_value('z', ...) for any list
of targets) on the CPU: tests/emu/zall_emu.cpp compiles the kernel's per-thread body with g++ and runs it for a grid of virtual
threads; the sums must match the oracle's one-target-at-a-time expectation values (core.rs:222-264) within 1e-12."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle as orc

ROOT = Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tests" / "emu"
CUDA_INC = Path("/usr/local/cuda/include")


@pytest.fixture(scope="module")
def emu():
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    out = EMU_DIR / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libzall_emu.so"
    cmd = [gxx, "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-shared", "-fPIC", f"-I{CUDA_INC}",
           "-include", str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "zall_emu.cpp"), "-o", str(lib)]
    subprocess.run(cmd, check=True, cwd=ROOT)
    h = C.CDLL(str(lib))
    h.emu_z_all.restype = C.c_int
    h.emu_z_all.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return h


@pytest.mark.parametrize("n,grid,threads", [(2, 1, 256), (3, 1, 256), (5, 2, 256), (9, 3, 256), (11, 1, 256), (12, 1, 256), (14, 7, 256), (17, 4, 256), (18, 296, 256)])
def test_every_qubits_z_from_one_pass(emu, n, grid, threads):
    s = orc.gen_random_state(n, 50 + n)
    out = np.zeros(n + 1)
    assert emu.emu_z_all(n, s.reals.ctypes.data, s.imags.ctypes.data, grid, threads, out.ctypes.data) == 0
    assert abs(out[0] - orc.norm2(s)) < 1e-12
    want = orc.xyz_expectation_value("z", s, list(range(n)))
    got = out[0] - 2.0 * out[1:]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    # prob0 of every qubit falls out of the same numbers (measurement.rs:16-29)
    for t in (0, 1, n // 2, n - 1):
        assert abs((out[0] - out[1 + t]) - orc.prob0(s, t)) < 1e-12


def test_basis_states_are_exact(emu):
    n = 10
    for idx in (0, 1, 2, 5, 0x2aa, (1 << n) - 1):
        s = orc.State(n)
        s.reals[0] = 0.0
        s.reals[idx] = 1.0
        out = np.zeros(n + 1)
        assert emu.emu_z_all(n, s.reals.ctypes.data, s.imags.ctypes.data, 3, 256, out.ctypes.data) == 0
        assert out[0] == 1.0
        assert [int(x) for x in out[1:]] == [(idx >> t) & 1 for t in range(n)]


@pytest.fixture(scope="module")
def emu_xy():
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    out = EMU_DIR / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libxyall_emu.so"
    cmd = [gxx, "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-pthread", f"-I{CUDA_INC}",
           "-include", str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "xyall_emu.cpp"), "-o", str(lib)]
    subprocess.run(cmd, check=True, cwd=ROOT)
    h = C.CDLL(str(lib))
    h.emu_xy_tile.restype = C.c_int
    h.emu_xy_tile.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return h


@pytest.mark.parametrize("n,L,high,grid,threads", [
    (7, 7, [], 1, 32), (9, 9, [], 2, 64), (12, 12, [], 1, 64), (14, 12, [], 3, 64),      # contiguous tiles: the targets below bit 12
    (14, 6, [12, 13], 2, 32), (15, 6, [7, 9, 10, 12, 13, 14], 3, 64), (16, 6, [12, 15], 5, 32), (13, 6, [6], 1, 64)])
def test_x_and_y_of_every_tile_qubit_from_one_pass(emu_xy, n, L, high, grid, threads):
    """The tile body of the batched <X> / <Y> pass (csrc/kernels_xyall.cuh) against the oracle's one-target-at-a-time values
    (core.rs:222-264), within 1e-12: contiguous tiles, and tiles of 64 contiguous amplitudes x arbitrary higher qubits."""
    s = orc.gen_random_state(n, 80 + n + len(high))
    hi = np.array(high + [0] * (8 - len(high)), dtype=np.int32)
    tile_qubits = list(range(L)) + high
    tmask = sum(1 << b for b in range(len(tile_qubits)))
    for obs, name in ((0, "x"), (1, "y")):
        out = np.zeros(12)
        assert emu_xy.emu_xy_tile(n, s.reals.ctypes.data, s.imags.ctypes.data, L, len(high), hi.ctypes.data, tmask, obs, grid, threads,
                                  out.ctypes.data) == 0
        want = orc.xyz_expectation_value(name, s, tile_qubits)
        np.testing.assert_allclose(out[: len(tile_qubits)], want, rtol=0, atol=1e-12)
    # a subset of the tile bits: the others are left alone
    out = np.full(12, 7.0)
    sub = tmask & 0b101001000001
    assert emu_xy.emu_xy_tile(n, s.reals.ctypes.data, s.imags.ctypes.data, L, len(high), hi.ctypes.data, sub, 0, grid, threads, out.ctypes.data) == 0
    for b, q in enumerate(tile_qubits):
        if (sub >> b) & 1:
            assert abs(out[b] - orc.xyz_expectation_value("x", s, [q])[0]) < 1e-12
        else:
            assert out[b] == 0.0
