"""The fused exchange + gate kernel (csrc/kernels_xgate.cuh, opt-in SPZ_DIST_FUSE_GATE=1) on the CPU emulation: two ranks in one
process, every thread block alive at the same time (the protocol is a conversation between block b of rank 0 and block b of
rank 1), system-scope release / acquire mapped to C++ atomics.

Checked here: the result equals "exchange the rank bit with local bit l, then apply the gate on l" bit for bit (oracle
arithmetic), for every gate kind, every l, several grid sizes (including grids larger than the work and work that does not
divide evenly); under ThreadSanitizer no slot is overwritten while the partner still reads it -- and when one rank is told
not to wait for the acknowledgements, ThreadSanitizer does report it, so the check has teeth.
"""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle as orc
from spinoza_b200 import Gate
from tests import _dense as D

ROOT = Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tests" / "emu"
CUDA_INC = Path("/usr/local/cuda/include")
KINDS = [(Gate.KIND_H, ()), (Gate.KIND_X, ()), (Gate.KIND_Y, ()), (Gate.KIND_RX, (0.3,)), (Gate.KIND_RY, (0.3,)), (Gate.KIND_U, (0.3, 0.5, 0.7))]
PARAMS = (0.3, 0.5, 0.7)   # what the stand-alone driver uses


def _gxx():
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    return gxx


@pytest.fixture(scope="module")
def emu():
    out = EMU_DIR / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libxgate_emu.so"
    subprocess.run([_gxx(), "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-pthread", f"-I{CUDA_INC}", "-include",
                    str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "xgate_emu.cpp"), "-o", str(lib)], check=True, cwd=ROOT)
    h = C.CDLL(str(lib))
    h.emu_xgate.restype = C.c_int
    h.emu_xgate.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return h


def expected(n_local, lq, kind, params, re, im):
    """exchange (swap of the rank bit = top bit with local bit lq), then the gate on lq, with the oracle's arithmetic"""
    n = n_local + 1
    psi = D.apply_swap(re + 1j * im, n, n_local, lq)
    s = orc.State(n)
    s.reals[:], s.imags[:] = psi.real, psi.imag
    orc.apply(kind, s, lq, params)
    return s.reals.copy(), s.imags.copy()


def run(emu, n_local, lq, kind, params, grid, seed):
    n = n_local + 1
    init = orc.gen_random_state(n, seed)
    re, im = init.reals.copy(), init.imags.copy()
    half = 1 << n_local
    shards = [x.copy() for x in (re[:half], im[:half], re[half:], im[half:])]   # copies: re / im stay the input
    p = (C.c_double * 3)(*(list(params) + [0.0] * (3 - len(params))))
    rc = emu.emu_xgate(n_local, lq, kind, p, *(s.ctypes.data for s in shards), grid, 0)
    assert rc == 0, rc
    want_re, want_im = expected(n_local, lq, kind, params, re, im)
    got_re = np.concatenate([shards[0], shards[2]])
    got_im = np.concatenate([shards[1], shards[3]])
    assert np.array_equal(got_re, want_re) and np.array_equal(got_im, want_im), (n_local, lq, kind, grid)


@pytest.mark.parametrize("kind,params", KINDS)
def test_every_gate_every_local_bit(emu, kind, params):
    n_local = 10                     # 2^9 pairs = 128 vectors; a block of 32 threads x U = 4 takes 128 vectors per step
    for lq in range(2, n_local):
        run(emu, n_local, lq, kind, params, grid=2, seed=10 * lq + kind)


@pytest.mark.parametrize("n_local,grid", [(4, 1), (5, 3), (8, 1), (9, 4), (10, 3), (11, 2), (12, 3), (12, 5)])
def test_grid_shapes(emu, n_local, grid):
    """more blocks than work, work that does not divide by the grid, several steps per block"""
    run(emu, n_local, n_local - 1, Gate.KIND_H, (), grid, seed=n_local * 7 + grid)
    run(emu, n_local, 2, Gate.KIND_U, (0.3, 0.5, 0.7), grid, seed=n_local * 7 + grid + 1)


# ---- the protocol under ThreadSanitizer ------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def emu_tsan():
    out = EMU_DIR / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "xgate_emu_tsan"
    r = subprocess.run([_gxx(), "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-w", "-fsanitize=thread", "-DSPZ_EMU_TSAN", "-DSPZ_EMU_MAIN", "-pthread",
                        f"-I{CUDA_INC}", "-include", str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "xgate_emu.cpp"), "-o", str(exe)],
                       cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    if subprocess.run([str(exe)], capture_output=True).returncode != 64:
        pytest.skip("ThreadSanitizer cannot run here")
    return exe


def tsan(exe, tmp_path, n_local, lq, kind, grid, brk):
    n = n_local + 1
    init = orc.gen_random_state(n, 5)
    half = 1 << n_local
    state = tmp_path / "state.bin"
    np.concatenate([init.reals[:half], init.imags[:half], init.reals[half:], init.imags[half:]]).tofile(state)
    r = subprocess.run([str(exe), str(n_local), str(lq), str(kind), str(grid), str(brk), str(state)], capture_output=True, text=True,
                       env={"TSAN_OPTIONS": "halt_on_error=0 exitcode=0"}, timeout=600)
    out = np.fromfile(state)
    return r, init, out


@pytest.mark.parametrize("kind", [Gate.KIND_H, Gate.KIND_U])
@pytest.mark.parametrize("n_local,lq,grid", [(11, 10, 2), (11, 3, 3), (12, 5, 2)])
def test_no_race_between_the_two_ranks(emu_tsan, tmp_path, n_local, lq, kind, grid):
    r, init, out = tsan(emu_tsan, tmp_path, n_local, lq, kind, grid, 0)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    half = 1 << n_local
    params = () if kind == Gate.KIND_H else PARAMS
    want_re, want_im = expected(n_local, lq, kind, params, init.reals, init.imags)
    got_re = np.concatenate([out[:half], out[2 * half:3 * half]])
    got_im = np.concatenate([out[half:2 * half], out[3 * half:]])
    assert np.array_equal(got_re, want_re) and np.array_equal(got_im, want_im)


def test_the_race_check_has_teeth(emu_tsan, tmp_path):
    """rank 1 is given a flag base that makes every acknowledgement look as if it had arrived already"""
    r, _, _ = tsan(emu_tsan, tmp_path, 12, 5, Gate.KIND_H, 2, 1)
    assert "data race" in r.stderr
