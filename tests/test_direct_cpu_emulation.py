"""The one-gate-per-pass path (csrc/kernels_direct.cu: launch_gate / launch_swap and the k_pair_* / k_swap_* kernels) executed on
the CPU: tests/emu/direct_emu.cpp compiles the product's dispatch code and kernel bodies with g++ (the grid becomes two nested
loops) and this file checks them against the oracle -- bit for bit, as the GPU parity tests do -- for every gate, every target,
and every shape of control mask (none / lane-level / vector-level / mixed / many), at register sizes from 1 qubit up.

It cannot say anything about coalescing or bandwidth; it guards the index arithmetic (zero-bit insertion, lane predicates for
controls below log2(W), the in-register low-target path, the scalar fallback for tiny registers) in the CPU suite.
"""
import ctypes as C
import itertools
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
PI = np.pi

ONE_Q = [(Gate.KIND_H, ()), (Gate.KIND_X, ()), (Gate.KIND_Y, ()), (Gate.KIND_Z, ()), (Gate.KIND_P, (0.7,)), (Gate.KIND_RX, (1.0,)),
         (Gate.KIND_RY, (1.3,)), (Gate.KIND_RZ, (1.0,)), (Gate.KIND_U, (1.0, 2.0, 3.0))]
C_OK = [g for g in ONE_Q if g[0] != Gate.KIND_Z]                                          # c_apply  gates.rs:257-269
MC_OK = [g for g in ONE_Q if g[0] in (Gate.KIND_X, Gate.KIND_P, Gate.KIND_RX, Gate.KIND_RY)]  # mc_apply gates.rs:290-320


@pytest.fixture(scope="module")
def emu():
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    out = EMU_DIR / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libdirect_emu.so"
    cmd = [gxx, "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-shared", "-fPIC", f"-I{CUDA_INC}",
           "-include", str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "direct_emu.cpp"), "-o", str(lib)]
    subprocess.run(cmd, check=True, cwd=ROOT)
    h = C.CDLL(str(lib))
    h.emu_apply.restype = C.c_int
    h.emu_apply.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_ulonglong, C.c_int]
    h.emu_apply_signed.restype = C.c_int
    h.emu_apply_signed.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_ulonglong, C.c_ulonglong, C.c_int]
    h.emu_swap.restype = C.c_int
    h.emu_swap.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    h.emu_last_error.restype = C.c_char_p
    return h


def emu_apply(emu, n, re, im, kind, params, cmask, target):
    p = (C.c_double * 3)(*(list(params) + [0.0] * (3 - len(params))))
    return emu.emu_apply(n, re.ctypes.data, im.ctypes.data, kind, p, cmask, target)


def fresh(n, seed):
    s = orc.gen_random_state(n, seed)
    return s, s.reals.copy(), s.imags.copy()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 11])
def test_apply_every_gate_every_target_bit_exact(emu, n):
    for (kind, params), t in itertools.product(ONE_Q, range(n)):
        cpu, re, im = fresh(n, 10 * n + t)
        assert emu_apply(emu, n, re, im, kind, params, 0, t) == 0, emu.emu_last_error()
        orc.apply(kind, cpu, t, params)
        assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags), (kind, t)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 10])
def test_c_apply_every_gate_every_pair_bit_exact(emu, n):
    for (kind, params), c, t in itertools.product(C_OK, range(n), range(n)):
        if c == t:
            continue
        cpu, re, im = fresh(n, 100 + 7 * c + t)
        assert emu_apply(emu, n, re, im, kind, params, 1 << c, t) == 0, emu.emu_last_error()
        orc.c_apply(kind, cpu, c, t, params)
        assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags), (kind, c, t)


@pytest.mark.parametrize("n", [3, 4, 6, 9])
def test_mc_apply_all_small_masks_bit_exact(emu, n):
    rng = np.random.default_rng(n)
    for kind, params in MC_OK:
        for t in range(n):
            others = [q for q in range(n) if q != t]
            masks = set()
            for k in (2, 3, min(5, len(others))):
                if k <= len(others):
                    for _ in range(4):
                        masks.add(tuple(sorted(int(c) for c in rng.choice(others, size=k, replace=False))))
            masks.add(tuple(others))  # every other qubit controls: one pair only
            for cs in masks:
                cm = sum(1 << c for c in cs)
                cpu, re, im = fresh(n, 300 + t)
                assert emu_apply(emu, n, re, im, kind, params, cm, t) == 0, emu.emu_last_error()
                orc.mc_apply(kind, cpu, list(cs), None, t, params)
                assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags), (kind, cs, t)


@pytest.mark.parametrize("n", [2, 3, 5, 6, 9])
def test_cells_the_reference_panics_on_match_the_dense_statement(emu, n):
    """controlled-Z, multi-controlled H / Y / Z / RZ / U: not in the oracle (the reference hits todo!()), computed anyway."""
    rng = np.random.default_rng(40 + n)
    for kind, params in ONE_Q:
        for _ in range(6):
            t = int(rng.integers(n))
            others = [q for q in range(n) if q != t]
            k = int(rng.integers(1, min(4, len(others)) + 1))
            cs = [int(c) for c in rng.choice(others, size=k, replace=False)]
            cm = sum(1 << c for c in cs)
            psi = D.random_state(n, int(rng.integers(1 << 30)))
            re, im = np.ascontiguousarray(psi.real), np.ascontiguousarray(psi.imag)
            assert emu_apply(emu, n, re, im, kind, params, cm, t) == 0
            want = D.apply_matrix(psi, n, D.matrix(kind, params), t, cm)
            np.testing.assert_allclose(re + 1j * im, want, rtol=0, atol=1e-14)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 10])
def test_signed_controls_equal_x_conjugation_bit_exact(emu, n):
    """The spz_mc_apply_signed extension (true negative controls; the reference's Mixed { zeros } drops them, gates.rs:298-311):
    one launch must give exactly what X on the zero-controls, the all-ones gate of the oracle, and X again give -- X is an
    exchange, so the same pair updates run on the same values.  Controls below and above log2(W), every target."""
    rng = np.random.default_rng(900 + n)
    for kind, params in ONE_Q:
        for t in range(n):
            others = [q for q in range(n) if q != t]
            for _ in range(6):
                k = int(rng.integers(1, min(4, len(others)) + 1))
                cs = [int(c) for c in rng.choice(others, size=k, replace=False)]
                zs = [c for c in cs if rng.random() < 0.6] or [cs[0]]
                cm, zm = sum(1 << c for c in cs), sum(1 << c for c in zs)
                cpu, re, im = fresh(n, 700 + t)
                psi = cpu.amps()
                p = (C.c_double * 3)(*(list(params) + [0.0] * (3 - len(params))))
                assert emu.emu_apply_signed(n, re.ctypes.data, im.ctypes.data, kind, p, cm, zm, t) == 0, emu.emu_last_error()
                want = D.apply_matrix(psi, n, D.matrix(kind, params), t, cm, zm)
                np.testing.assert_allclose(re + 1j * im, want, rtol=0, atol=1e-14)
                if kind == Gate.KIND_Z or (k > 1 and (kind, params) not in MC_OK):
                    continue  # cells the oracle (like the reference) has no loop for: the dense statement above stands alone
                for z in zs:
                    orc.apply(Gate.KIND_X, cpu, z, ())
                if k == 1:
                    orc.c_apply(kind, cpu, cs[0], t, params)
                else:
                    orc.mc_apply(kind, cpu, cs, None, t, params)
                for z in zs:
                    orc.apply(Gate.KIND_X, cpu, z, ())
                assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags), (kind, cs, zs, t)


def test_signed_controls_argument_errors(emu):
    n = 4
    _, re, im = fresh(n, 2)
    p = (C.c_double * 3)(0.0, 0.0, 0.0)
    assert emu.emu_apply_signed(n, re.ctypes.data, im.ctypes.data, Gate.KIND_X, p, 0b0010, 0b0100, 0) != 0  # zeros not in controls
    assert emu.emu_apply_signed(n, re.ctypes.data, im.ctypes.data, Gate.KIND_X, p, 0b0011, 0b0001, 0) != 0  # target is a control
    before = re.copy()
    assert emu.emu_apply_signed(n, re.ctypes.data, im.ctypes.data, Gate.KIND_X, p, 0b0110, 0, 0) == 0       # no zeros: plain mc gate
    assert not np.array_equal(before, re)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 10])
def test_swap_every_pair_bit_exact(emu, n):
    for a, b in itertools.product(range(n), range(n)):
        cpu, re, im = fresh(n, 500 + a * n + b)
        assert emu.emu_swap(n, re.ctypes.data, im.ctypes.data, a, b) == 0
        if a != b:
            orc.swap(cpu, a, b)
        assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags), (a, b)


def test_argument_errors_are_status_codes(emu):
    n = 4
    _, re, im = fresh(n, 1)
    assert emu_apply(emu, n, re, im, Gate.KIND_H, (), 0, 4) != 0          # target out of range
    assert emu_apply(emu, n, re, im, Gate.KIND_H, (), 1 << 2, 2) != 0     # target is also a control
    assert emu_apply(emu, n, re, im, Gate.KIND_H, (), 1 << 5, 0) != 0     # control outside the register
    assert emu_apply(emu, n, re, im, Gate.KIND_SWAP, (), 0, 0) != 0       # not a pair-update gate
    assert emu.emu_swap(n, re.ctypes.data, im.ctypes.data, 0, 4) != 0
    assert b"out of range" in emu.emu_last_error()
