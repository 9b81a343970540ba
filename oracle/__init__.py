"""ctypes loader for the CPU oracle (oracle/spinoza_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  Nothing under ``spinoza_b200/``
imports this package.

Parity status: pinned against the reference's golden vectors (tests/test_oracle_golden.py).
The real reference (Rust) cannot be compiled in this image, so this is a "port".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libspinoza_oracle.so"

# gate kinds, same numbering as include/spinoza_b200.h
H, M, X, Y, Z, P, RX, RY, RZ, SWAP, U, UNITARY, BITFLIP = range(13)
OK, ERR_INVALID, ERR_UNSUPPORTED = 0, 1, 2

CTRL_NONE, CTRL_SINGLE, CTRL_ONES, CTRL_MIXED = 0, 1, 2, 3


class Op(C.Structure):
    """Same layout as spz_op (include/spinoza_b200.h) / orc_op."""

    _fields_ = [
        ("kind", C.c_int32),
        ("target", C.c_int32),
        ("t0", C.c_int32),
        ("t1", C.c_int32),
        ("p", C.c_double * 3),
        ("ctrl_kind", C.c_int32),
        ("reserved", C.c_int32),
        ("ctrl_mask", C.c_uint64),
        ("zeros_mask", C.c_uint64),
    ]


def build(force: bool = False) -> Path:
    src = _HERE / "spinoza_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_get_threads.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_state_init.argtypes = [dp, dp, C.c_int]
        L.orc_u_scalars.argtypes = [C.c_double] * 3 + [dp]
        L.orc_swap.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int]
        L.orc_apply.argtypes = [C.c_int, dp, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int]
        L.orc_c_apply.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_int, C.c_int]
        L.orc_mc_mask.argtypes = [ip, C.c_int, ip, C.c_int]
        L.orc_mc_mask.restype = C.c_uint64
        L.orc_mc_apply_mask.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_uint64, C.c_int]
        L.orc_mc_apply.argtypes = [C.c_int, dp, dp, dp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int]
        L.orc_cc_apply.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_mc_scan_literal.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_uint64, C.c_int]
        L.orc_iqft.argtypes = [dp, dp, C.c_int, ip, C.c_int]
        L.orc_prob0.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_prob0.restype = C.c_double
        L.orc_norm2.argtypes = [dp, dp, C.c_int]
        L.orc_norm2.restype = C.c_double
        L.orc_qubit_expectation_value.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_qubit_expectation_value.restype = C.c_double
        L.orc_measure_qubit.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, dp]
        L.orc_xyz_expectation_value.argtypes = [C.c_char, dp, dp, C.c_int, ip, C.c_int, dp]
        L.orc_uniforms.argtypes = [C.c_uint64, C.c_int64, dp]
        L.orc_reservoir_sampling.argtypes = [dp, dp, C.c_int, C.c_int64, C.c_int64, C.c_uint64,
                                             C.POINTER(C.c_int64)]
        L.orc_sample_cdf.argtypes = [dp, dp, C.c_int, dp, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_gen_random_state.argtypes = [dp, dp, C.c_int, C.c_uint64]
        L.orc_execute.argtypes = [dp, dp, C.c_int, C.POINTER(Op), C.c_int, C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_uint64), dp, C.c_int]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ia(xs):
    arr = (C.c_int * max(len(xs), 1))(*xs)
    return arr


def _params(p):
    p = list(p) + [0.0] * (3 - len(p))
    return (C.c_double * 3)(*p)


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__({1: "invalid argument", 2: "unsupported (reference panics: todo!/unimplemented!)"}.get(code, str(code)))
        self.code = code


def _chk(rc):
    if rc != OK:
        raise OracleError(rc)


def set_threads(t: int):
    lib().orc_set_threads(int(t))


def max_threads() -> int:
    return lib().orc_max_threads()


class State:
    """Mirror of spinoza::core::State (core.rs:18-51): split re/im f64 arrays, |0..0> on creation."""

    def __init__(self, n: int, reals=None, imags=None):
        assert n > 0  # core.rs:33
        self.n = n
        if reals is None:
            self.reals = np.zeros(1 << n, dtype=np.float64)
            self.imags = np.zeros(1 << n, dtype=np.float64)
            self.reals[0] = 1.0
        else:
            self.reals = np.ascontiguousarray(reals, dtype=np.float64).copy()
            self.imags = np.ascontiguousarray(imags, dtype=np.float64).copy()
            assert self.reals.shape == (1 << n,) and self.imags.shape == (1 << n,)

    def __len__(self):
        return 1 << self.n

    def clone(self):
        return State(self.n, self.reals, self.imags)

    def amps(self):
        return self.reals + 1j * self.imags


def gen_random_state(n: int, seed: int) -> State:
    s = State(n)
    lib().orc_gen_random_state(_dp(s.reals), _dp(s.imags), n, seed)
    return s


def apply(kind, state: State, target: int, params=(), t0=0, t1=0):
    _chk(lib().orc_apply(kind, _params(params), t0, t1, _dp(state.reals), _dp(state.imags), state.n, target))


def c_apply(kind, state: State, control: int, target: int, params=()):
    _chk(lib().orc_c_apply(kind, _params(params), _dp(state.reals), _dp(state.imags), state.n, control, target))


def cc_apply(kind, state: State, c0: int, c1: int, target: int, params=()):
    _chk(lib().orc_cc_apply(kind, _params(params), _dp(state.reals), _dp(state.imags), state.n, c0, c1, target))


def mc_apply(kind, state: State, controls, zeros, target: int, params=()):
    zs = sorted(zeros) if zeros else []
    _chk(lib().orc_mc_apply(kind, _params(params), _dp(state.reals), _dp(state.imags), state.n,
                            _ia(list(controls)), len(controls), _ia(zs), len(zs), target))


def mc_apply_mask(kind, state: State, mask: int, target: int, params=()):
    _chk(lib().orc_mc_apply_mask(kind, _params(params), _dp(state.reals), _dp(state.imags), state.n, mask, target))


def mc_scan_literal(kind, state: State, mask: int, target: int, params=()) -> int:
    """Literal scan-and-skip loop of the reference (bounds-checked). Returns the status code."""
    return lib().orc_mc_scan_literal(kind, _params(params), _dp(state.reals), _dp(state.imags), state.n, mask, target)


def swap(state: State, t0: int, t1: int):
    _chk(lib().orc_swap(_dp(state.reals), _dp(state.imags), state.n, t0, t1))


def iqft(state: State, targets):
    _chk(lib().orc_iqft(_dp(state.reals), _dp(state.imags), state.n, _ia(list(targets)), len(targets)))


def prob0(state: State, target: int) -> float:
    return lib().orc_prob0(_dp(state.reals), _dp(state.imags), state.n, target)


def norm2(state: State) -> float:
    return lib().orc_norm2(_dp(state.reals), _dp(state.imags), state.n)


def qubit_expectation_value(state: State, target: int) -> float:
    return lib().orc_qubit_expectation_value(_dp(state.reals), _dp(state.imags), state.n, target)


def measure_qubit(state: State, target: int, reset: bool, v=None, u01: float = 0.5):
    p0 = C.c_double()
    bit = lib().orc_measure_qubit(_dp(state.reals), _dp(state.imags), state.n, target, int(reset),
                                  -1 if v is None else int(v), u01, C.byref(p0))
    return bit, p0.value


def xyz_expectation_value(observable: str, state: State, targets):
    out = np.zeros(len(targets), dtype=np.float64)
    rc = lib().orc_xyz_expectation_value(observable.encode()[:1], _dp(state.reals), _dp(state.imags), state.n,
                                         _ia(list(targets)), len(targets), _dp(out))
    _chk(rc)
    return out


def uniforms(seed: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.float64)
    lib().orc_uniforms(seed, count, _dp(out))
    return out


def reservoir_sampling(state: State, k: int, num_tests: int, seed: int) -> np.ndarray:
    out = np.zeros(k, dtype=np.int64)
    lib().orc_reservoir_sampling(_dp(state.reals), _dp(state.imags), state.n, k, num_tests, seed,
                                 out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def sample_cdf(state: State, u: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros(len(u), dtype=np.int64)
    lib().orc_sample_cdf(_dp(state.reals), _dp(state.imags), state.n, _dp(u), len(u),
                         out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def make_op(kind, target=0, params=(), ctrl_kind=CTRL_NONE, ctrl_mask=0, zeros_mask=0, t0=0, t1=0) -> Op:
    op = Op()
    op.kind, op.target, op.t0, op.t1 = kind, target, t0, t1
    for i, v in enumerate(list(params)[:3]):
        op.p[i] = v
    op.ctrl_kind, op.ctrl_mask, op.zeros_mask = ctrl_kind, ctrl_mask, zeros_mask
    return op


def execute(state: State, ops, measured: int = 0, vals: int = 0, u01=()):
    """QuantumCircuit::execute (circuit.rs:552-600). Returns (measured_mask, measured_vals)."""
    arr = (Op * max(len(ops), 1))(*ops)
    m, v = C.c_uint64(measured), C.c_uint64(vals)
    u = np.ascontiguousarray(np.asarray(list(u01) + [0.5], dtype=np.float64))
    rc = lib().orc_execute(_dp(state.reals), _dp(state.imags), state.n, arr, len(ops), C.byref(m), C.byref(v),
                           _dp(u), len(u01))
    _chk(rc)
    return m.value, v.value
