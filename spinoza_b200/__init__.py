"""spinoza_b200 -- B200-native state-vector engine behind Spinoza's gate-application API.

This module is the Python host-side mirror of the reference's public surface for the hot path
(`State`, `Gate`, `apply`, `c_apply`, `cc_apply`, `mc_apply`, `iqft`, `QuantumCircuit`,
`measure_qubit`, `qubit_expectation_value`, `xyz_expectation_value`, sampling), written over the
C ABI of ``libspinoza_b200.so`` (include/spinoza_b200.h) with ctypes.  Names, argument order and
error behaviour follow /root/reference/spinoza/src/{gates,core,circuit,measurement}.rs and
/root/reference/spynoza/src/lib.rs (cited per item).

There is no CPU fallback: if the CUDA library is missing the import fails; if no GPU is visible,
creating a `State` raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path
from typing import Iterable, List, Optional, Sequence

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "lib" / "libspinoza_b200.so"

PI = math.pi  # math.rs:8


class SpinozaError(RuntimeError):
    """Raised where the reference would panic (todo!/unimplemented!/assert!)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[spz status {status}] {message}")
        self.status = status


OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_OOM, ERR_COMM, ERR_NO_DEVICE = range(7)


def _load() -> C.CDLL:
    """Load the CUDA engine; rebuild it in-tree first when its sources changed.  Never falls back."""
    # see abi.cu: kernels that spin on peer flags must not meet CUDA's lazy module loading
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    if os.environ.get("SPINOZA_B200_NO_AUTOBUILD"):
        if not _LIB_PATH.exists():
            raise ImportError(f"{_LIB_PATH} is missing: build it with `python -m spinoza_b200._build` "
                              "(there is no CPU fallback)")
    else:
        from . import _build
        try:
            _build.build()
        except Exception as e:  # no nvcc on this host: an existing in-tree library is still the product
            if not _LIB_PATH.exists():
                raise ImportError(f"cannot build {_LIB_PATH}: {e} (there is no CPU fallback)") from e
    return C.CDLL(str(_LIB_PATH))


_lib = _load()


class _Gate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("t0", C.c_int32), ("t1", C.c_int32), ("reserved", C.c_int32),
                ("p", C.c_double * 3)]


class _DistAction(C.Structure):
    _fields_ = [("type", C.c_int32), ("kind", C.c_int32), ("target", C.c_int32), ("hi", C.c_int32),
                ("cmask", C.c_uint64), ("gbit", C.c_int32), ("lq", C.c_int32), ("partner", C.c_int32),
                ("grefs", C.c_int32), ("p", C.c_double * 3)]


class _Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("target", C.c_int32), ("t0", C.c_int32), ("t1", C.c_int32),
                ("p", C.c_double * 3), ("ctrl_kind", C.c_int32), ("reserved", C.c_int32),
                ("ctrl_mask", C.c_uint64), ("zeros_mask", C.c_uint64)]


_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes).  tests/test_abi_symbols.py checks this table against include/spinoza_b200.h.
_SIGNATURES = {
    "spz_abi_version": (C.c_int, []),
    "spz_last_error": (C.c_char_p, []),
    "spz_device_count": (C.c_int, []),
    "spz_status_string": (C.c_char_p, [C.c_int]),
    "spz_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "spz_destroy": (C.c_int, [_vp]),
    "spz_clone": (C.c_int, [_vp, C.POINTER(_vp)]),
    "spz_num_qubits": (C.c_int, [_vp]),
    "spz_len": (C.c_int64, [_vp]),
    "spz_reset_zero": (C.c_int, [_vp]),
    "spz_set_basis": (C.c_int, [_vp, C.c_uint64]),
    "spz_init_random": (C.c_int, [_vp, C.c_uint64]),
    "spz_upload": (C.c_int, [_vp, _dp, _dp, C.c_int64, C.c_int64]),
    "spz_upload_async": (C.c_int, [_vp, _dp, _dp]),
    "spz_download": (C.c_int, [_vp, _dp, _dp, C.c_int64, C.c_int64]),
    "spz_sync": (C.c_int, [_vp]),
    "spz_alloc_host": (C.c_int, [C.c_uint64, C.POINTER(_vp)]),
    "spz_free_host": (C.c_int, [_vp]),
    "spz_apply": (C.c_int, [_vp, C.POINTER(_Gate), C.c_int]),
    "spz_c_apply": (C.c_int, [_vp, C.POINTER(_Gate), C.c_int, C.c_int]),
    "spz_cc_apply": (C.c_int, [_vp, C.POINTER(_Gate), C.c_int, C.c_int, C.c_int]),
    "spz_mc_apply": (C.c_int, [_vp, C.POINTER(_Gate), _i32p, C.c_int, _i32p, C.c_int, C.c_int]),
    "spz_mc_apply_mask": (C.c_int, [_vp, C.POINTER(_Gate), C.c_uint64, C.c_int]),
    "spz_mc_apply_signed": (C.c_int, [_vp, C.POINTER(_Gate), C.c_uint64, C.c_uint64, C.c_int]),
    "spz_iqft": (C.c_int, [_vp, _i32p, C.c_int]),
    "spz_execute": (C.c_int, [_vp, C.POINTER(_Op), C.c_int64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "spz_set_seed": (C.c_int, [_vp, C.c_uint64]),
    "spz_plan_fusion": (C.c_int, [C.c_int, C.POINTER(_Op), C.c_int64, C.c_uint32, _i32p, _i32p, _i32p]),
    "spz_debug_compile_sharded": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_Op), C.c_int64, C.c_uint32, C.c_void_p, C.c_int64,
                                            C.POINTER(C.c_int64)]),
    "spz_debug_compile_pass": (C.c_int, [C.c_int, C.POINTER(_Op), C.c_int64, C.c_uint32, C.c_int, C.c_void_p, C.c_int64,
                                         C.POINTER(C.c_int64)]),
    "spz_prob0": (C.c_int, [_vp, C.c_int, _dp]),
    "spz_norm2": (C.c_int, [_vp, _dp]),
    "spz_measure_qubit": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "spz_qubit_expectation_value": (C.c_int, [_vp, C.c_int, _dp]),
    "spz_xyz_expectation_value": (C.c_int, [_vp, C.c_char, _i32p, C.c_int, _dp]),
    "spz_sample": (C.c_int, [_vp, _dp, C.c_int64, C.POINTER(C.c_int64)]),
    "spz_dist_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "spz_dist_export": (C.c_int, [_vp, C.c_char_p]),
    "spz_dist_connect": (C.c_int, [_vp, C.c_char_p]),
    "spz_dist_connect_local": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "spz_dist_perm": (C.c_int, [_vp, _i32p]),
    "spz_dist_local_qubits": (C.c_int, [_vp]),
    "spz_dist_copy_from": (C.c_int, [_vp, _vp]),
    "spz_dist_stats": (C.c_int, [_vp, _dp]),
    "spz_rdv_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "spz_rdv_allgather": (C.c_int, [_vp, C.c_void_p, C.c_int64, C.c_void_p]),
    "spz_rdv_barrier": (C.c_int, [_vp]),
    "spz_rdv_close": (C.c_int, [_vp]),
    "spz_dist_connect_rdv": (C.c_int, [_vp, _vp]),
    "spz_dist_plan_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "spz_dist_plan_destroy": (C.c_int, [_vp]),
    "spz_dist_plan_lower": (C.c_int, [_vp, C.c_int, C.POINTER(_Op), C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "spz_dist_plan_perm": (C.c_int, [_vp, _i32p]),
    "spz_timer_start": (C.c_int, [_vp]),
    "spz_timer_stop": (C.c_int, [_vp, _dp]),
    "spz_launch_count": (C.c_int64, []),
    "spz_device_name": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "spz_mem_info": (C.c_int, [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(_lib, _name)  # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def _check(status: int):
    if status != OK:
        raise SpinozaError(status, _lib.spz_last_error().decode(errors="replace"))


def library_path() -> str:
    return str(_LIB_PATH)


def device_count() -> int:
    return _lib.spz_device_count()


def device_name(device: int = 0) -> str:
    buf = C.create_string_buffer(256)
    _check(_lib.spz_device_name(device, buf, 256))
    return buf.value.decode()


def mem_info(device: int = 0):
    f, t = C.c_uint64(), C.c_uint64()
    _check(_lib.spz_mem_info(device, C.byref(f), C.byref(t)))
    return f.value, t.value


def launch_count() -> int:
    return _lib.spz_launch_count()


# ---- Gate enum (gates.rs:44-74) -------------------------------------------------------------------
class Gate:
    """`Gate::H`, `Gate::P(theta)`, ... -> `Gate.H`, `Gate.P(theta)`, ..."""

    KIND_H, KIND_M, KIND_X, KIND_Y, KIND_Z, KIND_P, KIND_RX, KIND_RY, KIND_RZ, KIND_SWAP, KIND_U, KIND_UNITARY, \
        KIND_BITFLIP = range(13)
    _NAMES = ["H", "M", "X", "Y", "Z", "P", "RX", "RY", "RZ", "SWAP", "U", "Unitary", "BitFlipNoise"]

    __slots__ = ("kind", "params", "t0", "t1")

    def __init__(self, kind: int, params: Sequence[float] = (), t0: int = 0, t1: int = 0):
        self.kind, self.params, self.t0, self.t1 = kind, tuple(float(x) for x in params), int(t0), int(t1)

    def __repr__(self):
        args = ", ".join(map(repr, self.params)) if self.kind != Gate.KIND_SWAP else f"{self.t0}, {self.t1}"
        return f"Gate.{self._NAMES[self.kind]}" + (f"({args})" if args else "")

    def __eq__(self, other):
        return isinstance(other, Gate) and (self.kind, self.params, self.t0, self.t1) == (
            other.kind, other.params, other.t0, other.t1)

    def inverse(self) -> "Gate":
        """Gate::inverse gates.rs:78-92."""
        k = self.kind
        if k in (Gate.KIND_H, Gate.KIND_X, Gate.KIND_Y, Gate.KIND_Z, Gate.KIND_SWAP):
            return self
        if k in (Gate.KIND_P, Gate.KIND_RX, Gate.KIND_RY, Gate.KIND_RZ):
            return Gate(k, (-self.params[0],))
        if k == Gate.KIND_U:
            theta, phi, lam = self.params
            return Gate(k, (-theta, -lam, -phi))
        raise SpinozaError(ERR_UNSUPPORTED, f"{self!r}.inverse(): unimplemented!() (gates.rs:86)")

    def to_matrix(self) -> np.ndarray:
        """Gate::to_matrix gates.rs:95-190 (row-major 2x2)."""
        k, p = self.kind, self.params
        r = math.sqrt(0.5)
        if k == Gate.KIND_H:
            return np.array([[r, r], [r, -r]], dtype=complex)
        if k == Gate.KIND_X:
            return np.array([[0, 1], [1, 0]], dtype=complex)
        if k == Gate.KIND_Y:
            return np.array([[0, -1j], [1j, 0]], dtype=complex)
        if k == Gate.KIND_Z:
            return np.array([[1, 0], [0, -1]], dtype=complex)
        if k == Gate.KIND_P:
            return np.array([[1, 0], [0, complex(math.cos(p[0]), math.sin(p[0]))]], dtype=complex)
        if k == Gate.KIND_RX:
            c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
            return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
        if k == Gate.KIND_RY:
            c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
            return np.array([[c, -s], [s, c]], dtype=complex)
        if k == Gate.KIND_RZ:
            c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
            return np.array([[complex(c, -s), 0], [0, complex(c, s)]], dtype=complex)
        if k == Gate.KIND_U:
            th, ph, la = p
            c, s = math.cos(th / 2), math.sin(th / 2)
            return np.array([[c, complex(-math.cos(la) * s, -math.sin(la) * s)],
                             [complex(math.cos(ph) * s, math.sin(ph) * s),
                              complex(math.cos(ph + la) * c, math.sin(ph + la) * c)]], dtype=complex)
        raise SpinozaError(ERR_UNSUPPORTED, f"{self!r}.to_matrix(): unimplemented!() (gates.rs:188)")

    def _c(self) -> _Gate:
        g = _Gate()
        g.kind, g.t0, g.t1 = self.kind, self.t0, self.t1
        for i, v in enumerate(self.params[:3]):
            g.p[i] = v
        return g

    # constructors mirroring the enum variants
    @staticmethod
    def P(theta: float) -> "Gate":
        return Gate(Gate.KIND_P, (theta,))

    @staticmethod
    def RX(theta: float) -> "Gate":
        return Gate(Gate.KIND_RX, (theta,))

    @staticmethod
    def RY(theta: float) -> "Gate":
        return Gate(Gate.KIND_RY, (theta,))

    @staticmethod
    def RZ(theta: float) -> "Gate":
        return Gate(Gate.KIND_RZ, (theta,))

    @staticmethod
    def U(theta: float, phi: float, lam: float) -> "Gate":
        return Gate(Gate.KIND_U, (theta, phi, lam))

    @staticmethod
    def SWAP(t0: int, t1: int) -> "Gate":
        return Gate(Gate.KIND_SWAP, (), t0, t1)

    @staticmethod
    def BitFlipNoise(prob: float) -> "Gate":
        return Gate(Gate.KIND_BITFLIP, (prob,))


Gate.H = Gate(Gate.KIND_H)
Gate.M = Gate(Gate.KIND_M)
Gate.X = Gate(Gate.KIND_X)
Gate.Y = Gate(Gate.KIND_Y)
Gate.Z = Gate(Gate.KIND_Z)


class HostBuffer:
    """Page-locked host array of f64 (cudaMallocHost) exposed as a NumPy view: `buf.array`."""

    def __init__(self, count: int):
        p = _vp()
        _check(_lib.spz_alloc_host(8 * int(count), C.byref(p)))
        self._p = p
        self.count = int(count)
        self.array = np.ctypeslib.as_array(C.cast(p, _dp), shape=(self.count,))

    def ptr(self):
        return C.cast(self._p, _dp)

    def __del__(self):
        p, self._p = getattr(self, "_p", None), None
        if p is not None and _lib is not None:
            self.array = None
            _lib.spz_free_host(p)


def _as_f64(a, n):
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if arr.shape != (n,):
        raise ValueError(f"expected {n} values, got shape {arr.shape}")
    return arr


# ---- State (core.rs:18-51) --------------------------------------------------------------------------
class State:
    """Device-resident `State { reals, imags, n }`.

    Unavoidable deviation from the reference: `reals` / `imags` are not `Vec` fields
    but download the arrays from HBM; `amp(i)` reads one amplitude.
    """

    def __init__(self, n: int, device: int = 0, _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            h = _vp()
            _check(_lib.spz_create(int(n), int(device), C.byref(h)))  # State::new core.rs:32-42
            self._h = h
        self.n = _lib.spz_num_qubits(self._h)
        self.device = device

    @classmethod
    def from_arrays(cls, reals, imags, device: int = 0) -> "State":
        reals = np.asarray(reals)
        n = int(reals.shape[0]).bit_length() - 1
        if reals.shape[0] != 1 << n:
            raise ValueError("state length must be a power of two")
        st = cls(n, device)
        st.upload(reals, imags)
        return st

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.spz_destroy(h)

    def __len__(self) -> int:  # State::len core.rs:48
        return int(_lib.spz_len(self._h))

    def len(self) -> int:
        return len(self)

    def clone(self) -> "State":  # #[derive(Clone)]
        h = _vp()
        _check(_lib.spz_clone(self._h, C.byref(h)))
        return State(self.n, self.device, _handle=h)

    def upload(self, reals, imags, offset: int = 0):
        reals = np.ascontiguousarray(reals, dtype=np.float64)
        imags = np.ascontiguousarray(imags, dtype=np.float64)
        if reals.shape != imags.shape or reals.ndim != 1:
            raise ValueError("reals/imags must be equal-length 1-D arrays")
        _check(_lib.spz_upload(self._h, reals.ctypes.data_as(_dp), imags.ctypes.data_as(_dp), offset, reals.shape[0]))

    def download(self, offset: int = 0, count: Optional[int] = None):
        count = len(self) - offset if count is None else count
        re = np.empty(count, dtype=np.float64)
        im = np.empty(count, dtype=np.float64)
        _check(_lib.spz_download(self._h, re.ctypes.data_as(_dp), im.ctypes.data_as(_dp), offset, count))
        return re, im

    def upload_from(self, re: "HostBuffer", im: "HostBuffer", offset: int = 0):
        _check(_lib.spz_upload(self._h, re.ptr(), im.ptr(), offset, re.count))

    def upload_async(self, re: "HostBuffer", im: "HostBuffer"):
        """Whole-state upload from page-locked buffers that returns at once: the gates issued next follow the pieces as they
        arrive.  The buffers must stay alive and untouched until the next sync() / download."""
        if re.count != len(self) or im.count != len(self):
            raise ValueError("upload_async moves the whole state: buffers of len(state) doubles")
        self._async_src = (re, im)  # keep the buffers alive
        _check(_lib.spz_upload_async(self._h, re.ptr(), im.ptr()))

    def download_into(self, re: "HostBuffer", im: "HostBuffer", offset: int = 0):
        _check(_lib.spz_download(self._h, re.ptr(), im.ptr(), offset, re.count))

    @property
    def reals(self) -> np.ndarray:
        return self.download()[0]

    @property
    def imags(self) -> np.ndarray:
        return self.download()[1]

    def amp(self, i: int) -> complex:
        re, im = self.download(i, 1)
        return complex(re[0], im[0])

    def amps(self) -> np.ndarray:
        re, im = self.download()
        return re + 1j * im

    def reset(self):
        _check(_lib.spz_reset_zero(self._h))

    def set_basis(self, index: int):
        _check(_lib.spz_set_basis(self._h, index))

    def init_random(self, seed: int):
        """utils.rs:168-201 gen_random_state, generated on the device from a counter-based RNG."""
        _check(_lib.spz_init_random(self._h, seed))

    def set_seed(self, seed: int):
        _check(_lib.spz_set_seed(self._h, seed))

    def sync(self):
        _check(_lib.spz_sync(self._h))

    def timer_start(self):
        _check(_lib.spz_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        _check(_lib.spz_timer_stop(self._h, C.byref(ms)))
        return ms.value


def _i32(xs: Iterable[int]):
    xs = list(xs)
    return (C.c_int32 * max(len(xs), 1))(*xs), len(xs)


# ---- gates.rs:215-320 ------------------------------------------------------------------------------------
def apply(gate: Gate, state: State, target: int):
    """`apply(gate, &mut state, target)` gates.rs:215."""
    g = gate._c()
    _check(_lib.spz_apply(state._h, C.byref(g), target))


def c_apply(gate: Gate, state: State, control: int, target: int):
    """`c_apply(gate, &mut state, control, target)` gates.rs:257."""
    g = gate._c()
    _check(_lib.spz_c_apply(state._h, C.byref(g), control, target))


def cc_apply(gate: Gate, state: State, control0: int, control1: int, target: int):
    """`cc_apply(gate, &mut state, control0, control1, target)` gates.rs:272."""
    g = gate._c()
    _check(_lib.spz_cc_apply(state._h, C.byref(g), control0, control1, target))


def mc_apply(gate: Gate, state: State, controls: Sequence[int], zeros: Optional[Iterable[int]], target: int):
    """`mc_apply(gate, &mut state, controls, zeros, target)` gates.rs:290."""
    g = gate._c()
    cs, nc = _i32(controls)
    if zeros is None:
        _check(_lib.spz_mc_apply(state._h, C.byref(g), cs, nc, None, 0, target))
    else:
        zs, nz = _i32(sorted(zeros))
        _check(_lib.spz_mc_apply(state._h, C.byref(g), cs, nc, zs, nz, target))


def mc_apply_mask(gate: Gate, state: State, ctrl_mask: int, target: int):
    g = gate._c()
    _check(_lib.spz_mc_apply_mask(state._h, C.byref(g), ctrl_mask, target))


def mc_apply_signed(gate: Gate, state: State, ones: Iterable[int], zeros: Iterable[int], target: int):
    """Extension (not in the reference): `gate` on `target` where every qubit of `ones` is 1 and every qubit of `zeros` is 0 --
    the negative controls `Controls::Mixed { zeros }` describes but mc_apply drops (gates.rs:298-311).  One pass."""
    g = gate._c()
    om = sum(1 << int(q) for q in set(ones))
    zm = sum(1 << int(q) for q in set(zeros))
    _check(_lib.spz_mc_apply_signed(state._h, C.byref(g), om, zm, target))


def iqft(state: State, targets: Sequence[int]):
    """`iqft(&mut state, targets)` core.rs:184."""
    ts, n = _i32(targets)
    _check(_lib.spz_iqft(state._h, ts, n))


# ---- measurement.rs / core.rs reductions -----------------------------------------------------------------
def measure_qubit(state: State, target: int, reset: bool, v: Optional[int] = None) -> int:
    """`measure_qubit(&mut state, target, reset, v)` measurement.rs:12."""
    bit = C.c_int()
    _check(_lib.spz_measure_qubit(state._h, target, int(reset), -1 if v is None else int(v), C.byref(bit)))
    return bit.value


def prob0(state: State, target: int) -> float:
    out = C.c_double()
    _check(_lib.spz_prob0(state._h, target, C.byref(out)))
    return out.value


def norm2(state: State) -> float:
    out = C.c_double()
    _check(_lib.spz_norm2(state._h, C.byref(out)))
    return out.value


def qubit_expectation_value(state: State, target: int) -> float:
    """core.rs:198."""
    out = C.c_double()
    _check(_lib.spz_qubit_expectation_value(state._h, target, C.byref(out)))
    return out.value


def xyz_expectation_value(observable: str, state: State, targets: Sequence[int]) -> List[float]:
    """core.rs:222."""
    ts, n = _i32(targets)
    out = (C.c_double * max(n, 1))()
    _check(_lib.spz_xyz_expectation_value(state._h, observable.encode()[:1] or b"?", ts, n, out))
    return [out[i] for i in range(n)]


def uniforms(seed: int, count: int) -> np.ndarray:
    """splitmix64 stream -> [0,1): the repo's shared deterministic source of host randomness."""
    mask = (1 << 64) - 1
    out = np.empty(count, dtype=np.float64)
    x = seed & mask
    for i in range(count):
        x = (x + 0x9E3779B97F4A7C15) & mask
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        z ^= z >> 31
        out[i] = (z >> 11) * (1.0 / 9007199254740992.0)
    return out


def sample(state: State, shots: int, seed: int = 0, u01: Optional[np.ndarray] = None) -> np.ndarray:
    """Exact inverse-CDF sampling of `shots` basis states (replaces reservoir_sampling core.rs:125)."""
    if u01 is None:
        u01 = np.random.default_rng(seed).random(shots)
    u01 = np.ascontiguousarray(u01, dtype=np.float64)
    out = np.empty(len(u01), dtype=np.int64)
    _check(_lib.spz_sample(state._h, u01.ctypes.data_as(_dp), len(u01), out.ctypes.data_as(C.POINTER(C.c_int64))))
    return out


class Reservoir:
    """`spinoza::core::Reservoir` (core.rs:65-121): same construction and read-out.  The filling is the engine's exact
    inverse-CDF sampler (spz_sample: one read pass over the device state) instead of `num_tests` rounds of weighted
    replacement, so every entry is an exact draw from |amplitude|^2 whatever `num_tests` is."""

    def __init__(self, k: int, seed: int = 0x9E3779B97F4A7C15):
        self.entries = [0] * k
        self._seed = seed

    def sampling(self, state: State, num_tests: int):  # core.rs:96-113
        u = np.random.default_rng(self._seed).random(len(self.entries))
        self._seed += 1
        self.entries = [int(i) for i in sample(state, len(u), u01=u)]

    def get_outcome_count(self) -> dict:  # core.rs:115-121
        out: dict = {}
        for e in self.entries:
            out[e] = out.get(e, 0) + 1
        return out


def reservoir_sampling(state: State, reservoir_size: int, num_tests: int) -> Reservoir:
    """`reservoir_sampling(&state, reservoir_size, num_tests)` core.rs:125-129."""
    r = Reservoir(reservoir_size)
    r.sampling(state, num_tests)
    return r


from .circuit import (Controls, QuantumCircuit, QuantumRegister, QuantumTransformation,  # noqa: E402
                      EXEC_FUSE, EXEC_NO_FUSE, EXEC_EXACT)
from . import openqasm  # noqa: E402
from . import distributed  # noqa: E402

__all__ = [
    "PI", "SpinozaError", "Gate", "State", "HostBuffer", "apply", "c_apply", "cc_apply", "mc_apply", "mc_apply_mask", "mc_apply_signed", "iqft",
    "measure_qubit", "prob0", "norm2", "qubit_expectation_value", "xyz_expectation_value", "sample", "uniforms", "Reservoir", "reservoir_sampling",
    "Controls", "QuantumCircuit", "QuantumRegister", "QuantumTransformation", "EXEC_FUSE", "EXEC_NO_FUSE", "EXEC_EXACT",
    "openqasm", "device_count", "device_name", "mem_info", "launch_count", "library_path",
]
