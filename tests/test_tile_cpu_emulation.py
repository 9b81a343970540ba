"""The bodies of the fused tile kernels -- k_tile3 (csrc/kernels_tile3.cu: the TMA kernel that runs every merged-mode pass on
full 12-bit tiles) and k_tile (csrc/kernels_tile.cu: exact mode, registers below 12 qubits, oversized programs) -- executed on
the CPU: tests/emu/ compiles the kernel sources themselves with g++ (one OS thread per CUDA thread, a barrier for
__syncthreads, the TMA box copies as synchronous copies under the same 128-byte swizzle) and runs them block by block on the
micro-programs that spz_execute would upload (spz_debug_compile_pass), for k_tile3 through the host lowering the launcher uses
(tile3_lower).  Unlike tests/test_tile_program.py -- an independent NumPy statement of the execution model -- this exercises
the real code: swizzled register layouts, the lowered instruction format, table-driven phase accumulators, rescaled
butterflies, exact-mode diagonal arithmetic.

It does not replace the GPU parity tests (no warps, no TMA engine, no memory model, no timing); it lets the CPU suite
(-m "not gpu") guard the logic of both kernels.
"""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import Gate, QuantumCircuit, QuantumRegister, workloads
from tests import _dense as D
from tests.test_gpu_parity import oracle_ops_from
from tests.test_scheduler_plan import random_circuit, run_dense_order
from tests.test_tile_program import compile_pass

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
    lib = out / "libtile_emu.so"
    cmd = [gxx, "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-pthread", f"-I{CUDA_INC}",
           "-include", str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "tile_emu.cpp"), "-o", str(lib)]
    subprocess.run(cmd, check=True, cwd=ROOT)
    h = C.CDLL(str(lib))
    h.emu_tile3_run.restype = C.c_int
    h.emu_tile3_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong, C.POINTER(C.c_int)]
    h.emu_tile1_run.restype = C.c_int
    h.emu_tile1_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong, C.c_int, C.c_int]
    return h


def raw_pass(qc, pass_index):
    arr, n = qc._encode()
    buf = (C.c_char * (8 << 20))()
    used = C.c_int64()
    sb._check(sb._lib.spz_debug_compile_pass(qc.n_qubits, arr, n, qc._flags(), pass_index, buf, len(buf), C.byref(used)))
    return bytes(buf[: used.value])


def emu_pass(emu, kernel, n, re, im, blob, exact, option=1, info=None):
    """One fused pass through the emulated kernel.  kernel 1 = k_tile (option: program decoded from shared memory 1 / global
    memory 0), kernel 3 = k_tile3 (merged mode only).  Returns 1 when k_tile3 is not eligible (the launcher uses k_tile)."""
    if kernel == 1:
        return emu.emu_tile1_run(n, re.ctypes.data, im.ctypes.data, blob, len(blob), 1 if exact else 0, option)
    assert not exact, "k_tile3 runs merged mode only"
    info = info if info is not None else (C.c_int * 4)()
    return emu.emu_tile3_run(n, re.ctypes.data, im.ctypes.data, blob, len(blob), info)


def run_emulated(emu, qc, re, im, stats, kernel=3):
    """Every fused pass through the emulated kernel; single-op passes through the dense statement."""
    n = qc.n_qubits
    trs = list(qc.transformations)
    plan, n_pass = qc.plan()
    by_pass = {}
    for idx, p in plan:
        by_pass.setdefault(p, []).append(idx)
    for p in range(n_pass):
        blob = raw_pass(qc, p)
        status = int(np.frombuffer(blob, dtype="<i4", count=1)[0])
        if status == 1:
            psi = run_dense_order(n, re + 1j * im, trs, by_pass[p])
            re[:], im[:] = psi.real, psi.imag
            stats["direct"] = stats.get("direct", 0) + 1
            continue
        info = (C.c_int * 4)()
        rc = emu_pass(emu, kernel, n, re, im, blob, qc.exact, (p & 1), info)
        if kernel == 1:
            assert rc == 0, f"pass {p}: emulation failed (rc={rc})"
            stats["k_tile"] = stats.get("k_tile", 0) + 1
            continue
        if rc == 1:  # too long for k_tile3's shared-memory budget: the launcher falls back to k_tile
            assert emu_pass(emu, 1, n, re, im, blob, qc.exact, 1) == 0
            stats["fallback"] = stats.get("fallback", 0) + 1
            continue
        assert rc == 0, f"pass {p}: emulation failed (rc={rc})"
        key = "ctrl" if info[0] else "noctrl"
        stats[key] = stats.get(key, 0) + 1
        stats["groups"] = stats.get("groups", 0) + info[2]
    return re, im


def reference_cells_circuit(n, count, seed, **kw):
    """Random circuit restricted to the gate x control cells the reference (and therefore the oracle) supports
    (c_apply has no Z; mc_apply only X, P, RX, RY)."""
    rng = np.random.default_rng(seed)
    one_q = [Gate.KIND_H, Gate.KIND_X, Gate.KIND_Y, Gate.KIND_Z, Gate.KIND_P, Gate.KIND_RX, Gate.KIND_RY, Gate.KIND_RZ, Gate.KIND_U]
    c_ok = [k for k in one_q if k != Gate.KIND_Z]
    mc_ok = [Gate.KIND_X, Gate.KIND_P, Gate.KIND_RX, Gate.KIND_RY]
    qc = QuantumCircuit(QuantumRegister(n), **kw)
    for _ in range(count):
        r = rng.random()
        t = int(rng.integers(n))
        p = tuple(float(x) for x in rng.random(3) * 2 * np.pi)
        others = [q for q in range(n) if q != t]
        if r < 0.05:
            qc.swap(int(rng.integers(n)), int(rng.integers(n)))
        elif r < 0.5:
            qc.add(sb.QuantumTransformation(Gate(one_q[int(rng.integers(len(one_q)))], p), t))
        elif r < 0.85:
            c = int(rng.choice(others))
            qc.add(sb.QuantumTransformation(Gate(c_ok[int(rng.integers(len(c_ok)))], p), t, sb.Controls.single(c)))
        else:
            cs = [int(c) for c in rng.choice(others, size=int(rng.integers(2, 4)), replace=False)]
            qc.add(sb.QuantumTransformation(Gate(mc_ok[int(rng.integers(len(mc_ok)))], p), t, sb.Controls.mixed(cs, set())))
    return qc


def dense_reference(qc, psi0):
    trs = list(qc.transformations)
    return run_dense_order(qc.n_qubits, psi0.copy(), trs, range(len(trs)))


def start(n, seed):
    psi0 = D.random_state(n, seed)
    return psi0, np.ascontiguousarray(psi0.real), np.ascontiguousarray(psi0.imag)


def test_qft_merged_mode(emu):
    n = 14
    qc = QuantumCircuit(QuantumRegister(n))
    qc.qft()
    psi0, re, im = start(n, 3)
    stats = {}
    run_emulated(emu, qc, re, im, stats)
    want = dense_reference(qc, psi0)
    np.testing.assert_allclose(re + 1j * im, want, rtol=0, atol=1e-12)
    # QFT has no in-tile controls on its butterflies, and its controlled phases reach outside the first pass's tile
    assert stats.get("noctrl", 0) >= 1 and stats.get("groups", 0) >= 1, stats


@pytest.mark.parametrize("n", [16, 18])
def test_qft_several_passes(emu, n):
    """QFT-16 / QFT-18: two passes, the first with high tile qubits and per-tile constants from the qubits below them."""
    qc = QuantumCircuit(QuantumRegister(n))
    qc.qft()
    psi0, re, im = start(n, n)
    stats = {}
    run_emulated(emu, qc, re, im, stats)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    assert stats.get("noctrl", 0) >= 2, stats


@pytest.mark.parametrize("select", ["0", "1"], ids=["first-come-tile", "chosen-tile"])
@pytest.mark.parametrize("kernel", [1, 3], ids=["k_tile", "k_tile3"])
@pytest.mark.parametrize("n,count,seed", [(13, 160, 31), (14, 220, 32), (15, 120, 33)])
def test_random_circuits_merged_mode(emu, n, count, seed, kernel, select, monkeypatch):
    monkeypatch.setenv("SPZ_TILE_SELECT", select)  # both ways of choosing a pass's tile qubits (default: chosen from 24 qubits up)
    qc = random_circuit(n, count, seed)
    psi0, re, im = start(n, seed)
    stats = {}
    run_emulated(emu, qc, re, im, stats, kernel=kernel)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    if kernel == 3:
        assert stats.get("ctrl", 0) >= 1, stats
    else:
        assert stats.get("k_tile", 0) >= 1


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_reference_cells_merged_mode(emu, seed):
    """Every gate x control cell the reference supports (incl. U, Y, multi-controlled RX / RY / P, SWAP) through k_tile3."""
    n = 13
    qc = reference_cells_circuit(n, 200, 70 + seed)
    psi0, re, im = start(n, seed)
    stats = {}
    run_emulated(emu, qc, re, im, stats)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    assert stats.get("ctrl", 0) >= 1, stats


class _SignedView:
    """A circuit with SIGNED controls as run_emulated wants it: programs compiled from the list as the caller wrote it (the
    native expansion in spz_execute), plan and dense fallback from the expanded list -- they must describe the same schedule."""

    def __init__(self, qc):
        self._qc, self._ex = qc, qc.expanded()
        self.n_qubits, self.exact = qc.n_qubits, qc.exact
        self.transformations = self._ex.transformations

    def plan(self):
        return self._ex.plan()

    def _encode(self):
        return self._qc._encode()

    def _flags(self):
        return self._qc._flags()


@pytest.mark.parametrize("kernel", [1, 3], ids=["k_tile", "k_tile3"])
def test_signed_controls_in_fused_passes(emu, kernel):
    """Negative controls (Controls.signed, the SPZ_CTRL_SIGNED extension) inside fused passes: spz_execute schedules X on the
    zero-controls around the all-ones gate; the compiled programs must reproduce the dense statement with the zeros at 0."""
    n = 13
    rng = np.random.default_rng(77)
    qc = random_circuit(n, 60, 78)
    kinds = [Gate.KIND_X, Gate.KIND_P, Gate.KIND_RX, Gate.KIND_RY, Gate.KIND_H, Gate.KIND_RZ, Gate.KIND_U, Gate.KIND_Z, Gate.KIND_Y]
    for i in range(24):
        t = int(rng.integers(n))
        others = [q for q in range(n) if q != t]
        cs = [int(c) for c in rng.choice(others, size=int(rng.integers(1, 4)), replace=False)]
        zs = {c for c in cs if rng.random() < 0.5} or {cs[-1]}
        p = tuple(float(x) for x in rng.random(3) * 2 * np.pi)
        qc.add(sb.QuantumTransformation(Gate(kinds[i % len(kinds)], p), t, sb.Controls.signed(cs, zs)))
        qc.h(int(rng.integers(n)))
    with pytest.raises(sb.SpinozaError):
        qc.plan()  # one entry per op of the caller's list cannot describe the expansion
    psi0, re, im = start(n, 9)
    stats = {}
    run_emulated(emu, _SignedView(qc), re, im, stats, kernel=kernel)
    want = run_dense_order(n, psi0.copy(), list(qc.transformations), range(len(qc.transformations)))
    np.testing.assert_allclose(re + 1j * im, want, rtol=0, atol=1e-12)
    assert stats.get("ctrl", 0) + stats.get("k_tile", 0) >= 1, stats


def test_rotations_near_pi_keep_their_matrix(emu):
    """RX / RY with |cos(theta/2)| tiny are not rescaled (the factored form divides by the cosine), and a long run of strongly
    rescaled rotations must not drive the pass scale out of range."""
    n = 12
    qc = QuantumCircuit(QuantumRegister(n))
    for t in range(n):
        qc.rx(np.pi - 1e-9 * (t + 1), t)
        qc.ry(np.pi + 1e-7 * (t + 1), t)
        qc.rx(np.pi, t)
    for rep in range(40):
        for t in range(n):
            qc.rx(np.pi - 0.02, t)   # cos(theta/2) ~ 0.01 each: 480 of them would scale by 1e-960
    psi0, re, im = start(n, 5)
    run_emulated(emu, qc, re, im, {})
    assert np.all(np.isfinite(re)) and np.all(np.isfinite(im))
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)


@pytest.mark.parametrize("n", [5, 8, 11])
def test_k_tile_small_registers_use_partial_tiles(emu, n):
    """n < 12: the tile is the whole register and the block has 2^(n-4) threads (k_tile only; k_tile3 needs full tiles)."""
    qc = random_circuit(n, 80, 60 + n)
    psi0, re, im = start(n, n)
    stats = {}
    run_emulated(emu, qc, re, im, stats, kernel=1)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    assert stats.get("k_tile", 0) >= 1


def test_qft_k_tile(emu):
    n = 14
    qc = QuantumCircuit(QuantumRegister(n))
    qc.qft()
    psi0, re, im = start(n, 4)
    run_emulated(emu, qc, re, im, {}, kernel=1)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)


@pytest.mark.parametrize("kernel", [1, 3], ids=["k_tile", "k_tile3"])
@pytest.mark.parametrize("lmin", ["4", "5"])
def test_shorter_tile_segments(emu, monkeypatch, lmin, kernel):
    """SPZ_TILE_LMIN = 4 / 5: up to 8 / 7 arbitrary high qubits per pass, segments of 16 / 32 amplitudes."""
    monkeypatch.setenv("SPZ_TILE_LMIN", lmin)
    n = 15
    qc = QuantumCircuit(QuantumRegister(n))
    workloads.random_layered_circuit(qc, depth=6, seed=11)
    for t in range(n - 1, 6, -1):   # a run of high targets so that one pass really takes more than 6 high qubits
        qc.h(t)
    widest = 0
    _, n_pass = qc.plan()
    for p in range(n_pass):
        c = compile_pass(qc, p)
        if c[0] == "tile":
            assert c[1]["L"] >= int(lmin)
            widest = max(widest, len(c[1]["high"]))
    assert widest == 12 - int(lmin)
    psi0, re, im = start(n, 21)
    run_emulated(emu, qc, re, im, {}, kernel=kernel)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)


def test_every_register_layout_shape(emu):
    """Layouts with register bit 0 = tile bit 0 move amplitude pairs with 128-bit shared-memory accesses, the others one
    amplitude at a time; both at the first and at the last layout of a pass."""
    n = 13
    for first, last in (((0, 1, 2, 3), (0, 1, 6, 7)), ((0, 2, 4, 6), (0, 3, 5, 9)), ((2, 3, 4, 5), (1, 2, 10, 11)), ((0, 1, 10, 11), (5, 6, 7, 8))):
        qc = QuantumCircuit(QuantumRegister(n))
        for t in first:
            qc.h(t)
        qc.cp(0.3, first[0], 12)
        for t in last:
            if t in first:
                qc.h(8)       # force a layout change: 8 is in neither cluster's first half
            qc.ry(0.2 + 0.1 * t, t)
        psi0, re, im = start(n, sum(first) + 7 * sum(last))
        run_emulated(emu, qc, re, im, {})
        np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)


@pytest.mark.parametrize("n,count,seed", [(13, 60, 41), (14, 80, 42)])
def test_exact_mode_is_bit_identical_to_the_oracle(emu, n, count, seed, kernel=1):
    """EXACT programs replay the reference arithmetic operation by operation; the emulation is built with
    -ffp-contract=off like the oracle, so even on the CPU the two must agree in every bit."""
    qc = reference_cells_circuit(n, count, seed, exact=True)
    init = orc.gen_random_state(n, seed)
    re, im = init.reals.copy(), init.imags.copy()
    ops = oracle_ops_from(qc)
    # single-op passes: let the oracle apply them, so that the whole chain stays bit-exact
    trs = list(qc.transformations)
    plan, n_pass = qc.plan()
    by_pass = {}
    for idx, p in plan:
        by_pass.setdefault(p, []).append(idx)
    n_tile = 0
    for p in range(n_pass):
        blob = raw_pass(qc, p)
        if int(np.frombuffer(blob, dtype="<i4", count=1)[0]) == 1:
            s = orc.State(n)
            s.reals[:], s.imags[:] = re, im
            orc.execute(s, [ops[i] for i in by_pass[p]])
            re, im = s.reals.copy(), s.imags.copy()
            continue
        info = (C.c_int * 4)()
        rc = emu_pass(emu, kernel, n, re, im, blob, True, 1, info)
        assert rc == 0
        n_tile += 1
    assert n_tile >= 1
    cpu = init.clone()
    orc.execute(cpu, ops)
    assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags)


def test_layered_circuit(emu):
    n = 13
    qc = QuantumCircuit(QuantumRegister(n))
    workloads.random_layered_circuit(qc, depth=6, seed=5)
    psi0, re, im = start(n, 9)
    run_emulated(emu, qc, re, im, {})
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)


# ---- shared-memory protocol under ThreadSanitizer ------------------------------------------------------------------
# The emulation's barrier is annotated per generation (tests/emu/cuda_cpu_shim.h), so a missing __syncthreads() between a
# shared-memory store and a load by another thread is a happens-before race that TSan reports deterministically.  k_tile3 has
# one barrier after its tables are staged (which, in the emulation, also publishes the tile thread 0 copied in; on the GPU
# that is the mbarrier the TMA boxes complete on), one per layout change between store_regs and load_regs, and one before
# the tile goes out.  The per-thread accumulator slots F[c][tid] need none.

@pytest.fixture(scope="module")
def emu_tsan():
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    out = EMU_DIR / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "tile_emu_tsan"
    cmd = [gxx, "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-w", "-fsanitize=thread", "-DSPZ_EMU_TSAN", "-DSPZ_EMU_MAIN", "-pthread",
           f"-I{CUDA_INC}", "-include", str(EMU_DIR / "cuda_cpu_shim.h"), "-x", "c++", str(EMU_DIR / "tile_emu.cpp"), "-o", str(exe)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    probe = subprocess.run([str(exe)], capture_output=True, text=True)
    if probe.returncode != 64:  # the driver's usage exit code; anything else means TSan cannot start in this sandbox
        pytest.skip("ThreadSanitizer cannot run here: " + probe.stderr[-200:])
    return exe


def tsan_run(exe, tmp_path, kernel, n, exact, re, im, blob, option=1):
    state = tmp_path / "state.bin"
    np.concatenate([re, im]).tofile(state)
    (tmp_path / "blob.bin").write_bytes(blob)
    r = subprocess.run([str(exe), str(kernel), str(n), "1" if exact else "0", str(state), str(tmp_path / "blob.bin"), str(option)],
                       capture_output=True, text=True,
                       env={"TSAN_OPTIONS": "halt_on_error=0 exitcode=0"}, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    out = np.fromfile(state)
    return out[: 1 << n].copy(), out[1 << n:].copy(), r.stdout


@pytest.mark.parametrize("kernel", [1, 3], ids=["k_tile", "k_tile3"])
@pytest.mark.parametrize("case", ["qft", "random", "random-exact"])
def test_no_shared_memory_race_in_any_pass(emu, emu_tsan, tmp_path, case, kernel):
    if kernel == 3 and case == "random-exact":
        pytest.skip("k_tile3 runs merged mode only")
    n = 15 if case == "qft" else 13
    if case == "qft":
        qc = QuantumCircuit(QuantumRegister(n)); qc.qft()
    elif case == "random":
        qc = random_circuit(n, 160, 31)
    else:
        qc = reference_cells_circuit(n, 60, 41, exact=True)
    psi0, re, im = start(n, 77)
    _, n_pass = qc.plan()
    seen = set()
    for p in range(n_pass):
        blob = raw_pass(qc, p)
        if int(np.frombuffer(blob, dtype="<i4", count=1)[0]) != 0:
            continue
        info = (C.c_int * 4)()
        r2, i2 = re.copy(), im.copy()
        option = p % 2  # k_tile: both decode variants
        rc = emu_pass(emu, kernel, n, r2, i2, blob, qc.exact, option, info)
        if rc == 1:
            continue  # not eligible for k_tile3
        rt, it, out = tsan_run(emu_tsan, tmp_path, kernel, n, qc.exact, re, im, blob, option)
        assert np.array_equal(rt, r2) and np.array_equal(it, i2)  # same code, same arithmetic, with and without the sanitizer
        seen.add(out.strip())
        re, im = r2, i2  # feed the next pass with this pass's output, as execute would
    assert seen, "no fused pass was exercised"


def test_qcbm_circuit_benchmark(emu):
    """benches/benchmark.rs `qcbm` (ring of CX + RZ RX RZ layers, depth 3 here) through the emulated k_tile3."""
    n = 13
    qc = QuantumCircuit(QuantumRegister(n))
    assert workloads.qcbm(qc, depth=3, seed=42) == n * (5 + 4 * 3)
    psi0, re, im = start(n, 15)
    stats = {}
    run_emulated(emu, qc, re, im, stats)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    assert stats.get("ctrl", 0) >= 1, stats
