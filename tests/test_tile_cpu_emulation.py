"""The bodies of the fused tile kernels -- k_tile (csrc/kernels_tile.cu, the default) and k_tile2 (csrc/kernels_tile2.cu,
opt-in) -- executed on the CPU: tests/emu/ compiles the kernel sources themselves with g++ (one OS thread per CUDA thread, a
barrier for __syncthreads) and runs them block by block on the micro-programs that spz_execute would upload
(spz_debug_compile_pass).  Unlike tests/test_tile_program.py -- an independent NumPy statement of the execution model -- this
exercises the real indexing code: swizzled staging, register layouts, merged phase runs, exact-mode diagonal arithmetic, and
for k_tile2 the direct global<->register transfers, the lazy phase flush and the CTRL=false instantiation.

It does not replace the GPU parity tests (no warps, no memory model, no timing).  It exists because k_tile2 was written when
no GPU time was left, and it lets the CPU suite (-m "not gpu") guard the logic of both kernels.
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
    h.emu_tile2_run.restype = C.c_int
    h.emu_tile2_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_int)]
    h.emu_tile1_run.restype = C.c_int
    h.emu_tile1_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong, C.c_int, C.c_int]
    return h


def raw_pass(qc, pass_index):
    arr, n = qc._encode()
    buf = (C.c_char * (8 << 20))()
    used = C.c_int64()
    sb._check(sb._lib.spz_debug_compile_pass(qc.n_qubits, arr, n, qc._flags(), pass_index, buf, len(buf), C.byref(used)))
    return bytes(buf[: used.value])


def emu_pass(emu, kernel, n, re, im, blob, exact, option, info=None):
    """One fused pass through the emulated kernel.  kernel 1 = k_tile (option: program decoded from shared memory 1 / global
    memory 0), kernel 2 = k_tile2 (option: direct-transfer level 0..3).  Returns 1 when k_tile2 is not eligible."""
    if kernel == 1:
        return emu.emu_tile1_run(n, re.ctypes.data, im.ctypes.data, blob, len(blob), 1 if exact else 0, option)
    info = info if info is not None else (C.c_int * 4)()
    return emu.emu_tile2_run(n, re.ctypes.data, im.ctypes.data, blob, len(blob), 1 if exact else 0, option, info)


def run_emulated(emu, qc, re, im, stats, direct_level=1, kernel=2):
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
        rc = emu_pass(emu, kernel, n, re, im, blob, qc.exact, direct_level if kernel == 2 else (p & 1), info)
        if kernel == 1:
            assert rc == 0, f"pass {p}: emulation failed (rc={rc})"
            stats["k_tile"] = stats.get("k_tile", 0) + 1
            continue
        if rc == 1:  # too long for k_tile2's shared-memory budget: the launcher falls back to k_tile
            psi = run_dense_order(n, re + 1j * im, trs, by_pass[p])
            re[:], im[:] = psi.real, psi.imag
            stats["fallback"] = stats.get("fallback", 0) + 1
            continue
        assert rc == 0, f"pass {p}: emulation failed (rc={rc})"
        key = ("ctrl" if info[0] else "noctrl", "ld-direct" if info[1] else "ld-staged", "st-direct" if info[2] else "st-staged")
        stats[key] = stats.get(key, 0) + 1
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


def test_qft_merged_mode_uses_every_transfer_path(emu):
    n = 14
    qc = QuantumCircuit(QuantumRegister(n))
    qc.qft()
    psi0, re, im = start(n, 3)
    stats = {}
    run_emulated(emu, qc, re, im, stats)
    want = dense_reference(qc, psi0)
    np.testing.assert_allclose(re + 1j * im, want, rtol=0, atol=1e-12)
    # QFT has no in-tile controls on its butterflies; its first layout is {8..11} (direct load), its last {0..3} (staged store)
    assert stats.get(("noctrl", "ld-direct", "st-staged"), 0) >= 1, stats


@pytest.mark.parametrize("select", ["0", "1"], ids=["first-come-tile", "chosen-tile"])
@pytest.mark.parametrize("kernel", [1, 2], ids=["k_tile", "k_tile2"])
@pytest.mark.parametrize("n,count,seed", [(13, 160, 31), (14, 220, 32), (15, 120, 33)])
def test_random_circuits_merged_mode(emu, n, count, seed, kernel, select, monkeypatch):
    monkeypatch.setenv("SPZ_TILE_SELECT", select)  # both ways of choosing a pass's tile qubits (default: chosen from 24 qubits up)
    qc = random_circuit(n, count, seed)
    psi0, re, im = start(n, seed)
    stats = {}
    run_emulated(emu, qc, re, im, stats, kernel=kernel)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    if kernel == 2:
        assert any(k[0] == "ctrl" for k in stats if isinstance(k, tuple)), stats
    else:
        assert stats.get("k_tile", 0) >= 1


@pytest.mark.parametrize("n", [5, 8, 11])
def test_k_tile_small_registers_use_partial_tiles(emu, n):
    """n < 12: the tile is the whole register and the block has 2^(n-4) threads (k_tile only; k_tile2 needs full tiles)."""
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


@pytest.mark.parametrize("level", [0, 2, 3])
def test_every_direct_transfer_level_is_correct(emu, level):
    """SPZ_TILE_V2_DIRECT only moves the coalescing trade-off; results must not depend on it."""
    n = 13
    qc = random_circuit(n, 160, 35)
    psi0, re, im = start(n, 35)
    stats = {}
    run_emulated(emu, qc, re, im, stats, direct_level=level)
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    tiles = [k for k in stats if isinstance(k, tuple)]
    if level == 0:
        assert all(k[1:] == ("ld-staged", "st-staged") for k in tiles), stats
    if level == 3:
        assert all(k[1:] == ("ld-direct", "st-direct") for k in tiles), stats


@pytest.mark.parametrize("kernel", [1, 2], ids=["k_tile", "k_tile2"])
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


def test_vector_direct_transfers(emu):
    """Level 3 picks 256-bit accesses when register bits 0,1 are tile bits 0,1, 128-bit when only bit 0 is, one amplitude
    otherwise; build one pass of each shape at both ends and check the modes were really taken."""
    n = 13
    modes = set()
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
        plan, n_pass = qc.plan()
        for p in range(n_pass):
            blob = raw_pass(qc, p)
            if int(np.frombuffer(blob, dtype="<i4", count=1)[0]) != 0:
                trs = list(qc.transformations)
                psi = run_dense_order(n, re + 1j * im, trs, [i for i, pp in plan if pp == p])
                re[:], im[:] = psi.real, psi.imag
                continue
            info = (C.c_int * 4)()
            assert emu_pass(emu, 2, n, re, im, blob, False, 3, info) == 0
            modes.add(info[1]); modes.add(info[2])
        np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
    assert {1, 2, 4} <= modes, modes


@pytest.mark.parametrize("kernel", [1, 2], ids=["k_tile", "k_tile2"])
@pytest.mark.parametrize("n,count,seed", [(13, 60, 41), (14, 80, 42)])
def test_exact_mode_is_bit_identical_to_the_oracle(emu, n, count, seed, kernel):
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
        if rc == 1:   # program too long for k_tile2's shared-memory budget: the launcher falls back to k_tile
            s = orc.State(n)
            s.reals[:], s.imags[:] = re, im
            orc.execute(s, [ops[i] for i in by_pass[p]])
            re, im = s.reals.copy(), s.imags.copy()
            continue
        assert rc == 0
        n_tile += 1
    assert n_tile >= 1
    cpu = init.clone()
    orc.execute(cpu, ops)
    assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags)


def test_layouts_that_mix_direct_and_staged_transfers(emu):
    n = 14
    cases = []
    qc = QuantumCircuit(QuantumRegister(n))       # high layout first, low layout last
    for t in (11, 10, 9, 8):
        qc.h(t)
    qc.cp(0.3, 11, 2)
    for t in (0, 1, 2, 3):
        qc.ry(0.1 * (t + 1), t)
    cases.append((qc, ("ld-direct", "st-staged")))
    qc = QuantumCircuit(QuantumRegister(n))       # low layout first, high layout last (targets outside the low 12 bits)
    for t in (0, 1, 2, 3):
        qc.rx(0.2 * (t + 1), t)
    qc.cp(0.7, 1, 9)
    for t in (13, 12, 11, 10):
        qc.h(t)
    qc.cx(13, 12)
    cases.append((qc, ("ld-staged", "st-direct")))
    qc = QuantumCircuit(QuantumRegister(n))       # one high layout only: no shared-memory staging at all
    for t in (13, 12, 9, 8):
        qc.u(0.3, 0.2, 0.1 * t, t)
    qc.crz(0.4, 0, 13) if hasattr(qc, "crz") else qc.cp(0.4, 0, 13)
    cases.append((qc, ("ld-direct", "st-direct")))
    for i, (qc, want_paths) in enumerate(cases):
        psi0, re, im = start(n, 50 + i)
        stats = {}
        run_emulated(emu, qc, re, im, stats)
        np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)
        assert any(isinstance(k, tuple) and k[1:] == want_paths for k in stats), (i, stats)


def test_layered_circuit(emu):
    n = 13
    qc = QuantumCircuit(QuantumRegister(n))
    workloads.random_layered_circuit(qc, depth=6, seed=5)
    psi0, re, im = start(n, 9)
    run_emulated(emu, qc, re, im, {})
    np.testing.assert_allclose(re + 1j * im, dense_reference(qc, psi0), rtol=0, atol=1e-12)


# ---- shared-memory protocol under ThreadSanitizer ------------------------------------------------------------------
# The emulation's barrier is annotated per generation (tests/emu/cuda_cpu_shim.h), so a missing __syncthreads() between a
# shared-memory store and a load by another thread is a happens-before race that TSan reports deterministically.
# Mutation check (QFT-15, replace one __syncthreads() at a time by a no-op, both transfer modes): each of the six barriers
# k_tile2 has -- after seg_off; after the per-term scratch is written; before the staged tile overwrites that scratch; after
# the tables (and the staged tile) are in place; between store_regs and load_regs of a LAYOUT; before the final copy-out -- is
# reported when deleted.  The two further barriers k_tile carries per layout change ("everyone has read before anyone's next
# store_regs") were reported as unnecessary -- a thread's next store goes to the cells it has just read -- and are not in
# k_tile2.

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


@pytest.mark.parametrize("kernel", [1, 2], ids=["k_tile", "k_tile2"])
@pytest.mark.parametrize("case", ["qft", "random", "random-exact"])
def test_no_shared_memory_race_in_any_pass(emu, emu_tsan, tmp_path, case, kernel):
    n = 15 if case == "qft" else 13   # QFT-15: groups of several outer terms, so the term-parallel reduction of k_tile2 shares scratch
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
        option = p % 4 if kernel == 2 else p % 2  # cycle through the transfer / decode variants as well
        rc = emu_pass(emu, kernel, n, r2, i2, blob, qc.exact, option, info)
        if rc == 1:
            continue  # not eligible for k_tile2
        rt, it, out = tsan_run(emu_tsan, tmp_path, kernel, n, qc.exact, re, im, blob, option)
        assert np.array_equal(rt, r2) and np.array_equal(it, i2)  # same code, same arithmetic, with and without the sanitizer
        seen.add(out.strip())
        re, im = r2, i2  # feed the next pass with this pass's output, as execute would
    assert seen, "no fused pass was exercised"
