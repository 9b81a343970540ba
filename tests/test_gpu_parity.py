"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): index mapping bit-exact; amplitudes within 1e-12 absolute.  Because the
kernels mirror the reference's IEEE operation order (gate_math.cuh) and take host-computed scalars, every
unfused AND fused gate is in fact required to be BIT-IDENTICAL to the oracle here; reductions (different
summation order) are held to 1e-12.
"""
import math

import numpy as np
import pytest

import oracle as orc
import spinoza_b200 as sb
from spinoza_b200 import Gate, QuantumCircuit, QuantumRegister
from tests import _dense as D

pytestmark = pytest.mark.gpu
PI = math.pi

GATES = [(orc.H, ()), (orc.X, ()), (orc.Y, ()), (orc.Z, ()), (orc.P, (0.37,)), (orc.RX, (1.0,)),
         (orc.RY, (-0.6,)), (orc.RZ, (1.0,)), (orc.U, (0.3, 1.4, -0.8))]


def G(kind, p=()):
    return Gate(kind, p)


def to_gpu(s: orc.State) -> sb.State:
    return sb.State.from_arrays(s.reals, s.imags)


def assert_same(gpu: sb.State, cpu: orc.State, exact=True, tol=1e-12, what=""):
    re, im = gpu.download()
    if exact:
        assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags), \
            f"{what}: max abs diff {max(np.max(np.abs(re - cpu.reals)), np.max(np.abs(im - cpu.imags)))}"
    else:
        assert np.max(np.abs(re - cpu.reals)) <= tol and np.max(np.abs(im - cpu.imags)) <= tol, what


def test_loaded_library_is_the_in_tree_cuda_build():
    assert sb.device_count() >= 1
    assert sb.library_path().endswith("spinoza_b200/lib/libspinoza_b200.so")
    assert "sm_100" in sb.device_name(0) or "sm_10" in sb.device_name(0), sb.device_name(0)


def test_state_new_is_zero_ket():  # core.rs:32-42
    for n in (1, 2, 3, 10):
        s = sb.State(n)
        re, im = s.download()
        assert re[0] == 1.0 and np.count_nonzero(re) == 1 and np.count_nonzero(im) == 0 and len(s) == 1 << n


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 13])
def test_apply_every_gate_every_target_bit_exact(n):
    launches0 = sb.launch_count()
    for kind, p in GATES:
        for t in range(n):
            cpu = orc.gen_random_state(n, 1000 + 17 * n + t)
            gpu = to_gpu(cpu)
            orc.apply(kind, cpu, t, p)
            sb.apply(G(kind, p), gpu, t)
            assert_same(gpu, cpu, what=f"apply kind={kind} n={n} t={t}")
    assert sb.launch_count() > launches0  # our kernels ran, not a fallback


@pytest.mark.parametrize("n", [2, 3, 5, 8, 12])
def test_c_apply_every_gate_every_pair_bit_exact(n):
    for kind, p in GATES:
        if kind == orc.Z:
            continue  # not in the reference's c_apply (gates.rs:258-267); covered as an extension below
        pairs = [(c, t) for c in range(n) for t in range(n) if c != t]
        if n >= 8:
            pairs = [pr for i, pr in enumerate(pairs) if i % 5 == 0 or 0 in pr or 1 in pr or n - 1 in pr]
        for c, t in pairs:
            cpu = orc.gen_random_state(n, 2000 + 31 * c + t)
            gpu = to_gpu(cpu)
            orc.c_apply(kind, cpu, c, t, p)
            sb.c_apply(G(kind, p), gpu, c, t)
            assert_same(gpu, cpu, what=f"c_apply kind={kind} n={n} c={c} t={t}")


@pytest.mark.parametrize("n", [3, 4, 6, 10])
def test_mc_and_cc_apply_all_masks(n):
    rng = np.random.default_rng(n)
    masks = range(1, 1 << n) if n <= 6 else [int(x) for x in rng.integers(1, 1 << n, 60)]
    for kind, p in [(orc.X, ()), (orc.P, (3.14,)), (orc.RX, (0.9,)), (orc.RY, (0.7,))]:
        for mask in masks:
            for t in range(n):
                if (mask >> t) & 1 or bin(mask).count("1") >= n:
                    continue
                cpu = orc.gen_random_state(n, 3000 + mask)
                gpu = to_gpu(cpu)
                orc.mc_apply_mask(kind, cpu, mask, t, p)
                controls = [q for q in range(n) if (mask >> q) & 1]
                sb.mc_apply(G(kind, p), gpu, controls, None, t)
                assert_same(gpu, cpu, what=f"mc_apply kind={kind} n={n} mask={mask:b} t={t}")
    cpu = orc.gen_random_state(n, 5)
    gpu = to_gpu(cpu)
    orc.cc_apply(orc.X, cpu, 0, n - 1, 1)
    sb.cc_apply(Gate.X, gpu, 0, n - 1, 1)
    assert_same(gpu, cpu, what="cc_apply")


def test_mc_apply_zeros_dropped_like_reference():  # gates.rs:298-311 (B4 mirrored)
    cpu = orc.gen_random_state(5, 8)
    gpu = to_gpu(cpu)
    orc.mc_apply(orc.X, cpu, [0, 1, 2], {1}, 4)
    sb.mc_apply(Gate.X, gpu, [0, 1, 2], {1}, 4)
    assert_same(gpu, cpu)


def test_reference_doc_and_example_cases():
    # doc-test gates.rs:282-289; examples/ccx.rs:23 cc(0,2->1); examples/multicontrol.rs:12 mc X [1,2]->0
    cpu, gpu = orc.State(3), sb.State(3)
    orc.mc_apply(orc.P, cpu, [0, 1], None, 2, (3.14,))
    sb.mc_apply(Gate.P(3.14), gpu, [0, 1], None, 2)
    assert_same(gpu, cpu)
    cpu = orc.gen_random_state(3, 1); gpu = to_gpu(cpu)
    orc.cc_apply(orc.X, cpu, 0, 2, 1); sb.cc_apply(Gate.X, gpu, 0, 2, 1)
    assert_same(gpu, cpu)
    cpu = orc.gen_random_state(3, 2); gpu = to_gpu(cpu)
    orc.mc_apply(orc.X, cpu, [1, 2], None, 0); sb.mc_apply(Gate.X, gpu, [1, 2], None, 0)
    assert_same(gpu, cpu)


@pytest.mark.parametrize("n", [2, 3, 9, 12])
def test_swap_matches_oracle_and_three_cx(n):  # gates.rs:1873-1901 (bit-exact)
    for t0 in range(n):
        for t1 in range(n):
            if n > 9 and (t0 + t1) % 3:
                continue
            cpu = orc.gen_random_state(n, 70 + t0 * n + t1)
            gpu = to_gpu(cpu)
            gpu2 = to_gpu(cpu)
            orc.swap(cpu, t0, t1)
            sb.apply(Gate.SWAP(t0, t1), gpu, 0)
            assert_same(gpu, cpu, what=f"swap {t0},{t1}")
            if t0 != t1:
                sb.c_apply(Gate.X, gpu2, t0, t1); sb.c_apply(Gate.X, gpu2, t1, t0); sb.c_apply(Gate.X, gpu2, t0, t1)
                assert_same(gpu2, cpu, what=f"3cx {t0},{t1}")


def test_golden_vectors_through_the_gpu():
    # the reference's own 1e-10 known answers (gates.rs:1541, 1659-1681, 1727-1728, 1820-1842, 1854-1869)
    s = sb.State(3)
    for t in range(3):
        sb.apply(Gate.H, s, t)
    assert np.max(np.abs(s.reals - 0.35355339059327384)) < 1e-10
    s = sb.State(3)
    for t in range(3):
        sb.apply(Gate.RZ(1.0), s, t)
    assert abs(s.amp(0) - complex(0.07073720166770296, -0.9974949866040546)) < 1e-10
    s = sb.State(3)
    for t in range(3):
        sb.apply(Gate.U(1.0, 1.0, 1.0), s, t)
    assert abs(s.amp(7) - complex(-0.1090926263889472, 0.015550776766638148)) < 1e-10
    s = sb.State(1)
    sb.apply(Gate.U(2.0, 3.0, 1.0), s, 0)
    assert abs(s.amp(1) - complex(-0.833049961066805, 0.11874839215823475)) < 1e-10


def qcbm(n, state, apply, c_apply):  # gates.rs:1499-1529
    pairs = [(i, (i + 1) % n) for i in range(n)]
    for i in range(n):
        apply("rx", state, i); apply("rz", state, i)
    for p0, p1 in pairs[: n - 1]:
        c_apply(state, p0, p1)
    for _ in range(9):
        for i in range(n):
            apply("rz", state, i); apply("rx", state, i); apply("rz", state, i)
        for p0, p1 in pairs[: n - 1]:
            c_apply(state, p0, p1)
    for i in range(n):
        apply("rz", state, i); apply("rx", state, i)


def test_qcbm_20_qubits_golden():  # gates.rs:1798-1809
    n = 20
    s = sb.State(n)
    qcbm(n, s, lambda g, st, t: sb.apply(Gate.RX(1.0) if g == "rx" else Gate.RZ(1.0), st, t),
         lambda st, c, t: sb.c_apply(Gate.X, st, c, t))
    assert abs(s.amp(0) - complex(-0.0022221321676945643, 0.001743068112560825)) < 1e-10
    assert abs(s.amp(7) - complex(-0.0031017461877124453, -0.0034043237120339686)) < 1e-10
    assert abs(s.amp(12) - complex(0.0005494086136357235, -0.00009827749580581964)) < 1e-10
    c = orc.State(n)
    qcbm(n, c, lambda g, st, t: orc.apply(orc.RX if g == "rx" else orc.RZ, st, t, (1.0,)),
         lambda st, cc, t: orc.c_apply(orc.X, st, cc, t))
    assert_same(s, c, what="qcbm-20 vs oracle")


# ---- execute: fused vs unfused vs oracle ---------------------------------------------------------------
def random_ops(n, count, seed, with_swap=True):
    rng = np.random.default_rng(seed)
    ops = []
    kinds = [k for k, _ in GATES]
    for _ in range(count):
        r = rng.random()
        t = int(rng.integers(n))
        kind = kinds[int(rng.integers(len(kinds)))]
        p = tuple(rng.random(3) * 2 * PI)
        if r < 0.45 or n == 1:
            ops.append(("g", kind, p, t, 0))
        elif r < 0.8:
            c = int(rng.integers(n - 1)); c += c >= t
            ops.append(("g", kind, p, t, 1 << c))
        elif r < 0.92 and n >= 3:
            k = int(rng.integers(2, min(n, 4)))
            cs = rng.choice([q for q in range(n) if q != t], size=k, replace=False)
            ops.append(("g", kind, p, t, int(sum(1 << int(c) for c in cs))))
        elif with_swap and n >= 2:
            a = int(rng.integers(n)); b = int(rng.integers(n))
            ops.append(("s", a, b))
        else:
            ops.append(("g", kind, p, t, 0))
    return ops


def run_dense(n, psi, ops):
    for o in ops:
        if o[0] == "s":
            psi = D.apply_swap(psi, n, o[1], o[2])
        else:
            _, kind, p, t, cm = o
            psi = D.apply_matrix(psi, n, D.matrix(kind, p), t, cm)
    return psi


def build_circuit(n, ops, state, fuse, exact=False):
    qc = QuantumCircuit.from_state(state, fuse=fuse, exact=exact)
    for o in ops:
        if o[0] == "s":
            qc.swap(o[1], o[2])
        else:
            _, kind, p, t, cm = o
            cs = [q for q in range(n) if (cm >> q) & 1]
            ctrl = sb.Controls.none() if not cs else sb.Controls.single(cs[0]) if len(cs) == 1 else sb.Controls.mixed(cs, set())
            qc.add(sb.QuantumTransformation(G(kind, p), t, ctrl))
    return qc


@pytest.mark.parametrize("n,count,seed", [(1, 20, 1), (2, 40, 2), (3, 60, 3), (4, 60, 10), (5, 80, 11), (6, 120, 4),
                                          (11, 150, 5), (12, 200, 6), (13, 200, 7), (16, 300, 8), (20, 160, 9)])
def test_fused_execute_vs_unfused(n, count, seed):
    """Fused + EXACT must be bit-identical to unfused; default fused (merged diagonal runs) within 1e-12."""
    ops = random_ops(n, count, seed)
    init = orc.gen_random_state(n, 500 + seed)
    a, b, c = to_gpu(init), to_gpu(init), to_gpu(init)
    build_circuit(n, ops, a, fuse=True, exact=True).execute()
    build_circuit(n, ops, b, fuse=False).execute()
    build_circuit(n, ops, c, fuse=True, exact=False).execute()
    ra, ia = a.download(); rb, ib = b.download(); rc, ic = c.download()
    assert np.array_equal(ra, rb) and np.array_equal(ia, ib)
    assert np.max(np.abs(rc - rb)) <= 1e-12 and np.max(np.abs(ic - ib)) <= 1e-12
    if n <= 16:
        want = run_dense(n, init.amps(), ops)
        assert np.max(np.abs((ra + 1j * ia) - want)) < 1e-12
        assert np.max(np.abs((rc + 1j * ic) - want)) < 1e-12


@pytest.mark.parametrize("n", [8, 14, 18])
def test_fused_diagonal_heavy_circuit(n):
    """Long runs of P / CP / RZ / Z / multi-controlled P between Hadamards: the merged-phase path."""
    rng = np.random.default_rng(n)
    ops = []
    for layer in range(6):
        for t in range(n):
            ops.append(("g", orc.H, (), t, 0))
        for _ in range(12 * n):
            kind = [orc.P, orc.RZ, orc.Z][int(rng.integers(3))]
            t = int(rng.integers(n))
            k = int(rng.integers(0, 4))
            cs = rng.choice([q for q in range(n) if q != t], size=min(k, n - 1), replace=False)
            ops.append(("g", kind, (float(rng.random() * 6.28),), t, int(sum(1 << int(c) for c in cs))))
    init = orc.gen_random_state(n, 800 + n)
    a, b, c = to_gpu(init), to_gpu(init), to_gpu(init)
    build_circuit(n, ops, a, fuse=True, exact=True).execute()
    build_circuit(n, ops, b, fuse=False).execute()
    build_circuit(n, ops, c, fuse=True).execute()
    ra, ia = a.download(); rb, ib = b.download(); rc, ic = c.download()
    assert np.array_equal(ra, rb) and np.array_equal(ia, ib)
    assert np.max(np.abs(rc - rb)) <= 1e-12 and np.max(np.abs(ic - ib)) <= 1e-12
    assert abs(sb.norm2(c) - 1.0) < 1e-10


def oracle_ops_from(qc):
    out = []
    for t in qc.transformations:
        out.append(orc.make_op(t.gate.kind, t.target, t.gate.params, ctrl_kind=t.controls.kind,
                               ctrl_mask=t.controls.mask(), zeros_mask=t.controls.zeros_mask(), t0=t.gate.t0, t1=t.gate.t1))
    return out


@pytest.mark.parametrize("fuse,exact", [(False, False), (True, True), (True, False)])
def test_all_gates_as_transformations_n17(fuse, exact):  # circuit.rs:771-822
    n = 17
    qc = QuantumCircuit(QuantumRegister(n), fuse=fuse, exact=exact)
    for t in range(n):
        qc.h(t)
    qc.x(0); qc.y(1); qc.z(2); qc.p(PI, 3); qc.cp(PI, 3, 4); qc.rx(PI, 5); qc.ry(PI, 6); qc.rz(PI, 7)
    qc.u(PI, PI, PI, 8); qc.cy(9, 10); qc.crx(PI, 11, 12); qc.cry(PI, 13, 14)
    cpu = orc.State(n)
    orc.execute(cpu, oracle_ops_from(qc))
    qc.execute()
    assert qc.transformations == []
    assert_same(qc.state, cpu, exact=(exact or not fuse), tol=1e-12, what="circuit.rs all_gates_as_transformations")


@pytest.mark.parametrize("n,fuse", [(5, False), (5, True), (14, True), (20, True), (24, True)])
def test_qft_closed_form_and_roundtrip(n, fuse):  # QFT|x>[k] = 2^(-n/2) exp(+2 pi i x rev(k) / 2^n)
    x = 0x9E3779B97F4A7C15 % (1 << n)
    s = sb.State(n)
    s.set_basis(x)
    qc = QuantumCircuit.from_state(s, fuse=fuse)
    qc.qft()
    ops = oracle_ops_from(qc)
    qc.execute()
    k = np.arange(1 << n)
    rev = np.zeros_like(k)
    for b in range(n):
        rev |= ((k >> b) & 1) << (n - 1 - b)
    want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ((x * rev) % (1 << n)) / (1 << n))
    assert np.max(np.abs(s.amps() - want)) < 1e-12
    if n <= 20:
        cpu = orc.State(n); cpu.reals[0] = 0.0; cpu.reals[x] = 1.0
        orc.execute(cpu, ops)
        assert_same(s, cpu, exact=not fuse, tol=1e-12, what="QFT vs oracle")
    # QFT then IQFT returns the start state (circuit.rs:929-960)
    qc.iqft(list(reversed(range(n))))
    qc.execute()
    re, im = s.download()
    assert abs(re[x] - 1.0) < 1e-12 and np.max(np.abs(np.delete(re, x))) < 1e-12 and np.max(np.abs(im)) < 1e-12
    assert abs(sb.norm2(s) - 1.0) < 1e-10


@pytest.mark.parametrize("n,mode", [(9, None), (17, None), (20, None), (17, "exact"), (20, "0")])
def test_iqft_functional_matches_oracle(n, mode, monkeypatch):  # core.rs:184-191
    """From 16 qubits up spz_iqft hands its gate list to the fused scheduler: merged mode by default (within 1e-12, a handful of
    passes), SPZ_IQFT_FUSE=exact bit-identical in as few passes, SPZ_IQFT_FUSE=0 the gate-by-gate loop."""
    if mode is not None:
        monkeypatch.setenv("SPZ_IQFT_FUSE", mode)
    cpu = orc.gen_random_state(n, 77)
    gpu = to_gpu(cpu)
    targets = list(reversed(range(n)))
    orc.iqft(cpu, targets)
    before = sb.launch_count()
    sb.iqft(gpu, targets)
    launches = sb.launch_count() - before
    loop = n < 16 or mode == "0"
    assert_same(gpu, cpu, exact=loop or mode == "exact", tol=1e-12)
    assert launches == n + n * (n - 1) // 2 if loop else launches <= 8, launches
    if n >= 16:  # a subset of the qubits, in an order of the caller's choosing
        sub = [3, n - 1, 0, 8, 12]
        orc.iqft(cpu, sub)
        sb.iqft(gpu, sub)
        assert_same(gpu, cpu, exact=False, tol=1e-12)
        with pytest.raises(sb.SpinozaError):
            sb.iqft(gpu, [0, 1, 1])
        with pytest.raises(sb.SpinozaError):
            sb.iqft(gpu, [0, 1, n])


def test_value_encoding_yields_basis_state():  # circuit.rs:1076-1113
    n, v = 3, 4.0
    for fuse in (False, True):
        qc = QuantumCircuit(QuantumRegister(n), fuse=fuse)
        for t in range(n):
            qc.h(t)
        for t in range(n):
            qc.p(2.0 * PI / (2.0 ** (t + 1)) * v, t)
        qc.iqft(list(reversed(range(n))))
        qc.execute()
        a = qc.state.amps()
        want = np.zeros(8); want[4] = 1.0
        assert np.max(np.abs(a - want)) < 1e-4


# ---- reductions / measurement -----------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 7, 11, 14, 19, 22])
def test_prob0_norm_expectations(n):
    cpu = orc.gen_random_state(n, 900 + n)
    gpu = to_gpu(cpu)
    assert abs(sb.norm2(gpu) - orc.norm2(cpu)) < 1e-12
    targets = list(range(n)) if n <= 7 else [0, 1, n // 2, n - 1]
    for t in targets:
        assert abs(sb.prob0(gpu, t) - orc.prob0(cpu, t)) < 1e-12
        assert abs(sb.qubit_expectation_value(gpu, t) - orc.qubit_expectation_value(cpu, t)) < 1e-12
    for obs in "xyz":
        got = sb.xyz_expectation_value(obs, gpu, targets)
        want = orc.xyz_expectation_value(obs, cpu, targets)
        assert np.max(np.abs(np.array(got) - want)) < 1e-12, obs
    if n >= 2:  # 'z' on several targets is ONE read pass for all of them (csrc/kernels_zall.cuh): pass + final sum
        before = sb.launch_count()
        got = sb.xyz_expectation_value("z", gpu, list(range(n)))
        assert sb.launch_count() - before == 2
        assert np.max(np.abs(np.array(got) - orc.xyz_expectation_value("z", cpu, list(range(n))))) < 1e-12
        with pytest.raises(sb.SpinozaError):
            sb.xyz_expectation_value("z", gpu, [0, n])
    if 7 <= n <= 19:  # 'x' / 'y' on three or more targets: tiles staged in shared memory, up to twelve targets per read pass
        every = list(range(n))
        for obs in "xy":
            before = sb.launch_count()
            got = sb.xyz_expectation_value(obs, gpu, every)
            assert sb.launch_count() - before == 2 * (1 + (max(0, n - 12) + 5) // 6)   # (pass + final sum) per group of targets
            assert np.max(np.abs(np.array(got) - orc.xyz_expectation_value(obs, cpu, every))) < 1e-12, obs
        mixed = [n - 1, 0, n - 1, 3, n - 2]           # duplicates and any order
        assert np.max(np.abs(np.array(sb.xyz_expectation_value("y", gpu, mixed)) - orc.xyz_expectation_value("y", cpu, mixed))) < 1e-12
        with pytest.raises(sb.SpinozaError):
            sb.xyz_expectation_value("x", gpu, [0, 1, n])
    re, im = gpu.download()  # reductions must not modify the state (the reference clones, core.rs:227)
    assert np.array_equal(re, cpu.reals) and np.array_equal(im, cpu.imags)


def test_xyz_expectation_reference_value_and_bad_observable():  # core.rs:294-310
    s = sb.State(1)
    sb.apply(Gate.RX(0.54), s, 0)
    sb.apply(Gate.RY(0.12), s, 0)
    assert abs(sb.xyz_expectation_value("z", s, [0])[0] - 0.8515405859048367) < 1e-4
    assert abs(sb.qubit_expectation_value(s, 0) - 0.8515405859048367) < 1e-4
    with pytest.raises(sb.SpinozaError):
        sb.xyz_expectation_value("a", s, [0])


@pytest.mark.parametrize("n", [1, 3, 10, 16])
def test_measure_qubit_forced_outcomes(n):  # measurement.rs:12-92
    for t in sorted({0, n // 2, n - 1}):
        for v in (0, 1):
            for reset in (False, True):
                cpu = orc.gen_random_state(n, 40 + t)
                gpu = to_gpu(cpu)
                bit_c, _ = orc.measure_qubit(cpu, t, reset, v)
                bit_g = sb.measure_qubit(gpu, t, reset, v)
                assert bit_c == bit_g == v
                assert_same(gpu, cpu, exact=False, tol=1e-12, what=f"measure n={n} t={t} v={v} reset={reset}")
                assert abs(sb.norm2(gpu) - 1.0) < 1e-10


def test_measure_qubit_known_state():  # measurement.rs:145-246
    vals = [0.034172256444052966, 0.29007027387615136, -0.1300556493088507, 0.47222164829858637,
            -0.032338373524095645, 0.26511510737291843, 0.1259630181898572, -0.09645897805840803,
            -0.31931099330088214, -0.24644972468157703, -0.15963222942036193, -0.14329373536970438,
            -0.1564141838467382, -0.4751067410290973, 0.1034273381193853, -0.32966556091031934]
    s = sb.State.from_arrays(vals[0::2], vals[1::2])
    sb.measure_qubit(s, 0, True, 0)
    assert abs(s.amp(0) - complex(0.04528096797370981, 0.38436627101331156)) < 1e-3
    assert abs(s.amp(6) - complex(-0.2072612811212442, -0.6295543626114914)) < 1e-3
    sb.measure_qubit(s, 1, True, 0)
    assert abs(s.amp(4) - complex(-0.6411848150109799, -0.49487748447346463)) < 1e-3
    sb.measure_qubit(s, 2, True, 1)
    assert abs(s.amp(0) - complex(-0.7916334352111761, -0.6109963209838112)) < 1e-3
    assert np.max(np.abs(s.amps()[1:])) < 1e-3


def test_measure_all_twice_gives_identical_bits():  # circuit.rs:825-889 (N = 21 in the reference)
    n = 21
    s = sb.State(n)
    s.init_random(42)
    s.set_seed(7)
    assert abs(sb.norm2(s) - 1.0) < 1e-3
    qc = QuantumCircuit.from_state(s)
    for t in range(n):
        qc.measure(t)
    qc.execute()
    vals = [qc.get_qubit_measured_val(t) for t in range(n)]
    assert all(v in (0, 1) for v in vals)
    for t in range(n):
        qc.measure(t)
    qc.execute()
    assert [qc.get_qubit_measured_val(t) for t in range(n)] == vals
    assert abs(s.amp(0)) == pytest.approx(1.0, abs=1e-9)  # reset=true collapses to |0..0>


@pytest.mark.parametrize("n,order", [(5, "up"), (12, "down"), (17, "mixed"), (21, "up")])
def test_run_of_measurements_visits_only_the_live_subspace(n, order):
    """A run of M ops inside one execute (measure everything, circuit.rs:825-889): each measurement resets its qubit to |0>, so
    the later ones read only the indices where the earlier qubits are 0.  Same seed, same outcomes and -- the partial sums differ
    only in order -- the same state within 1e-12 as measuring with one execute per qubit (no run, full passes)."""
    qs = {"up": list(range(n)), "down": list(reversed(range(n))), "mixed": [(7 * i + 3) % n for i in range(n)]}[order]
    init = orc.gen_random_state(n, 600 + n)
    a, b = to_gpu(init), to_gpu(init)
    a.set_seed(99); b.set_seed(99)
    qa = QuantumCircuit.from_state(a)
    qa.h(0)                                  # something before the run ...
    for t in qs[: n - 1]:
        qa.measure(t)
    qa.measure(qs[0])                        # ... a qubit measured twice inside it ...
    qa.measure(qs[n - 1])
    qa.execute()
    qb = QuantumCircuit.from_state(b)
    qb.h(0)
    qb.execute()
    for t in qs:
        qb.measure(t)
        qb.execute()                         # one execute per measurement: every pass is a full one
    va = [qa.get_qubit_measured_val(t) for t in range(n)]
    vb = [qb.get_qubit_measured_val(t) for t in range(n)]
    assert va == vb and all(v in (0, 1) for v in va)
    ra, ia = a.download(); rb, ib = b.download()
    assert np.max(np.abs(ra - rb)) <= 1e-12 and np.max(np.abs(ia - ib)) <= 1e-12
    assert abs(abs(a.amp(0)) - 1.0) < 1e-9 and np.count_nonzero(ra) + np.count_nonzero(ia) <= 2   # |0..0> up to a phase
    # a gate between two measurements ends the run: the next measurement must see the whole state again
    c = to_gpu(init); c.set_seed(5)
    d = to_gpu(init); d.set_seed(5)
    qc1 = QuantumCircuit.from_state(c)
    qc1.measure(1); qc1.h(1); qc1.measure(0); qc1.x(1); qc1.measure(2 % n)
    qc1.execute()
    for build in (lambda q: q.measure(1), lambda q: q.h(1), lambda q: q.measure(0), lambda q: q.x(1), lambda q: q.measure(2 % n)):
        q = QuantumCircuit.from_state(d)
        build(q)
        q.execute()
    rc, ic = c.download(); rd, idd = d.download()
    assert np.max(np.abs(rc - rd)) <= 1e-12 and np.max(np.abs(ic - idd)) <= 1e-12
    assert abs(sb.norm2(c) - 1.0) < 1e-10


def test_measurement_statistics_follow_born_rule():
    n = 3
    cpu = orc.gen_random_state(n, 3)
    p1 = 1.0 - orc.prob0(cpu, 1)
    ones = 0
    trials = 2000
    gpu = to_gpu(cpu)
    gpu.set_seed(123)
    for _ in range(trials):
        g = gpu.clone()
        g.set_seed(int(np.random.default_rng(_).integers(1 << 62)))
        ones += sb.measure_qubit(g, 1, False, None)
    sigma = math.sqrt(p1 * (1 - p1) / trials)
    assert abs(ones / trials - p1) < 5 * sigma


def test_classical_control_and_bitflip_in_execute():  # circuit.rs:570-574, gates.rs:1365-1374
    qc = QuantumCircuit(QuantumRegister(2))
    qc.x(0); qc.measure(0); qc.cx(0, 1)
    qc.execute()
    assert qc.get_qubit_measured_val(0) == 1
    assert abs(qc.state.amp(2) - 1.0) < 1e-12  # qubit 0 reset to 0, classically controlled X fired on qubit 1
    init = orc.gen_random_state(1, 5)
    s = to_gpu(init)
    qc = QuantumCircuit.from_state(s)
    qc.bit_flip_noise(0.0, 0)
    qc.execute()
    assert_same(s, init)
    qc.bit_flip_noise(1.0, 0)
    qc.execute()
    re, im = s.download()
    assert re[0] == init.reals[1] and re[1] == init.reals[0] and im[0] == init.imags[1]


def test_unsupported_combinations_return_errors_not_aborts():  # gates.rs:230,267,275,318; circuit.rs:597
    s = sb.State(3)
    with pytest.raises(sb.SpinozaError) as e:
        sb.apply(Gate.M, s, 0)
    assert e.value.status == sb.ERR_UNSUPPORTED
    for g in (Gate.SWAP(0, 1), Gate.M, Gate.BitFlipNoise(0.5)):
        with pytest.raises(sb.SpinozaError):
            sb.c_apply(g, s, 0, 1)
        with pytest.raises(sb.SpinozaError):
            sb.mc_apply(g, s, [0, 1], None, 2)
    with pytest.raises(sb.SpinozaError):
        sb.apply(Gate.H, s, 3)
    with pytest.raises(sb.SpinozaError):
        sb.c_apply(Gate.H, s, 1, 1)
    with pytest.raises(sb.SpinozaError):
        sb.apply(Gate.SWAP(0, 5), s, 0)  # assert! gates.rs:1377
    with pytest.raises(sb.SpinozaError):
        sb.State(0)  # assert!(n > 0) core.rs:33
    re, im = s.download()
    assert re[0] == 1.0  # failed calls left the state untouched


def test_extension_cells_match_dense():  # controlled Z / mc H etc.: the reference panics, we compute the obvious thing
    n = 5
    init = orc.gen_random_state(n, 61)
    s = to_gpu(init)
    sb.c_apply(Gate.Z, s, 4, 0)
    sb.mc_apply(Gate.H, s, [0, 3], None, 2)
    sb.cc_apply(Gate.U(0.1, 0.2, 0.3), s, 1, 2, 4)
    psi = init.amps()
    psi = D.apply_matrix(psi, n, D.matrix(D.Z), 0, 1 << 4)
    psi = D.apply_matrix(psi, n, D.matrix(D.H), 2, 0b01001)
    psi = D.apply_matrix(psi, n, D.matrix(D.U, (0.1, 0.2, 0.3)), 4, 0b00110)
    assert np.max(np.abs(s.amps() - psi)) < 1e-14


@pytest.mark.parametrize("n", [3, 6, 13])
def test_signed_controls_one_launch_equals_x_conjugation(n):
    """spz_mc_apply_signed (extension; the reference's Mixed { zeros } drops the zeros, gates.rs:298-311): true negative controls
    in ONE launch of the pair kernel.  X is an exchange, so X(zeros) . oracle gate . X(zeros) is bit-identical where the oracle
    has the cell, and the dense statement covers the rest."""
    rng = np.random.default_rng(300 + n)
    for kind, p in GATES:
        for _ in range(8):
            t = int(rng.integers(n))
            others = [q for q in range(n) if q != t]
            cs = [int(c) for c in rng.choice(others, size=int(rng.integers(1, min(3, len(others)) + 1)), replace=False)]
            zs = [c for c in cs if rng.random() < 0.6] or [cs[0]]
            ones = [c for c in cs if c not in zs]
            init = orc.gen_random_state(n, int(rng.integers(1 << 20)))
            s = to_gpu(init)
            before = sb.launch_count()
            sb.mc_apply_signed(G(kind, p), s, ones, zs, t)
            assert sb.launch_count() - before == 1
            cm, zm = sum(1 << c for c in cs), sum(1 << c for c in zs)
            want = D.apply_matrix(init.amps(), n, D.matrix(kind, p), t, cm, zm)
            assert np.max(np.abs(s.amps() - want)) < 1e-14, (kind, cs, zs, t)
            if kind == orc.Z or (len(cs) > 1 and kind not in (orc.X, orc.P, orc.RX, orc.RY)):
                continue
            cpu = init.clone()
            for z in zs:
                orc.apply(orc.X, cpu, z)
            if len(cs) == 1:
                orc.c_apply(kind, cpu, cs[0], t, p)
            else:
                orc.mc_apply(kind, cpu, cs, None, t, p)
            for z in zs:
                orc.apply(orc.X, cpu, z)
            assert_same(s, cpu, what=f"signed {kind} controls {cs} zeros {zs} target {t}")
    s = sb.State(4)
    with pytest.raises(sb.SpinozaError):
        sb.mc_apply_signed(Gate.X, s, [1], [1], 0)     # both signs
    with pytest.raises(sb.SpinozaError):
        sb.mc_apply_signed(Gate.X, s, [1], [0], 0)     # target among the controls
    with pytest.raises(sb.SpinozaError):
        sb.mc_apply_signed(Gate.SWAP(0, 1), s, [2], [3], 0)


@pytest.mark.parametrize("fuse,exact", [(False, False), (True, False), (True, True)])
def test_signed_controls_in_execute(fuse, exact):
    """Controls.signed through QuantumCircuit::execute: one launch per op unfused, X . gate . X inside the fused passes."""
    n = 14
    rng = np.random.default_rng(41)
    qc = QuantumCircuit(QuantumRegister(n), fuse=fuse, exact=exact)
    init = orc.gen_random_state(n, 77)
    qc.state = to_gpu(init)
    psi = init.amps()
    kinds = [D.X, D.P, D.RX, D.RY, D.H, D.RZ, D.U, D.Y, D.Z]
    for i in range(40):
        t = int(rng.integers(n))
        others = [q for q in range(n) if q != t]
        cs = [int(c) for c in rng.choice(others, size=int(rng.integers(1, 4)), replace=False)]
        zs = {c for c in cs if rng.random() < 0.5} or {cs[0]}
        kind, p = kinds[i % len(kinds)], tuple(float(x) for x in rng.random(3) * 2 * PI)
        qc.add(sb.QuantumTransformation(Gate(kind, p), t, sb.Controls.signed(cs, zs)))
        psi = D.apply_matrix(psi, n, D.matrix(kind, p), t, sum(1 << c for c in cs), sum(1 << c for c in zs))
        h = int(rng.integers(n))
        qc.h(h)
        psi = D.apply_matrix(psi, n, D.matrix(D.H), h)
    before = sb.launch_count()
    qc.execute()
    launches = sb.launch_count() - before
    assert (launches == 80) if not fuse else (launches < 40), launches
    assert np.max(np.abs(qc.state.amps() - psi)) < 1e-12


def test_async_upload_then_gates_equals_upload_then_gates():
    """spz_upload_async: gates and fused passes issued right after it follow the pieces of the state as they arrive (low targets
    chunk by chunk, high targets and controls after the last piece).  Bit-identical to the synchronous upload."""
    n = 20
    cpu = orc.gen_random_state(n, 5)
    hre, him = sb.HostBuffer(1 << n), sb.HostBuffer(1 << n)
    hre.array[:] = cpu.reals
    him.array[:] = cpu.imags

    def gates(s):
        sb.apply(Gate.H, s, 3)
        sb.apply(Gate.RX(0.7), s, n - 3)       # below the piece bits: follows the pieces
        sb.apply(Gate.RY(0.2), s, n - 1)       # a piece bit: after the last piece
        sb.c_apply(Gate.P(0.4), s, n - 2, 5)   # control on a piece bit
        sb.apply(Gate.RZ(1.3), s, 0)

    a = sb.State(n)
    a.upload_from(hre, him)
    gates(a)
    for rep in range(3):                       # repeated: the copy stream must wait for the gates of the previous round
        b = sb.State(n) if rep == 0 else b
        b.upload_async(hre, him)
        gates(b)
        assert np.array_equal(a.download()[0], b.download()[0]) and np.array_equal(a.download()[1], b.download()[1])
    # a fused pass right behind the upload
    c = sb.State(n)
    c.upload_from(hre, him)
    qa = QuantumCircuit.from_state(c, fuse=True); qa.qft(); qa.execute()
    d = sb.State(n)
    d.upload_async(hre, him)
    qb = QuantumCircuit.from_state(d, fuse=True); qb.qft(); qb.execute()
    assert np.array_equal(c.download()[0], d.download()[0]) and np.array_equal(c.download()[1], d.download()[1])
    # reductions and a second upload join the copy stream
    d.upload_async(hre, him)
    assert abs(sb.norm2(d) - 1.0) < 1e-12
    d.upload_async(hre, him)
    d.upload_async(hre, him)
    assert np.array_equal(d.download()[0], cpu.reals)


def test_streamed_round_trip_equals_the_synchronous_one():
    """upload_async -> a sweep of gates -> download_into: runs of gates that act inside the pieces go to one stream per piece
    (also after a join, from 24 qubits), diagonal gates on a piece bit become constant factors per piece, controls on piece bits
    select pieces, and the download leaves piece by piece from the lanes.  Bit-identical to upload -> gates -> download."""
    n = 24
    cpu = orc.gen_random_state(n, 11)
    hre, him = sb.HostBuffer(1 << n), sb.HostBuffer(1 << n)
    ore, oim = sb.HostBuffer(1 << n), sb.HostBuffer(1 << n)

    def sweep(s):
        for g in (Gate.H, Gate.RX(1.0), Gate.RZ(1.0)):
            for t in range(n):
                sb.apply(g, s, t)
        sb.apply(Gate.P(0.3), s, n - 1)
        sb.apply(Gate.Z, s, n - 2)
        sb.c_apply(Gate.RY(0.4), s, n - 1, 2)          # control on a piece bit, target inside
        sb.c_apply(Gate.P(0.9), s, n - 2, n - 1)       # diagonal, control and target on piece bits
        sb.mc_apply(Gate.X, s, [n - 1, 3], None, 7)
        sb.c_apply(Gate.RZ(0.2), s, 4, n - 1)          # diagonal on a piece bit under a local control

    hre.array[:] = cpu.reals
    him.array[:] = cpu.imags
    a = sb.State(n)
    a.upload_from(hre, him)
    sweep(a)
    want_re, want_im = a.download()
    b = sb.State(n)
    for rep in range(2):
        b.upload_async(hre, him)
        sweep(b)
        b.download_into(ore, oim)
        assert np.array_equal(ore.array, want_re) and np.array_equal(oim.array, want_im), rep
    # after the round trip the state is an ordinary one again
    sb.apply(Gate.H, b, 0)
    sb.apply(Gate.H, a, 0)
    assert np.array_equal(a.download()[0], b.download()[0])


# ---- sampling ------------------------------------------------------------------------------------------------
def test_sample_basis_state_is_exact():  # core.rs:272-291
    n = 3
    for i in range(1 << n):
        s = sb.State(n)
        s.set_basis(i)
        out = sb.sample(s, 8, seed=i)
        assert np.all(out == i)


@pytest.mark.parametrize("n", [4, 13, 18])
def test_sample_matches_oracle_cdf(n):
    cpu = orc.gen_random_state(n, 321 + n)
    gpu = to_gpu(cpu)
    u = orc.uniforms(42, 4096)
    got = sb.sample(gpu, len(u), u01=u)
    want = orc.sample_cdf(cpu, u)
    bad = np.nonzero(got != want)[0]
    if len(bad):  # only legal when u sits within rounding of a CDF boundary
        cdf = np.cumsum(cpu.reals ** 2 + cpu.imags ** 2)
        for k in bad:
            assert abs(int(got[k]) - int(want[k])) == 1
            b = min(got[k], want[k])
            assert abs(cdf[b] - u[k] * cdf[-1]) < 1e-12
    assert len(bad) <= 2


def test_sample_peaked_distribution_many_shots_matches_oracle_cdf():
    """Most of 2^17 shots land in one or two blocks of 4096 amplitudes: the warp-aggregated histogram, and the extra pass-2 CTAs
    of blocks that hold more than 1024 shots (k_hot_list), against the oracle's CDF walk."""
    n = 15
    cpu = orc.gen_random_state(n, 77)
    w = np.full(1 << n, 0.02 / (1 << n))
    w[5000] = 0.6
    w[5001] = 0.1
    w[20000] = 0.28
    phase = np.arctan2(cpu.imags, cpu.reals)
    amp = np.sqrt(w / w.sum())
    cpu.reals[:] = amp * np.cos(phase)
    cpu.imags[:] = amp * np.sin(phase)
    gpu = to_gpu(cpu)
    u = orc.uniforms(7, 1 << 17)
    got = sb.sample(gpu, len(u), u01=u)
    want = orc.sample_cdf(cpu, u)
    bad = np.nonzero(got != want)[0]
    cdf = np.cumsum(cpu.reals ** 2 + cpu.imags ** 2)
    for k in bad:  # only legal when u sits within rounding of a CDF boundary
        assert abs(int(got[k]) - int(want[k])) == 1
        assert abs(cdf[min(got[k], want[k])] - u[k] * cdf[-1]) < 1e-12
    assert len(bad) <= 8
    counts = np.bincount(got, minlength=1 << n)
    assert abs(counts[5000] / len(u) - 0.6) < 0.01 and abs(counts[20000] / len(u) - 0.28) < 0.01


def test_sample_chi_square():
    n = 6
    cpu = orc.gen_random_state(n, 99)
    gpu = to_gpu(cpu)
    shots = 1 << 16
    out = sb.sample(gpu, shots, seed=1)
    counts = np.bincount(out, minlength=1 << n)
    p = cpu.reals ** 2 + cpu.imags ** 2
    chi2 = np.sum((counts - shots * p) ** 2 / (shots * p))
    assert chi2 < 63 + 6 * math.sqrt(2 * 63)  # dof = 63


def test_init_random_matches_oracle_recipe():  # utils.rs:168-201
    n = 10
    s = sb.State(n)
    s.init_random(42)
    cpu = orc.gen_random_state(n, 42)
    assert_same(s, cpu, exact=False, tol=1e-14)
    assert abs(sb.norm2(s) - 1.0) < 1e-12


def test_clone_and_partial_transfer():
    cpu = orc.gen_random_state(6, 4)
    a = to_gpu(cpu)
    b = a.clone()
    sb.apply(Gate.X, a, 0)
    assert_same(b, cpu)
    re, im = b.download(8, 16)
    assert np.array_equal(re, cpu.reals[8:24])
    b.upload(np.zeros(4), np.ones(4), offset=4)
    assert np.array_equal(b.download(4, 4)[1], np.ones(4))


# ---- full-size properties (BASELINE sizes; no CPU state needed) ------------------------------------------------
@pytest.mark.parametrize("n", [26])
def test_large_state_properties(n):
    s = sb.State(n)
    s.init_random(42)
    assert abs(sb.norm2(s) - 1.0) < 1e-10
    head0 = s.download(0, 64)
    # G then G^-1 on every target: exact for H/X/Y/Z pairs only up to rounding -> compare at 1e-12
    for t in (0, 1, 2, 5, n // 2, n - 2, n - 1):
        for g in (Gate.H, Gate.RX(1.0), Gate.RZ(1.0), Gate.U(0.3, 0.2, 0.1)):
            sb.apply(g, s, t)
            sb.apply(g.inverse(), s, t)
    head1 = s.download(0, 64)
    assert np.max(np.abs(head0[0] - head1[0])) < 1e-12 and np.max(np.abs(head0[1] - head1[1])) < 1e-12
    assert abs(sb.norm2(s) - 1.0) < 1e-10
    # X on every qubit maps index i to ~i: exact permutation
    tail0 = s.download((1 << n) - 64, 64)
    for t in range(n):
        sb.apply(Gate.X, s, t)
    flipped = s.download(0, 64)
    assert np.array_equal(flipped[0], tail0[0][::-1]) and np.array_equal(flipped[1], tail0[1][::-1])
    # QFT . IQFT = identity at scale, fused
    s.set_basis(12345)
    qc = QuantumCircuit.from_state(s, fuse=True)
    qc.qft()
    qc.iqft(list(reversed(range(n))))
    qc.execute()
    assert abs(s.amp(12345) - 1.0) < 1e-12
    assert abs(sb.norm2(s) - 1.0) < 1e-10


def test_spynoza_facade_runs():  # spynoza/examples/*.py shape
    from spinoza_b200 import spynoza as sp
    q = sp.QuantumRegister(3)
    qc = sp.QuantumCircuit(q)
    qc.h(0); qc.cx(0, 1); qc.cx(1, 2)
    st = sp.run(qc)
    assert len(st) == 8
    assert abs(st[0][0] - math.sqrt(0.5)) < 1e-12 and abs(st[7][0] - math.sqrt(0.5)) < 1e-12 and st[3] == (0.0, 0.0)
    hist = sp.get_samples(st, 2000, 20000, seed=3)
    assert set(hist) == {0, 7} and sum(hist.values()) == 2000 and abs(hist[0] - 1000) < 150
    assert "Outcome" in sp.show_table(st) and str(st).count("\n") == 8
    assert abs(sp.qubit_expectation_value(st.data, 0)) < 1e-12
    assert abs(sp.xyz_expectation_value("z", st.data, [1])[0]) < 1e-12


def test_baseline_size_30_qubits_properties():
    """BASELINE config 2's register size (2^30 amplitudes, 17 GB): size-independent properties only."""
    n = 30
    free, _total = sb.mem_info(0)
    if free < 40 << 30:
        pytest.skip("needs ~35 GB of free HBM")
    s = sb.State(n)
    x = 0x9E3779B97F4A7C15 % (1 << n)
    s.set_basis(x)
    qc = QuantumCircuit.from_state(s, fuse=True)
    qc.qft()
    qc.execute()
    re, im = s.download(0, 1 << 12)
    k = np.arange(1 << 12, dtype=np.uint64)
    rev = np.zeros_like(k)
    for b in range(n):
        rev |= ((k >> np.uint64(b)) & np.uint64(1)) << np.uint64(n - 1 - b)
    ph = ((np.uint64(x) * rev) & np.uint64((1 << n) - 1)).astype(np.float64) / float(1 << n)
    want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ph)
    assert np.max(np.abs((re + 1j * im) - want)) < 1e-12
    assert abs(sb.norm2(s) - 1.0) < 1e-10
    for t in (0, 1, 2, 15, 28, 29):  # every kernel shape of the bandwidth sweep: G then G^-1
        for g in (Gate.H, Gate.RX(1.0), Gate.RZ(1.0)):
            sb.apply(g, s, t)
            sb.apply(g.inverse(), s, t)
    re2, im2 = s.download(0, 1 << 12)
    assert np.max(np.abs(re2 - re)) < 1e-12 and np.max(np.abs(im2 - im)) < 1e-12
    qc.iqft(list(reversed(range(n))))
    qc.execute()
    assert abs(s.amp(x) - 1.0) < 1e-12 and abs(sb.norm2(s) - 1.0) < 1e-10
    assert abs(sb.prob0(s, 3) - (0.0 if (x >> 3) & 1 else 1.0)) < 1e-10
