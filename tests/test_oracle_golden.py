"""Pins the CPU oracle against the reference's own known-answer tests.

Every expected number below is copied (as a test vector, with its file:line) from the
`#[cfg(test)]` modules of /root/reference/spinoza/src/*.rs.  Tolerances are the reference's own.
"""
import math

import numpy as np
import pytest

import oracle as orc
from tests import _dense as D

PI = math.pi


def close(actual, expected, eps):
    # utils.rs:163-165 assert_float_closeness
    assert abs(actual - expected) < eps, (actual, expected, eps)


def apply_all(kind, n, params=()):
    s = orc.State(n)
    for t in range(n):
        orc.apply(kind, s, t, params)
    return s


# ---- gates.rs:1531-1628 ---------------------------------------------------------------------
def test_h_gate_3_qubits():  # gates.rs:1532-1544
    s = apply_all(orc.H, 3)
    for i in range(8):
        close(s.reals[i], 0.35355339059327384, 1e-10)
        close(s.imags[i], 0.0, 1e-10)


@pytest.mark.parametrize("n", [3, 20])
def test_x_gate(n):  # gates.rs:1547-1576
    s = apply_all(orc.X, n)
    close(s.reals[0], 0.0, 1e-10)
    close(s.imags[0], 0.0, 1e-10)
    close(s.reals[(1 << n) - 1], 1.0, 1e-10)
    close(s.imags[(1 << n) - 1], 0.0, 1e-10)


def test_y_gate_3_qubits():  # gates.rs:1579-1592
    s = apply_all(orc.Y, 3)
    close(s.reals[0], 0.0, 1e-10)
    close(s.imags[0], 0.0, 1e-10)
    close(s.reals[7], 0.0, 1e-10)
    close(s.imags[7], -1.0, 1e-10)


def test_y_gate_20_qubits():  # gates.rs:1595-1608
    n = 20
    s = apply_all(orc.Y, n)
    close(s.reals[0], 0.0, 1e-10)
    close(s.imags[0], 0.0, 1e-10)
    close(s.reals[(1 << n) - 1], 1.0, 1e-10)
    close(s.imags[(1 << n) - 1], 0.0, 1e-10)


@pytest.mark.parametrize("kind,params", [(orc.Z, ()), (orc.P, (PI,))])
def test_z_and_p_gate_3_qubits(kind, params):  # gates.rs:1611-1648
    s = apply_all(kind, 3, params)
    for i in range(8):
        close(s.reals[i], 1.0 if i == 0 else 0.0, 1e-10)
        close(s.imags[i], 0.0, 1e-10)


def test_rx_gate_3_qubits():  # gates.rs:1651-1682
    s = apply_all(orc.RX, 3, (1.0,))
    a, b, c, d = 0.6758712218347053, -0.3692301313020644, -0.20171134005566746, 0.11019540730213864
    exp = [(a, 0), (0, b), (0, b), (c, 0), (0, b), (c, 0), (c, 0), (0, d)]
    for i, (re, im) in enumerate(exp):
        close(s.reals[i], re, 1e-10)
        close(s.imags[i], im, 1e-10)


def test_ry_gate_3_qubits():  # gates.rs:1685-1716
    s = apply_all(orc.RY, 3, (1.0,))
    a, b, c, d = 0.6758712218347053, 0.3692301313020644, 0.20171134005566746, 0.11019540730213864
    exp = [a, b, b, c, b, c, c, d]
    for i, re in enumerate(exp):
        close(s.reals[i], re, 1e-10)
        close(s.imags[i], 0.0, 1e-10)


def test_rz_gate_3_qubits():  # gates.rs:1719-1750
    s = apply_all(orc.RZ, 3, (1.0,))
    close(s.reals[0], 0.07073720166770296, 1e-10)
    close(s.imags[0], -0.9974949866040546, 1e-10)
    for i in range(1, 8):
        close(s.reals[i], 0.0, 1e-10)
        close(s.imags[i], 0.0, 1e-10)


def test_u_gate_3_qubits():  # gates.rs:1812-1843
    s = apply_all(orc.U, 3, (1.0, 1.0, 1.0))
    close(s.reals[0], 0.6758712218347053, 1e-10)
    close(s.imags[0], 0.0, 1e-10)
    for i in (1, 2):
        close(s.reals[i], 0.19949589133850137, 1e-10)
        close(s.imags[i], 0.3106964422074971, 1e-10)
    close(s.reals[3], -0.08394153605985091, 1e-10)
    close(s.imags[3], 0.18341560247417849, 1e-10)
    close(s.reals[7], -0.1090926263889472, 1e-10)
    close(s.imags[7], 0.015550776766638148, 1e-10)


def test_u_gate_1_qubit():  # gates.rs:1846-1870
    theta, phi, lam = 2.0, 3.0, 1.0
    s = orc.State(1)
    orc.apply(orc.U, s, 0, (theta, phi, lam))
    close(s.reals[0], 0.5403023058681398, 1e-10)
    close(s.imags[0], 0.0, 1e-10)
    close(s.reals[1], -0.833049961066805, 1e-10)
    close(s.imags[1], 0.11874839215823475, 1e-10)
    s = orc.State(1)
    orc.apply(orc.X, s, 0)
    orc.apply(orc.U, s, 0, (theta, phi, lam))
    close(s.reals[0], -0.4546487134128409, 1e-10)
    close(s.imags[0], -0.7080734182735712, 1e-10)
    close(s.reals[1], -0.35316515556860967, 1e-10)
    close(s.imags[1], -0.4089021333016357, 1e-10)


def qcbm_functional(n):  # gates.rs:1499-1529
    s = orc.State(n)
    pairs = [(i, (i + 1) % n) for i in range(n)]
    for i in range(n):
        orc.apply(orc.RX, s, i, (1.0,))
        orc.apply(orc.RZ, s, i, (1.0,))
    for p0, p1 in pairs[: n - 1]:
        orc.c_apply(orc.X, s, p0, p1)
    for _ in range(9):
        for i in range(n):
            orc.apply(orc.RZ, s, i, (1.0,))
            orc.apply(orc.RX, s, i, (1.0,))
            orc.apply(orc.RZ, s, i, (1.0,))
        for p0, p1 in pairs[: n - 1]:
            orc.c_apply(orc.X, s, p0, p1)
    for i in range(n):
        orc.apply(orc.RZ, s, i, (1.0,))
        orc.apply(orc.RX, s, i, (1.0,))
    return s


def test_qcbm_3_qubits():  # gates.rs:1787-1795
    s = qcbm_functional(3)
    close(s.reals[0], 0.18037770683997864, 1e-10)
    close(s.imags[0], -0.17626993141958947, 1e-10)
    close(s.reals[7], 0.014503954556966365, 1e-10)
    close(s.imags[7], -0.11198008105074927, 1e-10)


@pytest.mark.parametrize("threads", [1, 4])
def test_qcbm_20_qubits(threads):  # gates.rs:1798-1809 (exercises the rayon-mirroring OpenMP path at threads=4)
    orc.set_threads(threads)
    try:
        s = qcbm_functional(20)
    finally:
        orc.set_threads(1)
    close(s.reals[0], -0.0022221321676945643, 1e-10)
    close(s.imags[0], 0.001743068112560825, 1e-10)
    close(s.reals[7], -0.0031017461877124453, 1e-10)
    close(s.imags[7], -0.0034043237120339686, 1e-10)
    close(s.reals[12], 0.0005494086136357235, 1e-10)
    close(s.imags[12], -0.00009827749580581964, 1e-10)


def ref_swap_3cx(s, a, b):  # utils.rs:204-208
    orc.c_apply(orc.X, s, a, b)
    orc.c_apply(orc.X, s, b, a)
    orc.c_apply(orc.X, s, a, b)


def test_swap_9_qubits():  # gates.rs:1873-1885 (bit-exact)
    s0 = orc.gen_random_state(9, 1)
    s1 = s0.clone()
    ref_swap_3cx(s0, 0, 1)
    orc.swap(s1, 0, 1)
    assert np.array_equal(s0.reals, s1.reals) and np.array_equal(s0.imags, s1.imags)


@pytest.mark.parametrize("n", [3, 9])
def test_swap_all_qubits(n):  # gates.rs:1888-1901, circuit.rs:892-926 (bit-exact)
    s0 = orc.gen_random_state(n, 2)
    s1 = s0.clone()
    for i in range(n >> 1):
        ref_swap_3cx(s0, i, n - 1 - i)
        orc.apply(orc.SWAP, s1, 0, t0=i, t1=n - 1 - i)
    assert np.array_equal(s0.reals, s1.reals) and np.array_equal(s0.imags, s1.imags)


@pytest.mark.parametrize("kind,params,inv", [
    (orc.H, (), ()), (orc.X, (), ()), (orc.Y, (), ()), (orc.Z, (), ()),
    (orc.P, (2.03,), (-2.03,)), (orc.RX, (2.03,), (-2.03,)), (orc.RZ, (3.03,), (-3.03,)),
    (orc.RY, (3.03,), (-3.03,)), (orc.U, (1.0, 2.0, 3.0), (-1.0, -3.0, -2.0)),
])
def test_gate_times_inverse_is_identity(kind, params, inv):
    # gates.rs:1904-2045 (matrix identity at 1e-3); here on states: G^-1 G psi == psi
    s = orc.gen_random_state(4, 3)
    ref = s.clone()
    for t in range(4):
        orc.apply(kind, s, t, params)
        orc.apply(kind, s, t, inv)
    assert np.max(np.abs(s.amps() - ref.amps())) < 1e-12


def test_ch():  # gates.rs:2100-2133
    s = apply_all(orc.H, 3)
    orc.c_apply(orc.H, s, 0, 1)
    for i in range(0, 8, 4):
        for k, v in enumerate([0.353553391, 0.5, 0.353553391, 0.0]):
            close(s.reals[i + k], v, 1e-4)
            close(s.imags[i + k], 0.0, 1e-4)


def test_crz():  # gates.rs:2136-2171
    s = apply_all(orc.H, 3)
    orc.c_apply(orc.RZ, s, 0, 1, (PI / 2.0,))
    for i in range(0, 8, 4):
        exp = [(0.353553391, 0.0), (0.25, -0.25), (0.353553391, 0.0), (0.25, 0.25)]
        for k, (re, im) in enumerate(exp):
            close(s.reals[i + k], re, 1e-4)
            close(s.imags[i + k], im, 1e-4)


def test_controlled_u():  # gates.rs:2198-2236
    s = apply_all(orc.H, 3)
    orc.c_apply(orc.U, s, 0, 1, (1.0, 2.0, 3.0))
    er = [0.35355339, 0.47807852, 0.35355339, 0.01747458] * 2
    ei = [0.0, -0.0239202, 0.0, -0.14339942] * 2
    for i in range(8):
        close(s.reals[i], er[i], 1e-5)
        close(s.imags[i], ei[i], 1e-5)


def test_bit_flip_noise():  # gates.rs:2174-2195 through execute
    s0 = orc.gen_random_state(1, 5)
    s1 = s0.clone()
    orc.execute(s1, [orc.make_op(orc.BITFLIP, 0, (0.0,))], u01=[0.3])
    assert np.array_equal(s0.reals, s1.reals)
    orc.execute(s1, [orc.make_op(orc.BITFLIP, 0, (1.0,))], u01=[0.3])
    assert s1.reals[0] == s0.reals[1] and s1.reals[1] == s0.reals[0]
    assert s1.imags[0] == s0.imags[1] and s1.imags[1] == s0.imags[0]


# ---- core.rs:271-339 ------------------------------------------------------------------------
def test_xyz_exp_val():  # core.rs:294-301, 329-339
    s = orc.State(1)
    orc.apply(orc.RX, s, 0, (0.54,))
    orc.apply(orc.RY, s, 0, (0.12,))
    v = orc.xyz_expectation_value("z", s, [0])
    close(v[0], 0.8515405859048367, 1e-4)
    close(orc.qubit_expectation_value(s, 0), v[0], 1e-4)
    orc.xyz_expectation_value("x", s, [0])
    orc.xyz_expectation_value("y", s, [0])
    with pytest.raises(orc.OracleError):  # core.rs:303-310 panics
        orc.xyz_expectation_value("a", s, [0])


def test_reservoir_encoded_integers():  # core.rs:272-291
    n = 3
    for i in range(1 << n):
        s = orc.State(n)
        s.reals[0] = 0.0
        s.reals[i] = 1.0
        entries = orc.reservoir_sampling(s, len(s), len(s) * 10_000, seed=7 + i)
        assert np.all(entries == i)


def test_value_encoding_iqft():  # circuit.rs:1076-1113 (|4> at 1e-4), gates.rs:1753-1767
    n, v = 3, 4.0
    s = orc.State(n)
    for t in range(n):
        orc.apply(orc.H, s, t)
    for t in range(n):
        orc.apply(orc.P, s, t, (2.0 * PI / (2.0 ** (t + 1)) * v,))
    orc.iqft(s, list(range(n))[::-1])
    for i in range(1 << n):
        close(s.reals[i], 1.0 if i == int(v) else 0.0, 1e-4)
        close(s.imags[i], 0.0, 1e-4)


# ---- measurement.rs:100-246 ------------------------------------------------------------------
def test_measure_qubit_norm():  # measurement.rs:101-142
    s = orc.gen_random_state(3, 11)
    close(orc.norm2(s), 1.0, 1e-3)
    for t, v in [(0, 0), (1, 0), (2, 1)]:
        bit, _ = orc.measure_qubit(s, t, True, v)
        assert bit == v
        close(orc.norm2(s), 1.0, 1e-3)


def test_measure_qubit_known_state():  # measurement.rs:145-246
    vals = [0.034172256444052966, 0.29007027387615136, -0.1300556493088507, 0.47222164829858637,
            -0.032338373524095645, 0.26511510737291843, 0.1259630181898572, -0.09645897805840803,
            -0.31931099330088214, -0.24644972468157703, -0.15963222942036193, -0.14329373536970438,
            -0.1564141838467382, -0.4751067410290973, 0.1034273381193853, -0.32966556091031934]
    s = orc.State(3, vals[0::2], vals[1::2])
    eps = 1e-3
    orc.measure_qubit(s, 0, True, 0)
    exp = {0: (0.04528096797370981, 0.38436627101331156), 2: (-0.042850926694402595, 0.3512986830692283),
           4: (-0.42311255872092046, -0.32656556082875193), 6: (-0.2072612811212442, -0.6295543626114914)}
    for i in range(8):
        re, im = exp.get(i, (0.0, 0.0))
        close(s.reals[i], re, eps)
        close(s.imags[i], im, eps)
    orc.measure_qubit(s, 1, True, 0)
    exp = {0: (0.06861878352538178, 0.5824686866330654), 4: (-0.6411848150109799, -0.49487748447346463)}
    for i in range(8):
        re, im = exp.get(i, (0.0, 0.0))
        close(s.reals[i], re, eps)
        close(s.imags[i], im, eps)
    orc.measure_qubit(s, 2, True, 1)  # outcome 1 + reset -> X moves the survivor to index 0
    exp = {0: (-0.7916334352111761, -0.6109963209838112)}
    for i in range(8):
        re, im = exp.get(i, (0.0, 0.0))
        close(s.reals[i], re, eps)
        close(s.imags[i], im, eps)


# ---- circuit.rs:622-1250 through orc.execute ---------------------------------------------------
def op(kind, target, params=(), control=None, **kw):
    if control is None:
        return orc.make_op(kind, target, params, **kw)
    return orc.make_op(kind, target, params, ctrl_kind=orc.CTRL_SINGLE, ctrl_mask=1 << control, **kw)


def test_all_gates_as_transformations():  # circuit.rs:771-822 (n = 17, exercises n >= 15 paths)
    n = 17
    ops = [op(orc.H, t) for t in range(n)]
    ops += [op(orc.X, 0), op(orc.Y, 1), op(orc.Z, 2), op(orc.P, 3, (PI,)), op(orc.P, 4, (PI,), control=3),
            op(orc.RX, 5, (PI,)), op(orc.RY, 6, (PI,)), op(orc.RZ, 7, (PI,)), op(orc.U, 8, (PI, PI, PI)),
            op(orc.Y, 10, control=9), op(orc.RX, 12, (PI,), control=11), op(orc.RY, 14, (PI,), control=13)]
    a = orc.State(n)
    orc.execute(a, ops)
    b = orc.State(n)
    for t in range(n):
        orc.apply(orc.H, b, t)
    orc.apply(orc.X, b, 0); orc.apply(orc.Y, b, 1); orc.apply(orc.Z, b, 2); orc.apply(orc.P, b, 3, (PI,))
    orc.c_apply(orc.P, b, 3, 4, (PI,)); orc.apply(orc.RX, b, 5, (PI,)); orc.apply(orc.RY, b, 6, (PI,))
    orc.apply(orc.RZ, b, 7, (PI,)); orc.apply(orc.U, b, 8, (PI, PI, PI)); orc.c_apply(orc.Y, b, 9, 10)
    orc.c_apply(orc.RX, b, 11, 12, (PI,)); orc.c_apply(orc.RY, b, 13, 14, (PI,))
    assert np.array_equal(a.reals, b.reals) and np.array_equal(a.imags, b.imags)
    # and against an independent dense NumPy statement
    psi = np.zeros(1 << n, dtype=complex); psi[0] = 1
    for t in range(n):
        psi = D.apply_matrix(psi, n, D.matrix(D.H), t)
    seq = [(D.X, 0, (), 0), (D.Y, 1, (), 0), (D.Z, 2, (), 0), (D.P, 3, (PI,), 0), (D.P, 4, (PI,), 1 << 3),
           (D.RX, 5, (PI,), 0), (D.RY, 6, (PI,), 0), (D.RZ, 7, (PI,), 0), (D.U, 8, (PI, PI, PI), 0),
           (D.Y, 10, (), 1 << 9), (D.RX, 12, (PI,), 1 << 11), (D.RY, 14, (PI,), 1 << 13)]
    for kind, t, p, cm in seq:
        psi = D.apply_matrix(psi, n, D.matrix(kind, p), t, cm)
    assert np.max(np.abs(a.amps() - psi)) < 1e-12


def test_measure_all_twice_gives_same_bits():  # circuit.rs:825-889
    n = 12
    s = orc.gen_random_state(n, 21)
    ops = [op(orc.M, t) for t in range(n)]
    u = orc.uniforms(99, n)
    m, v = orc.execute(s, ops, u01=u)
    assert m == (1 << n) - 1
    m2, v2 = orc.execute(s, ops, measured=m, vals=v, u01=orc.uniforms(5, n))
    assert (m2, v2) == (m, v)
    # reset=true => state collapsed to |0..0>
    close(s.reals[0] ** 2 + s.imags[0] ** 2, 1.0, 1e-9)


def test_inverse_iqft_roundtrip():  # circuit.rs:929-960
    n = 5
    s = orc.gen_random_state(n, 31)
    ref = s.clone()
    targets = list(range(n))[::-1]
    fwd = []
    for j in reversed(range(n)):  # circuit.rs:438-445
        fwd.append(op(orc.H, targets[j]))
        for k in reversed(range(j)):
            fwd.append(op(orc.P, targets[k], (-PI / 2.0 ** (j - k),), control=targets[j]))
    orc.execute(s, fwd)
    inv = []
    for o in reversed(fwd):  # circuit.rs:206-211 + gates.rs:78-92
        q = orc.make_op(o.kind, o.target, [-x for x in o.p], ctrl_kind=o.ctrl_kind, ctrl_mask=o.ctrl_mask)
        inv.append(q)
    orc.execute(s, inv)
    assert np.max(np.abs(s.amps() - ref.amps())) < 1e-12


def test_execute_controlled_u_matches_functional():  # circuit.rs:1235-1250 (bit-exact)
    a = orc.State(3)
    orc.execute(a, [op(orc.U, 1, (1.0, 2.0, 3.0), control=0)])
    b = orc.State(3)
    orc.c_apply(orc.U, b, 0, 1, (1.0, 2.0, 3.0))
    assert np.array_equal(a.reals, b.reals) and np.array_equal(a.imags, b.imags)


def test_classical_control_after_measurement():  # circuit.rs:570-574
    s = orc.State(2)
    ops = [op(orc.X, 0), op(orc.M, 0), op(orc.X, 1, control=0)]
    m, v = orc.execute(s, ops, u01=[0.5])
    assert m == 1 and v == 1
    # qubit 0 measured 1 then reset to 0; the classically-controlled X fired on qubit 1 -> |10>
    close(s.reals[2], 1.0, 1e-12)


# ---- reference defects: intended semantics vs literal loop -------------------------
def test_mc_literal_scan_agrees_on_safe_domain_and_fails_outside():
    # B2: the literal scan-and-skip loop is correct iff no non-control qubit lies below the target.
    agree = disagree = 0
    for n in range(2, 6):
        for target in range(n):
            for mask in range(1, 1 << n):
                if (mask >> target) & 1 or bin(mask).count("1") >= n:
                    continue
                s0 = orc.gen_random_state(n, 100 + n)
                s1 = s0.clone()
                orc.mc_apply_mask(orc.X, s0, mask, target)
                rc = orc.mc_scan_literal(orc.X, s1, mask, target)
                safe = all(((mask >> q) & 1) for q in range(target))
                same = rc == 0 and np.array_equal(s0.reals, s1.reals) and np.array_equal(s0.imags, s1.imags)
                if safe:
                    assert same, (n, target, mask)
                    agree += 1
                else:
                    assert not same, (n, target, mask)
                    disagree += 1
    assert agree > 0 and disagree > 0
    # the cases the reference itself ships are inside the safe domain
    for n, mask, target in [(3, 0b011, 2), (3, 0b101, 1), (3, 0b110, 0)]:
        s0 = orc.gen_random_state(n, 7); s1 = s0.clone()
        orc.mc_apply_mask(orc.P, s0, mask, target, (3.14,))
        assert orc.mc_scan_literal(orc.P, s1, mask, target, (3.14,)) == 0
        assert np.array_equal(s0.reals, s1.reals)


def test_mc_ry_is_true_ry_not_rx():  # B3
    n, mask, target = 3, 0b011, 2
    s = orc.gen_random_state(n, 9)
    psi = s.amps()
    orc.mc_apply_mask(orc.RY, s, mask, target, (0.7,))
    want = D.apply_matrix(psi, n, D.matrix(D.RY, (0.7,)), target, mask)
    assert np.max(np.abs(s.amps() - want)) < 1e-14
    lit = orc.State(n, psi.real, psi.imag)
    assert orc.mc_scan_literal(orc.RY, lit, mask, target, (0.7,)) == 0
    rx = D.apply_matrix(psi, n, D.matrix(D.RX, (0.7,)), target, mask)
    assert np.max(np.abs(lit.amps() - rx)) < 1e-14  # the literal reference loop applies RX


def test_mc_zeros_are_dropped_from_mask():  # B4 mirrored (gates.rs:298-311)
    s0 = orc.gen_random_state(4, 13); s1 = s0.clone()
    orc.mc_apply(orc.X, s0, [0, 1, 2], {1}, 3)
    orc.mc_apply_mask(orc.X, s1, 0b101, 3)
    assert np.array_equal(s0.reals, s1.reals) and np.array_equal(s0.imags, s1.imags)


def test_unsupported_combinations_report_errors():  # gates.rs:230,267,275,318
    s = orc.State(3)
    for kind in (orc.Z, orc.SWAP, orc.M):
        with pytest.raises(orc.OracleError):
            orc.c_apply(kind, s, 0, 1)
    with pytest.raises(orc.OracleError):
        orc.cc_apply(orc.H, s, 0, 1, 2)
    with pytest.raises(orc.OracleError):
        orc.mc_apply(orc.H, s, [0, 1], None, 2)
    with pytest.raises(orc.OracleError):
        orc.apply(orc.M, s, 0)


# ---- oracle vs independent dense statement, all gates x all positions ---------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 8])
def test_oracle_vs_dense_all_gates(n):
    gates = [(orc.H, ()), (orc.X, ()), (orc.Y, ()), (orc.Z, ()), (orc.P, (0.37,)), (orc.RX, (1.1,)),
             (orc.RY, (-0.6,)), (orc.RZ, (2.2,)), (orc.U, (0.3, 1.4, -0.8))]
    for kind, p in gates:
        for t in range(n):
            s = orc.gen_random_state(n, 40 + t)
            want = D.apply_matrix(s.amps(), n, D.matrix(kind, p), t)
            orc.apply(kind, s, t, p)
            assert np.max(np.abs(s.amps() - want)) < 1e-14, (kind, t)
            if kind == orc.Z:
                continue
            for c in range(n):
                if c == t:
                    continue
                s = orc.gen_random_state(n, 50 + c)
                want = D.apply_matrix(s.amps(), n, D.matrix(kind, p), t, 1 << c)
                orc.c_apply(kind, s, c, t, p)
                assert np.max(np.abs(s.amps() - want)) < 1e-14, (kind, c, t)


def test_qft_closed_form():  # IQFT|x>[k] = 2^(-n/2) exp(-2 pi i rev(x) k / 2^n)
    n = 6
    x = 0x9E3779B97F4A7C15 % (1 << n)
    s = orc.State(n)
    s.reals[0] = 0.0
    s.reals[x] = 1.0
    orc.iqft(s, list(range(n))[::-1])
    rev = int(format(x, f"0{n}b")[::-1], 2)
    k = np.arange(1 << n)
    want = 2.0 ** (-n / 2) * np.exp(-2j * np.pi * rev * k / (1 << n))
    assert np.max(np.abs(s.amps() - want)) < 1e-13
