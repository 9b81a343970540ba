"""Independent NumPy statement of 'apply a 2x2 matrix to qubit t under a control mask'.

Used to cross-check the C oracle (a third opinion next to the reference's golden vectors).
Matrices follow Gate::to_matrix (gates.rs:95-190).
"""
import numpy as np

H, M, X, Y, Z, P, RX, RY, RZ, SWAP, U, UNITARY, BITFLIP = range(13)


def matrix(kind, p=()):
    r = 1.0 / np.sqrt(2.0)
    if kind == H:
        return np.array([[r, r], [r, -r]], dtype=complex)
    if kind == X:
        return np.array([[0, 1], [1, 0]], dtype=complex)
    if kind == Y:
        return np.array([[0, -1j], [1j, 0]], dtype=complex)
    if kind == Z:
        return np.array([[1, 0], [0, -1]], dtype=complex)
    if kind == P:
        return np.array([[1, 0], [0, np.exp(1j * p[0])]], dtype=complex)
    if kind == RX:
        c, s = np.cos(p[0] / 2), np.sin(p[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
    if kind == RY:
        c, s = np.cos(p[0] / 2), np.sin(p[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=complex)
    if kind == RZ:
        return np.array([[np.exp(-1j * p[0] / 2), 0], [0, np.exp(1j * p[0] / 2)]], dtype=complex)
    if kind == U:
        th, ph, la = p
        c, s = np.cos(th / 2), np.sin(th / 2)
        return np.array([[c, -np.exp(1j * la) * s], [np.exp(1j * ph) * s, np.exp(1j * (ph + la)) * c]],
                        dtype=complex)
    raise ValueError(kind)


def apply_matrix(psi, n, m, target, ctrl_mask=0, neg_mask=0):
    """psi: complex vector of length 2^n (qubit q <-> bit q of the index). Returns a new vector.
    Every qubit of ctrl_mask is a control; those also in neg_mask fire on 0 (the signed-controls extension)."""
    idx = np.arange(1 << n, dtype=np.int64)
    sel0 = ((idx >> target) & 1) == 0
    if ctrl_mask:
        sel0 &= (idx & ctrl_mask) == (ctrl_mask & ~neg_mask)
    s0 = idx[sel0]
    s1 = s0 | (1 << target)
    out = psi.copy()
    a, b = psi[s0], psi[s1]
    out[s0] = m[0, 0] * a + m[0, 1] * b
    out[s1] = m[1, 0] * a + m[1, 1] * b
    return out


def apply_swap(psi, n, t0, t1):
    idx = np.arange(1 << n, dtype=np.int64)
    b0, b1 = (idx >> t0) & 1, (idx >> t1) & 1
    j = idx ^ ((b0 ^ b1) << t0) ^ ((b0 ^ b1) << t1)
    return psi[j]


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    p = rng.random(1 << n)
    p /= p.sum()
    a = rng.random(1 << n) * 2 * np.pi
    return np.sqrt(p) * np.exp(1j * a)
