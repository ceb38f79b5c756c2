"""Host-side scalar algebra of the T256 scalar field on Python integers: the handful of field operations per round
that the reference also performs on the host between its parallel loops (UniPoly interpolation / evaluation,
claim updates; src/polys/univariate.rs:84-153, src/sumcheck.rs:786-917, src/neutronnova_zk.rs:703-735).
Values cross as (1, 4) u64 Montgomery limbs (the ABI layout); everything bulk stays on the device."""
import numpy as np

Q = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
R = 1 << 256
R_INV = pow(R, -1, Q)
TWO_INV = pow(2, -1, Q)
SIX_INV = pow(6, -1, Q)


def to_int(limbs):
    """(…, 4) u64 Montgomery limbs -> canonical integer (first element)."""
    l = np.asarray(limbs, dtype=np.uint64).reshape(-1, 4)[0]
    v = sum(int(l[i]) << (64 * i) for i in range(4))
    return v * R_INV % Q


def to_ints(limbs):
    a = np.asarray(limbs, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(r[i]) << (64 * i) for i in range(4)) * R_INV % Q for r in a]


def from_int(v):
    m = (v % Q) * R % Q
    return np.array([[(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)


def from_ints(vs):
    return np.concatenate([from_int(v) for v in vs], axis=0) if len(vs) else np.zeros((0, 4), dtype=np.uint64)


def unipoly_from_evals(evals):
    """UniPoly::from_evals (univariate.rs:84-120): evaluations at 0, 1, 2[, 3] -> coefficients, low to high."""
    if len(evals) == 3:
        e0, e1, e2 = evals
        c2 = (e2 - 2 * e1 + e0) * TWO_INV % Q
        c1 = (e1 - e0 - c2) % Q
        return [e0 % Q, c1, c2]
    e0, e1, e2, e3 = evals
    c3 = (e3 - 3 * e2 + 3 * e1 - e0) * SIX_INV % Q
    c2 = ((e2 - 2 * e1 + e0) * TWO_INV - 3 * c3) % Q
    c1 = (e1 - e0 - c2 - c3) % Q
    return [e0 % Q, c1, c2, c3]


def unipoly_eval(coeffs, r):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * r + c) % Q
    return acc
