"""Oracle restatements of the NeutronNova building blocks (PowPolynomial::split_evals, NIFS round evaluation / layer
fold, weights_from_r + fold_multiple, the pow-weighted cubic evaluation points) against their definitions in Python
big ints — mirrors reference tests polys/power.rs:95-161 (outer-product identity) and the definition-based checks
the oracle's other sum-check functions get in tests/test_oracle_sumcheck.py.  CPU only."""
import numpy as np
import pytest

P = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff


def rf(rng, n):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a


@pytest.fixture(scope="module")
def data(orc):
    rng = np.random.default_rng(1)
    left, right = 8, 4; N = left * right; m = 4
    t = rf(rng, 1)
    E = orc.pow_split_evals(t, left, right)
    A, B, Cm = rf(rng, m * N), rf(rng, m * N), rf(rng, m * N)
    return dict(rng=rng, left=left, right=right, N=N, m=m, t=t, E=E, A=A, B=B, C=Cm)


def test_pow_split_evals_outer_product(orc, data):
    fm = orc.from_mont; left, right = data["left"], data["right"]
    ti = fm(data["t"])[0]; Ei = fm(data["E"])
    assert all(Ei[j] * Ei[left + i] % P == pow(ti, i * left + j, P) for i in range(right) for j in range(left))


def test_nifs_round_vs_definition(orc, data):
    fm = orc.from_mont; left, right, N, m = data["left"], data["right"], data["N"], data["m"]
    ell_b = 2; rhos = rf(data["rng"], ell_b)
    Ei, Ai, Bi, Ci, rh = fm(data["E"]), fm(data["A"]), fm(data["B"]), fm(data["C"]), fm(rhos)
    for tt in (0, 1):
        mm = m >> tt
        out = fm(orc.nifs_round(tt, rhos, left, right, data["E"], data["A"], data["B"], data["C"], N, mm))
        e0 = q = 0
        for p in range(mm // 2):
            w = 1; k = p
            for s in range(tt + 1, ell_b):
                w = w * (rh[s] if k & 1 else (1 - rh[s])) % P; k >>= 1
            Ek = lambda k: Ei[k % left] * Ei[left + k // left]   # noqa: E731
            pe = sum(Ek(k) * (Ai[2 * p * N + k] * Bi[2 * p * N + k] - Ci[2 * p * N + k]) for k in range(N)) % P if tt else 0
            pq = sum(Ek(k) * (Ai[(2 * p + 1) * N + k] - Ai[2 * p * N + k]) * (Bi[(2 * p + 1) * N + k] - Bi[2 * p * N + k]) for k in range(N)) % P
            e0 = (e0 + w * pe) % P; q = (q + w * pq) % P
        assert out == [e0, q]


def test_fold_multiple_equals_pairwise_folds(orc, data):
    N = data["N"]; r_bs = rf(data["rng"], 2)
    w = orc.weights_from_r(r_bs, 4)
    f1 = orc.nifs_fold(data["A"], N, 4, r_bs[0:1]); f2 = orc.nifs_fold(f1, N, 2, r_bs[1:2])
    assert np.array_equal(orc.fold_vectors(data["A"], 4, N, w), f2)
    wi = orc.from_mont(w); r = orc.from_mont(r_bs)
    assert wi == [(1 - r[0]) * (1 - r[1]) % P, r[0] * (1 - r[1]) % P, (1 - r[0]) * r[1] % P, r[0] * r[1] % P]   # LSB-first


def test_pow_cubic_eval_vs_definition(orc, data):
    fm = orc.from_mont; left, N = data["left"], data["N"]
    Ei, Ai, Bi, Ci = fm(data["E"]), fm(data["A"]), fm(data["B"]), fm(data["C"])
    pl, pr = data["E"][:left], data["E"][left:]
    for tl in (N, left, left // 2):                     # len >= left (outer product) and len < left (single table)
        out = fm(orc.pow_cubic_eval(pl, pr, data["A"][:tl], data["B"][:tl], data["C"][:tl]))
        ln = tl // 2

        def wt(idx):
            return Ei[idx % left] * Ei[left + idx // left] % P if ln >= left else Ei[idx]
        ev = []
        for X in (0, 2, 3):
            s = 0
            for i in range(ln):
                bnd = lambda T: ((1 - X) * T[i] + X * T[i + ln]) % P   # noqa: E731
                s += ((1 - X) * wt(i) + X * wt(i + ln)) * (bnd(Ai) * bnd(Bi) - bnd(Ci))
            ev.append(s % P)
        assert out == ev, tl


def test_fold_commitments_vs_scalar_muls(orc):
    from tests.curve_util import points
    rng = np.random.default_rng(3)
    n, rows = 4, 3
    pts = points(orc, n * rows, seed=9)
    w = rf(rng, n)
    got = orc.fold_commitments(pts, n, rows, w)
    for r in range(rows):
        acc = np.zeros((1, 8), dtype=np.uint64)
        for i in range(n):
            acc = orc.point_add(acc, orc.scalar_mul(pts[i * rows + r:i * rows + r + 1], w[i:i + 1]))
        assert np.array_equal(got[r:r + 1], acc)
