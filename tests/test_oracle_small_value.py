"""The oracle's restatement of the reference's small-value path (src/big_num/small_value.rs; users in
src/neutronnova_zk.rs:255-325, 649-693) against the definitions on Python integers — the way the reference's own
`test_small_value!` property tests pin it (small_value.rs:224-330): classification thresholds, signed accumulation of
field * i128 products, reduction of the 448-bit sums, and agreement of the i64 NIFS round 0 / c_vals with the standard
field path on the same layers (with and without large positions)."""
import numpy as np

Q = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
SMALL_MAX = (1 << 62) - 1


def rand_fe(rng, n):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a


def test_to_small_vec_or_zero_thresholds(orc):
    vals = [0, 1, 5, SMALL_MAX, SMALL_MAX + 1, Q - 1, Q - SMALL_MAX, Q - SMALL_MAX - 1, 1 << 64, Q // 2, Q - 7]
    out, large = orc.to_small_vec_or_zero(orc.to_mont(vals))
    want = [0, 1, 5, SMALL_MAX, 0, -1, -SMALL_MAX, 0, 0, 0, -7]
    assert out.tolist() == want
    assert large.tolist() == [4, 7, 8, 9]


def test_small_accumulator_matches_integers(orc):
    rng = np.random.default_rng(11)
    for n, mag in ((1, 1), (7, 3), (64, 62), (300, 126), (2000, 126)):
        f = rand_fe(rng, n)
        fi = [orc.limbs_to_int(x) for x in f]                 # raw Montgomery limbs as integers: the accumulator sums f_mont * val
        vals = [int(rng.integers(0, 2**63)) * int(rng.integers(0, 2**63)) >> (126 - mag) for _ in range(n)]
        vals = [v if rng.integers(0, 2) else -v for v in vals]
        vals[0] = 0 if n > 1 else vals[0]
        got = orc.limbs_to_int(orc.small_acc_dot(f, vals)[0])
        assert got == sum(a * v for a, v in zip(fi, vals)) % Q
    # all-positive sums large enough to leave the 4-limb fast path of reduce_7_to_field
    f = rand_fe(rng, 50); vals = [(1 << 126) - 1 - i for i in range(50)]
    assert orc.limbs_to_int(orc.small_acc_dot(f, vals)[0]) == sum(orc.limbs_to_int(a) * v for a, v in zip(f, vals)) % Q


def _layers(orc, rng, n, N, n_big):
    """n layers of N small signed values (as field elements) with n_big full-width entries sprinkled in"""
    ints = rng.integers(-(2**40), 2**40, size=(n, N)).astype(object)
    ints[0, 0] = SMALL_MAX; ints[1, 0] = -SMALL_MAX          # extreme differences: (2V)^2 products
    L = np.concatenate([orc.to_mont([int(v) % Q for v in ints[b]]) for b in range(n)], axis=0)
    big = rand_fe(rng, n_big)
    pos = rng.choice(n * N, size=n_big, replace=False) if n_big else []
    for q, p in enumerate(pos):
        L[int(p)] = big[q]
    return L


def _small_layers(orc, L, n, N):
    out = np.zeros((n, N), dtype=np.int64); union = set()
    for b in range(n):
        v, lg = orc.to_small_vec_or_zero(L[b * N:(b + 1) * N]); out[b] = v; union |= set(int(x) for x in lg)
    lp = np.array(sorted(union), dtype=np.uint64)
    for p in lp:                                              # neutronnova_zk.rs:1575-1584: zero at ALL large positions in ALL layers
        out[:, int(p)] = 0
    return out.reshape(-1), lp


def test_nifs_round0_small_equals_field_path(orc):
    rng = np.random.default_rng(5)
    for n, left, right, n_big in ((2, 4, 2, 0), (4, 8, 4, 3), (8, 16, 8, 9)):
        N = left * right; ell_b = n.bit_length() - 1
        A, B, Cm = _layers(orc, rng, n, N, n_big), _layers(orc, rng, n, N, n_big), _layers(orc, rng, n, N, 0)
        E = orc.pow_split_evals(rand_fe(rng, 1), left, right); rhos = rand_fe(rng, ell_b)
        A64, lpa = _small_layers(orc, A, n, N); B64, lpb = _small_layers(orc, B, n, N)
        lp = np.array(sorted(set(lpa.tolist()) | set(lpb.tolist())), dtype=np.uint64)
        A64 = A64.reshape(n, N); B64 = B64.reshape(n, N)
        for p in lp:
            A64[:, int(p)] = 0; B64[:, int(p)] = 0
        got = orc.nifs_round0_small(rhos, left, right, E, A, B, A64.reshape(-1), B64.reshape(-1), lp, N, n)
        want = orc.nifs_round(0, rhos, left, right, E, A, B, Cm, N, n)
        assert np.array_equal(got, want), (n, n_big)
        assert (len(lp) > 0) == (n_big > 0)


def test_cvals_small_equals_definition(orc):
    rng = np.random.default_rng(6)
    n, left, right = 4, 8, 4; N = left * right
    Cl = _layers(orc, rng, n, N, 5); C64, lp = _small_layers(orc, Cl, n, N)
    E = orc.pow_split_evals(rand_fe(rng, 1), left, right)
    got = orc.nifs_cvals_small(left, right, E, Cl, C64, lp, N, n)
    R_INV = pow(1 << 256, -1, Q)
    Ei = [orc.limbs_to_int(x) * R_INV % Q for x in E]
    for b in range(n):
        c = [orc.limbs_to_int(x) * R_INV % Q for x in Cl[b * N:(b + 1) * N]]
        want = sum(Ei[k % left] * Ei[left + k // left] * c[k] for k in range(N)) % Q
        assert orc.limbs_to_int(got[b]) * R_INV % Q == want
