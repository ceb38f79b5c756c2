"""Small-value path on the device vs the oracle's restatement of src/big_num/small_value.rs and its NIFS
users: i64 conversion with the union of large positions zeroed, round-0 quad coefficient from the i64 layers (signed
448-bit accumulation, reduction, field correction at the large positions), c_vals — bit-exact, including the extreme
values +-(2^62 - 1), full-width (large) entries and empty / all-large position sets."""
import numpy as np
import pytest

from tests.gpu_util import ctx, rand_fe  # noqa: F401
from tests.test_oracle_small_value import _layers, SMALL_MAX, Q

pytestmark = pytest.mark.gpu


def _union_small(orc, tabs, n, N):
    outs, union = [], set()
    for T in tabs:
        o = np.zeros((n, N), dtype=np.int64)
        for b in range(n):
            v, lg = orc.to_small_vec_or_zero(T[b * N:(b + 1) * N]); o[b] = v; union |= set(int(x) for x in lg)
        outs.append(o)
    lp = np.array(sorted(union), dtype=np.uint64)
    for o in outs:
        for p in lp:
            o[:, int(p)] = 0
    return [o.reshape(-1) for o in outs], lp


@pytest.mark.parametrize("n,left,right,n_big", [(2, 4, 2, 0), (4, 8, 4, 3), (8, 32, 16, 40), (32, 256, 128, 7)])
def test_small_layers_round0_cvals(ctx, orc, n, left, right, n_big):
    import spartan2_b200 as sp
    rng = np.random.default_rng(n * 1000 + n_big)
    N = left * right; ell_b = n.bit_length() - 1
    A, B, Cm = _layers(orc, rng, n, N, n_big), _layers(orc, rng, n, N, n_big), _layers(orc, rng, n, N, n_big // 2)
    (A64, B64, C64), lp = _union_small(orc, (A, B, Cm), n, N)
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cm)
    (dA64, dB64, dC64), dpos, nl = sp.SmallValue.to_small_layers(ctx, [dA, dB, dC], n, N)
    assert nl == len(lp)
    assert np.array_equal(dpos.download((max(nl, 1),), dtype=np.uint64)[:nl], lp)
    for d, want in ((dA64, A64), (dB64, B64), (dC64, C64)):
        assert np.array_equal(d.download((n * N,), dtype=np.int64), want)
    E = orc.pow_split_evals(rand_fe(rng, 1), left, right); rhos = rand_fe(rng, ell_b)
    dE = ctx.upload(E)
    got = sp.SmallValue.nifs_round0(ctx, rhos, left, right, dE, dA64, dB64, dA, dB, dpos, nl, N, n)
    want = orc.nifs_round0_small(rhos, left, right, E, A, B, A64, B64, lp, N, n)
    assert np.array_equal(got, want)
    assert np.array_equal(got, orc.nifs_round(0, rhos, left, right, E, A, B, Cm, N, n))          # == the field path
    assert np.array_equal(sp.SmallValue.cvals(ctx, left, right, dE, dC, dC64, dpos, nl, N, n), orc.nifs_cvals_small(left, right, E, Cm, C64, lp, N, n))


def test_to_small_thresholds_and_all_large(ctx, orc):
    import spartan2_b200 as sp
    vals = [0, 1, 5, SMALL_MAX, SMALL_MAX + 1, Q - 1, Q - SMALL_MAX, Q - SMALL_MAX - 1, 1 << 64, Q // 2, Q - 7, 0, 0, 0, 0, 0]
    T = orc.to_mont(vals)
    (d64,), dpos, nl = sp.SmallValue.to_small_layers(ctx, [ctx.upload(T)], 1, 16)
    want, lg = orc.to_small_vec_or_zero(T)
    assert np.array_equal(d64.download((16,), dtype=np.int64), want) and nl == len(lg) == 4
    assert np.array_equal(dpos.download((16,), dtype=np.uint64)[:nl], lg)
    # every position large: the i64 layers are all zero and round 0 is the field correction alone
    rng = np.random.default_rng(2)
    n, left, right = 2, 4, 2; N = 8
    A, B = rand_fe(rng, n * N), rand_fe(rng, n * N)
    dA, dB = ctx.upload(A), ctx.upload(B)
    (dA64, dB64), dpos, nl = sp.SmallValue.to_small_layers(ctx, [dA, dB], n, N)
    assert nl == N and not dA64.download((n * N,), dtype=np.int64).any()
    E = orc.pow_split_evals(rand_fe(rng, 1), left, right); rhos = rand_fe(rng, 1)
    got = sp.SmallValue.nifs_round0(ctx, rhos, left, right, ctx.upload(E), dA64, dB64, dA, dB, dpos, nl, N, n)
    assert np.array_equal(got, orc.nifs_round(0, rhos, left, right, E, A, B, A, N, n))
