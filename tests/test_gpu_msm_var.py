"""The DlogGroupExt surface for ARBITRARY bases (src/provider/traits.rs:118-162) on the device — signed-digit Pippenger
(sp2_msm_var), u64 scalars (sp2_msm_small_var), batches over one base slice (sp2_msm_batch_var) and the shared-weight
multi-MSM behind fold_commitments (sp2_msm_shared_weights) — against the oracle's restatements of src/provider/msm.rs
(orc.msm :59-222, orc.msm_small :367-620, orc.fold_commitments).  Bases are NOT the commitment key's."""
import numpy as np
import pytest

from tests.curve_util import points
from tests.gpu_util import ctx, rand_fe  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 8, 33, 300, 2048, 5000])
def test_msm_var_matches_oracle(ctx, orc, n):
    import spartan2_b200 as sp
    rng = np.random.default_rng(n)
    bases = points(orc, n, seed=100 + n % 7)
    s = rand_fe(rng, n)
    # the edge scalars the reference treats specially: 0 (skipped), 1 (boolean sum), p - 1, small values
    for k, val in enumerate([0, 1, orc.P_T256_SCALAR - 1, 2, 128, 129, 255, 256, (1 << 255)][:n]):
        s[k] = orc.to_mont([val])[0]
    orc.set_threads(orc.max_threads())
    assert np.array_equal(sp.DlogGroupExt.vartime_multiscalar_mul_var(ctx, s, bases), orc.msm(s, bases))
    orc.set_threads(1)


def test_msm_var_degenerate_inputs(ctx, orc):
    import spartan2_b200 as sp
    bases = points(orc, 4, seed=3)
    z = np.zeros((4, 4), dtype=np.uint64)
    assert not sp.DlogGroupExt.vartime_multiscalar_mul_var(ctx, z, bases).any()              # all-zero scalars -> identity
    assert not sp.DlogGroupExt.vartime_multiscalar_mul_var(ctx, z[:0], bases[:0]).any()      # empty
    # P + (-P): s and p - s on the same base cancel
    s = rand_fe(np.random.default_rng(1), 1)
    neg = orc.f_sub(np.zeros((1, 4), dtype=np.uint64), s)
    same = np.concatenate([bases[:1], bases[:1]])
    assert not sp.DlogGroupExt.vartime_multiscalar_mul_var(ctx, np.concatenate([s, neg]), same).any()
    # doubling inside a bucket: the same base twice with the same scalar
    two = sp.DlogGroupExt.vartime_multiscalar_mul_var(ctx, np.concatenate([s, s]), same)
    assert np.array_equal(two, orc.scalar_mul(bases[:1], orc.f_add(s, s)))
    with pytest.raises(sp.SpartanError):
        sp.DlogGroupExt.vartime_multiscalar_mul_var(ctx, z, bases[:3])


@pytest.mark.parametrize("bits", [1, 2, 7, 8, 10, 16, 33, 63, 64])
def test_msm_small_var_matches_msm_small(ctx, orc, bits):
    import spartan2_b200 as sp
    rng = np.random.default_rng(bits); n = 700
    bases = points(orc, n, seed=11)
    hi = (1 << bits) - 1
    s = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
    s[0] = hi; s[1] = 0
    assert np.array_equal(sp.DlogGroupExt.vartime_multiscalar_mul_small(ctx, s, bases), orc.msm_small(s, bases))


def test_batch_and_shared_weights(ctx, orc):
    import spartan2_b200 as sp
    rng = np.random.default_rng(9)
    bases = points(orc, 96, seed=12)
    vecs = [rand_fe(rng, k) for k in (96, 1, 40, 0, 7)]
    got = sp.DlogGroupExt.batch_vartime_multiscalar_mul(ctx, vecs, bases)
    for g, v in zip(got, vecs):
        want = orc.msm(v, bases[:len(v)]) if len(v) else np.zeros((1, 8), dtype=np.uint64)
        assert np.array_equal(g, want[0])
    # shared weights: rows x n bases, one weight vector (fold_commitments' shape: n = #instances, rows = commitment rows)
    n, rows = 32, 13
    w = rand_fe(rng, n); w[3] = orc.to_mont([1])[0]; w[4] = 0
    rows_b = points(orc, n * rows, seed=13).reshape(rows, n, 8)
    got = sp.DlogGroupExt.vartime_multiscalar_mul_shared_weights(ctx, w, rows_b)
    comms = np.ascontiguousarray(rows_b.transpose(1, 0, 2)).reshape(n * rows, 8)          # fold_commitments takes [instance][row]
    assert np.array_equal(got, orc.fold_commitments(comms, n, rows, w))
