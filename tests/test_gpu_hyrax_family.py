"""The rest of the Hyrax commitment family and the zero-check first round on the device vs the oracle:
HyraxPCS::commit_without_blind / commit_incremental (hyrax_pc.rs:533-607), rerandomize_commitment (:321-344),
fold_blinds / fold_commitments_partial (:795-874), and EqSumCheckInstance::evaluation_points_zero_check_round0
(src/sumcheck.rs:1163-1271) — SURVEY.md §8 rows a5, a25, a27."""
import numpy as np
import pytest

from tests.curve_util import points
from tests.gpu_util import ctx, rand_fe  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def key(ctx, orc):
    import spartan2_b200 as sp
    width = 64
    pts = points(orc, width + 3, seed=44)
    ck = sp.CommitmentKey(ctx, pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3])
    yield ck, pts, width
    ck.free()


def test_commit_without_blind_and_incremental(ctx, orc, key):
    import spartan2_b200 as sp
    ck, pts, width = key
    rng = np.random.default_rng(2)
    v = rand_fe(rng, 5 * width + 9); v[2 * width:3 * width] = 0                    # a zero row, ragged last row
    raw = sp.HyraxPCS.commit_without_blind(ctx, ck, v)
    assert np.array_equal(raw, orc.hyrax_commit_without_blind(pts[:width], v)) and not raw[2].any()
    bits = orc.to_mont([int(b) for b in rng.integers(0, 2, size=3 * width)])
    assert np.array_equal(sp.HyraxPCS.commit_without_blind(ctx, ck, bits, is_small=True), orc.hyrax_commit_without_blind(pts[:width], bits, is_small=True))
    blinds = rand_fe(rng, 6)
    delta = np.zeros_like(v); delta[1] = rand_fe(rng, 1); delta[2 * width + 5] = rand_fe(rng, 1); delta[-1] = rand_fe(rng, 1)
    got = sp.HyraxPCS.commit_incremental(ctx, ck, raw, delta, blinds)
    assert np.array_equal(got, orc.hyrax_commit_incremental(pts[:width], pts[width:width + 1], raw, delta, blinds))
    assert np.array_equal(got, orc.hyrax_commit(pts[:width], pts[width:width + 1], orc.f_add(v, delta), blinds))   # = commit(v + delta)
    # fewer raw rows than delta rows: the missing ones are the identity (hyrax_pc.rs:590)
    got2 = sp.HyraxPCS.commit_incremental(ctx, ck, raw[:2], delta, blinds)
    assert np.array_equal(got2, orc.hyrax_commit_incremental(pts[:width], pts[width:width + 1], raw[:2], delta, blinds))


def test_rerandomize_commitment(ctx, orc, key):
    import spartan2_b200 as sp
    ck, pts, width = key
    rng = np.random.default_rng(3)
    v = rand_fe(rng, 4 * width); r_old = rand_fe(rng, 4); r_new = rand_fe(rng, 4); r_new[2] = r_old[2]     # one unchanged blind
    comm = orc.hyrax_commit(pts[:width], pts[width:width + 1], v, r_old)
    got = sp.HyraxPCS.rerandomize_commitment(ctx, ck, comm, r_old, r_new)
    assert np.array_equal(got, orc.hyrax_rerandomize(pts[width:width + 1], comm, r_old, r_new))
    assert np.array_equal(got, orc.hyrax_commit(pts[:width], pts[width:width + 1], v, r_new))
    with pytest.raises(sp.SpartanError) as ei:
        sp.HyraxPCS.rerandomize_commitment(ctx, ck, comm, r_old[:3], r_new)
    assert ei.value.kind == "InvalidInputLength"


@pytest.mark.parametrize("n,rows,data_rows", [(2, 4, 2), (8, 5, 3), (32, 16, 13), (4, 3, 3), (4, 3, 0)])
def test_fold_blinds_and_commitments_partial(ctx, orc, key, n, rows, data_rows):
    import spartan2_b200 as sp
    ck, pts, width = key
    rng = np.random.default_rng(n)
    h = pts[width:width + 1]
    Ws = rand_fe(rng, n * rows * width).reshape(n, rows * width, 4); Ws[:, data_rows * width:] = 0
    bl = rand_fe(rng, n * rows); w = rand_fe(rng, n); w[0] = orc.to_mont([1])[0]
    orc.set_threads(orc.max_threads())
    comms = np.concatenate([orc.hyrax_commit(pts[:width], h, Ws[i], bl[i * rows:(i + 1) * rows]) for i in range(n)])
    fb = sp.HyraxPCS.fold_blinds(ctx, bl, n, rows, w)
    assert np.array_equal(fb, orc.fold_blinds(bl, n, rows, w))
    got = sp.HyraxPCS.fold_commitments_partial(ctx, ck, comms, n, rows, w, data_rows, fb)
    assert np.array_equal(got, orc.fold_commitments_partial(comms, n, rows, w, data_rows, fb, h))
    assert np.array_equal(sp.fold_commitments(ctx, comms, n, rows, w), orc.fold_commitments(comms, n, rows, w))
    orc.set_threads(1)


@pytest.mark.parametrize("l", [1, 2, 3, 4, 7, 10, 13, 18])
def test_zero_check_round0(ctx, orc, l):
    import spartan2_b200 as sp
    rng = np.random.default_rng(500 + l); n = 1 << l
    A, B, taus = rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, l)
    got = sp.SumcheckProof.evaluation_points_zero_check_round0(ctx, taus, A, B)
    assert np.array_equal(got, orc.zero_check_round0(taus, A, B))
    if l <= 10:
        # ... which is the first round polynomial of the full cubic prover on a satisfied instance, at 0, 2, 3
        Cz = orc.f_mul(A, B)
        polys = orc.sumcheck_cubic_prove(np.zeros((1, 4), dtype=np.uint64), taus, A.copy(), B.copy(), Cz, orc.Transcript(b"x"))[0]
        c = orc.from_mont(np.asarray(polys).reshape(-1, 4)[:4]); P = orc.P_T256_SCALAR
        assert orc.from_mont(got) == [sum(ci * pow(x, k, P) for k, ci in enumerate(c)) % P for x in (0, 2, 3)]
    taus[0] = 0                                                                      # tau_0 = 0: the reference's fallback branch (:1244-1270)
    assert np.array_equal(sp.SumcheckProof.evaluation_points_zero_check_round0(ctx, taus, A, B), orc.zero_check_round0(taus, A, B))
