"""MSM / Hyrax commitment kernels vs the oracle (reference src/provider/msm.rs:878-934 differential tests:
Pippenger vs naive, msm_small vs msm for bit-widths {1,4,8,10,16,20,32,40,64}), through the C ABI.
Results are compared as affine points, the form the reference emits and absorbs into the transcript."""
import numpy as np
import pytest

from tests.curve_util import ORDER, points
from tests.gpu_util import ctx, rand_fe  # noqa: F401

pytestmark = pytest.mark.gpu
W = 64   # commitment width used by the tests (the reference's verifier-circuit keys are 16/32 wide; prover key 2048)


@pytest.fixture(scope="module")
def key(ctx, orc):
    import spartan2_b200 as sp
    pts = points(orc, W + 3, seed=11)
    ck, h, ck_s, h_s = pts[:W], pts[W:W + 1], pts[W + 1:W + 2], pts[W + 2:W + 3]
    return sp.CommitmentKey(ctx, ck, h, ck_s, h_s), ck, h


def test_msm_vs_oracle_and_naive(ctx, orc, key):
    import spartan2_b200 as sp
    dk, ck, h = key
    rng = np.random.default_rng(5)
    for n in (1, 2, 8, 33, W):
        s = rand_fe(rng, n)
        got = sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, s)
        assert np.array_equal(got, orc.msm(s, ck[:n]))
    # naive sum of scalar multiples (msm.rs:878-900), n = 8
    s = rand_fe(rng, 8)
    acc = np.zeros((1, 8), dtype=np.uint64)
    for i in range(8):
        acc = orc.point_add(acc, orc.scalar_mul(ck[i:i + 1], s[i:i + 1]))
    assert np.array_equal(sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, s), acc)


def test_msm_edge_scalars(ctx, orc, key):
    import spartan2_b200 as sp
    dk, ck, h = key
    vals = [0, 1, 2, 127, 128, 129, 255, 256, 0x8080808080808080, ORDER - 1, ORDER - 2, (ORDER - 1) // 2, 2**255 % ORDER, 2**248 + 129]
    s = orc.to_mont(vals)
    got = sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, s)
    assert np.array_equal(got, orc.msm(s, ck[:len(vals)]))
    z = np.zeros((5, 4), dtype=np.uint64)        # all-zero scalars -> identity (encoded as zeros)
    assert not sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, z).any()


@pytest.mark.parametrize("bits", [1, 4, 8, 10, 16, 20, 32, 40, 64])
def test_small_scalars_match_msm_small(ctx, orc, key, bits):
    import spartan2_b200 as sp
    dk, ck, h = key
    rng = np.random.default_rng(bits)
    raw = rng.integers(0, 2**64, size=W, dtype=np.uint64)
    if bits < 64:
        raw &= np.uint64((1 << bits) - 1)
    s = orc.to_mont([int(x) for x in raw])
    got = sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, s)
    assert np.array_equal(got, orc.msm_small(raw, ck))
    assert np.array_equal(got, orc.msm(s, ck))


def test_repeated_bases_hit_the_doubling_branch(ctx, orc):
    import spartan2_b200 as sp
    pts = points(orc, 4, seed=3)
    ck = np.concatenate([pts[:1]] * 8)        # the same base 8 times: equal points meet inside a bucket
    dk = sp.CommitmentKey(ctx, ck, pts[1:2], pts[2:3], pts[3:4])
    s = orc.to_mont([5, 5, 5, 5, 7, 7, ORDER - 5, 3])
    assert np.array_equal(sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, s), orc.msm(s, ck))


@pytest.mark.parametrize("n,small", [(W * 5, False), (W * 3 + 7, False), (W * 6, True), (0, False), (10, True)])
def test_hyrax_commit_rows(ctx, orc, key, n, small):
    import spartan2_b200 as sp
    dk, ck, h = key
    rng = np.random.default_rng(n + 1)
    rows = max(1, (n + W - 1) // W)
    if small:
        v = orc.to_mont([int(x) for x in rng.integers(0, 2, size=n)])
    else:
        v = rand_fe(rng, n)
        if n >= 2 * W:
            v[W:2 * W] = 0                        # an all-zero row: commit_zeros path (hyrax_pc.rs:305-319)
    blinds = rand_fe(rng, rows)
    got = sp.HyraxPCS.commit(ctx, dk, v, blinds, is_small=small)
    if n:
        want = orc.hyrax_commit(ck, h, v, blinds, is_small=small)
    else:
        want = orc.scalar_mul(h, blinds[:1])
    assert np.array_equal(got, want)


def test_hyrax_bind(ctx, orc):
    import spartan2_b200 as sp
    rng = np.random.default_rng(9)
    for rows, r_len in ((4, 64), (37, 128), (64, 2048)):
        poly, L = rand_fe(rng, rows * r_len), rand_fe(rng, rows)
        assert np.array_equal(sp.HyraxPCS.bind_with_delayed(ctx, poly, L, r_len), orc.hyrax_bind(poly, L, r_len))


def test_too_many_scalars_is_an_error(ctx, key):
    import spartan2_b200 as sp
    dk, ck, h = key
    with pytest.raises(sp.SpartanError) as ei:
        sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, np.zeros((W + 1, 4), dtype=np.uint64))
    assert ei.value.kind == "InvalidCommitmentKeyLength"


def test_tree_exceptional_cases_duplicate_and_opposite_bases(ctx, orc):
    """The reduction trees run lane-quad additions (curve_quad.cuh); their exceptional cases are resolved after the common path.
    A key with ck[1] = ck[0] and ck[3] = -ck[2] puts P + P and P + (-P) into the trees: with unit scalars the two equal (opposite)
    table entries sit alone in neighbouring lanes and meet in the last quad level; with equal random scalars every window pair meets
    somewhere (lane-serial mixed adds, quad levels, final kernel).  Compared with the oracle's MSM over the same bases."""
    import spartan2_b200 as sp
    pts = points(orc, W + 3, seed=13).copy()
    P_MOD = orc.MODS[orc.FP]
    pts[1] = pts[0]
    y = int(orc.from_mont(pts[2:3, 4:8], orc.FP)[0])
    pts[3, :4] = pts[2, :4]; pts[3, 4:8] = orc.to_mont([(P_MOD - y) % P_MOD], orc.FP)[0]
    ck, h, ck_s, h_s = pts[:W], pts[W:W + 1], pts[W + 1:W + 2], pts[W + 2:W + 3]
    dk = sp.CommitmentKey(ctx, ck, h, ck_s, h_s)
    rng = np.random.default_rng(17)
    r = int.from_bytes(rng.bytes(32), "little") % ORDER
    for vals in ([1, 1], [1, 1, 1, 1], [0, 0, 1, 1], [r, r], [r, r, r, r], [5, 5, 7, 7, 11], [r, r, 0, 0, 3]):
        s = orc.to_mont(vals)
        got = sp.DlogGroupExt.vartime_multiscalar_mul(ctx, dk, s)
        assert np.array_equal(got, orc.msm(s, ck[:len(vals)])), vals
    dk.free()
