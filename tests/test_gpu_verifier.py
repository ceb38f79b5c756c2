"""The on-box verifier (sp2_spartan_verify: matrix MLE evaluations, Hyrax row MSM and IPA checks on the device)
against the oracle's restatement of SpartanSNARK::verify (src/spartan.rs:469-578): both accept the device-made proof, and both
reject it — the device with ProofVerifyError — when ANY field is tampered; a proof made by the ORACLE prover is accepted by
the device verifier too (so prover and verifier are checked independently of each other)."""
import numpy as np
import pytest

from tests.curve_util import points
from tests.gpu_util import ctx, rand_fe  # noqa: F401
from tests.r1cs_util import dims, random_r1cs

pytestmark = pytest.mark.gpu


def _setup(ctx, orc, seed, lc, lv, width, npub, rest):
    import spartan2_b200 as sp
    inst = random_r1cs(seed, lc, lv, num_public=npub, rest_frac=rest, width=width)
    rng = np.random.default_rng(seed + 7)
    pts = points(orc, width + 3, seed=21)
    ck, h, ck_s, h_s = pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3]
    rows = inst["num_vars"] // width; cl = inst["num_shared"] + inst["num_precommitted"]; cr = cl // width
    rnd = (rand_fe(rng, rows), rand_fe(rng, 1), rand_fe(rng, width), rand_fe(rng, 1), rand_fe(rng, 1))
    vk = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
    S = sp.SplitR1CSShape(ctx, *dims(inst), inst["A"], inst["B"], inst["C"])
    K = sp.CommitmentKey(ctx, ck, h, ck_s, h_s)
    O = orc.Shape(*dims(inst), inst["A"], inst["B"], inst["C"]); keys = orc.Keys(ck, h, ck_s, h_s)
    W, X = inst["W"], inst["X"]
    prep = sp.SpartanSNARK.prep_prove(ctx, S, K, W[:cl], rnd[0][:cr], is_small=False)
    proof = sp.SpartanSNARK.prove(ctx, S, K, prep, vk, X, W[cl:], *rnd)
    comm_pre = orc.hyrax_commit(ck, h, W[:cl], rnd[0][:cr], is_small=False) if cr else np.zeros((0, 8), dtype=np.uint64)
    oproof = orc.spartan_prove(O, keys, vk, X, W, comm_pre, orc.Rand(*rnd))
    return sp, S, K, O, keys, vk, X, proof, oproof


@pytest.mark.parametrize("seed,lc,lv,width,npub,rest", [(1, 6, 6, 16, 2, 0.0), (2, 8, 7, 128, 3, 0.0), (3, 10, 10, 64, 5, 0.5), (4, 13, 13, 64, 30, 0.25)])
def test_device_verifier_accepts_and_rejects(ctx, orc, seed, lc, lv, width, npub, rest):
    sp, S, K, O, keys, vk, X, proof, oproof = _setup(ctx, orc, seed, lc, lv, width, npub, rest)
    sp.SpartanSNARK.verify(ctx, S, K, vk, X, proof)                               # accept
    # a proof made by the ORACLE prover, carried into the product's proof struct, is accepted as well
    carried = sp.SpartanProof(proof.l, proof.nry, proof.rows, proof.num_cols)
    for f in sp.SpartanProof.FIELDS:
        getattr(carried, f)[...] = getattr(oproof, f).reshape(getattr(carried, f).shape)
    sp.SpartanSNARK.verify(ctx, S, K, vk, X, carried)
    # every field, tampered: device rejects with ProofVerifyError, and so does the oracle verifier
    vp = orc.Proof(proof.l, proof.nry, proof.rows, proof.num_cols)
    for f in sp.SpartanProof.FIELDS:
        a = getattr(proof, f); idx = (a.shape[0] // 2, 1)
        old = a[idx]; a[idx] = old ^ np.uint64(2)
        with pytest.raises(sp.SpartanError) as ei:
            sp.SpartanSNARK.verify(ctx, S, K, vk, X, proof)
        assert ei.value.kind == "ProofVerifyError", f
        for g in sp.SpartanProof.FIELDS:
            getattr(vp, g)[...] = getattr(proof, g).reshape(getattr(vp, g).shape)
        assert orc.spartan_verify(O, keys, vk, X, vp) != 0, f
        a[idx] = old
    sp.SpartanSNARK.verify(ctx, S, K, vk, X, proof)
    # wrong public input / wrong vk digest
    if len(X):
        X2 = X.copy(); X2[0, 0] ^= np.uint64(1)
        with pytest.raises(sp.SpartanError):
            sp.SpartanSNARK.verify(ctx, S, K, vk, X2, proof)
    with pytest.raises(sp.SpartanError):
        sp.SpartanSNARK.verify(ctx, S, K, bytes(32), X, proof)


def test_device_verifier_on_the_sha256_circuit(ctx, orc):
    """the benchmark circuit on a 64-byte message (N = M = 2^16, 32 commitment rows of 2048): device prove -> device verify"""
    import spartan2_b200 as sp
    from spartan2_b200.frontend import Sha256Circuit
    circ = Sha256Circuit(bytes(range(64)))
    width = 2048
    pts = ctx.test_points(width + 3, seed=5)
    A, B, Cm = circ.matrices(); W, X = circ.witness()
    rows = circ.num_vars // width; cl = circ.num_precommitted; cr = cl // width
    rng = np.random.default_rng(64)
    rnd = (rand_fe(rng, rows), rand_fe(rng, 1), rand_fe(rng, width), rand_fe(rng, 1), rand_fe(rng, 1))
    S = sp.SplitR1CSShape(ctx, *circ.dims(), A, B, Cm)
    K = sp.CommitmentKey(ctx, pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3])
    prep = sp.SpartanSNARK.prep_prove(ctx, S, K, W[:cl], rnd[0][:cr], is_small=True)
    proof = sp.SpartanSNARK.prove(ctx, S, K, prep, bytes(32), X, None, *rnd)
    sp.SpartanSNARK.verify(ctx, S, K, bytes(32), X, proof)
    proof.z_vec[100, 0] ^= np.uint64(1)
    with pytest.raises(sp.SpartanError) as ei:
        sp.SpartanSNARK.verify(ctx, S, K, bytes(32), X, proof)
    assert ei.value.kind == "ProofVerifyError"
