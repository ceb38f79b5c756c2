"""SpartanSNARK prep_prove + prove on the device vs the oracle's restatement of src/spartan.rs:176-466, with the
same prover randomness: every proof field must be bit-identical, and the oracle's verifier (src/spartan.rs:469-578)
must accept the device-made proof — the reference's own integration test pattern (spartan.rs:653-688)."""
import numpy as np
import pytest

from tests.curve_util import points
from tests.gpu_util import ctx, rand_fe  # noqa: F401
from tests.r1cs_util import dims, random_r1cs

pytestmark = pytest.mark.gpu


def _run(ctx, orc, seed, lc, lv, width, num_public, rest_frac):
    import spartan2_b200 as sp
    inst = random_r1cs(seed, lc, lv, num_public=num_public, rest_frac=rest_frac, width=width)
    rng = np.random.default_rng(seed + 100)
    pts = points(orc, width + 3, seed=21)
    ck, h, ck_s, h_s = pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3]
    nv = inst["num_vars"]; rows = nv // width
    cached_len = inst["num_shared"] + inst["num_precommitted"]; cached_rows = cached_len // width
    blinds = rand_fe(rng, rows); blind_eval = rand_fe(rng, 1); d_vec = rand_fe(rng, width); r_delta = rand_fe(rng, 1); r_beta = rand_fe(rng, 1)
    vk = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
    W, X = inst["W"], inst["X"]
    # --- device
    S = sp.SplitR1CSShape(ctx, *dims(inst), inst["A"], inst["B"], inst["C"])
    K = sp.CommitmentKey(ctx, ck, h, ck_s, h_s)
    prep = sp.SpartanSNARK.prep_prove(ctx, S, K, W[:cached_len], blinds[:cached_rows], is_small=False)
    proof = sp.SpartanSNARK.prove(ctx, S, K, prep, vk, X, W[cached_len:], blinds, blind_eval, d_vec, r_delta, r_beta)
    # --- oracle
    O = orc.Shape(*dims(inst), inst["A"], inst["B"], inst["C"])
    keys = orc.Keys(ck, h, ck_s, h_s)
    comm_pre = orc.hyrax_commit(ck, h, W[:cached_len], blinds[:cached_rows], is_small=False) if cached_rows else np.zeros((0, 8), dtype=np.uint64)
    assert np.array_equal(prep.comm, comm_pre)
    oproof = orc.spartan_prove(O, keys, vk, X, W, comm_pre, orc.Rand(blinds, blind_eval, d_vec, r_delta, r_beta))
    for f in sp.SpartanProof.FIELDS:
        assert np.array_equal(getattr(proof, f).reshape(-1), getattr(oproof, f).reshape(-1)), f
    # the oracle's verifier accepts the device-made proof
    vp = orc.Proof(proof.l, proof.nry, proof.rows, proof.num_cols)
    for f in sp.SpartanProof.FIELDS:
        getattr(vp, f)[...] = getattr(proof, f).reshape(getattr(vp, f).shape)
    assert orc.spartan_verify(O, keys, vk, X, vp) == 0
    # ... and rejects a tampered one
    vp.eval_W[0, 0] ^= np.uint64(1)
    assert orc.spartan_verify(O, keys, vk, X, vp) != 0
    # prove is repeatable on the same prep state (the reference returns the prep state for reuse, snark.rs:39-47)
    proof2 = sp.SpartanSNARK.prove(ctx, S, K, prep, vk, X, W[cached_len:], blinds, blind_eval, d_vec, r_delta, r_beta)
    for f in sp.SpartanProof.FIELDS:
        assert np.array_equal(getattr(proof2, f), getattr(proof, f)), f
    return proof


@pytest.mark.parametrize("seed,lc,lv,width,npub,rest", [
    (1, 6, 6, 16, 2, 0.0),       # tiny: single-launch sum-checks
    (2, 8, 7, 128, 3, 0.0),      # one commitment row (rows == 1 special case of HyraxPCS::prove, hyrax_pc.rs:420-424)
    (3, 10, 10, 64, 5, 0.5),     # rest section present
    (4, 13, 13, 64, 30, 0.25),   # multi-CTA rounds + tail kernels, many public inputs
    (5, 12, 14, 256, 2, 0.0),    # more variables than constraints
    (6, 14, 12, 64, 2, 0.0),     # more constraints than variables
    (7, 18, 17, 256, 3, 0.25),   # streaming rounds of the persistent kernel (t(1) derived from the claim), pipelined mid rounds, early comm_LZ
])
def test_prove_bit_exact_and_verifies(ctx, orc, seed, lc, lv, width, npub, rest):
    _run(ctx, orc, seed, lc, lv, width, npub, rest)


def test_sha256_circuit_prove(ctx, orc):
    """The benchmark circuit itself (benches/sha256_spartan.rs:37-152) on a 64-byte message: 2 compressions,
    N = M = 2^16, Hyrax width 2048, witness all bits (is_small = true, as the bench passes), key from the device
    point generator (checked to be on the curve by the oracle)."""
    import hashlib
    import spartan2_b200 as sp
    from spartan2_b200.frontend import Sha256Circuit
    msg = bytes(range(64))
    circ = Sha256Circuit(msg)
    assert circ.digest == hashlib.sha256(msg).digest() and circ.is_satisfied()
    width = 2048
    pts = ctx.test_points(width + 3, seed=5)
    assert all(orc.on_curve(pts[i:i + 1]) for i in (0, 1, 77, width + 2))
    ck, h, ck_s, h_s = pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3]
    A, B, Cm = circ.matrices()
    W, X = circ.witness()
    nv = circ.num_vars; rows = nv // width; cached_len = circ.num_precommitted; cached_rows = cached_len // width
    rng = np.random.default_rng(64)
    blinds = rand_fe(rng, rows); blind_eval = rand_fe(rng, 1); d_vec = rand_fe(rng, width); r_delta = rand_fe(rng, 1); r_beta = rand_fe(rng, 1)
    vk = bytes(32)
    S = sp.SplitR1CSShape(ctx, *circ.dims(), A, B, Cm)
    K = sp.CommitmentKey(ctx, ck, h, ck_s, h_s)
    prep = sp.SpartanSNARK.prep_prove(ctx, S, K, W[:cached_len], blinds[:cached_rows], is_small=True)
    proof = sp.SpartanSNARK.prove(ctx, S, K, prep, vk, X, W[cached_len:], blinds, blind_eval, d_vec, r_delta, r_beta)
    O = orc.Shape(*circ.dims(), A, B, Cm)
    keys = orc.Keys(ck, h, ck_s, h_s)
    comm_pre = orc.hyrax_commit(ck, h, W[:cached_len], blinds[:cached_rows], is_small=True)
    assert np.array_equal(prep.comm, comm_pre)
    oproof = orc.spartan_prove(O, keys, vk, X, W, comm_pre, orc.Rand(blinds, blind_eval, d_vec, r_delta, r_beta))
    for f in sp.SpartanProof.FIELDS:
        assert np.array_equal(getattr(proof, f).reshape(-1), getattr(oproof, f).reshape(-1)), f
    vp = orc.Proof(proof.l, proof.nry, proof.rows, proof.num_cols)
    for f in sp.SpartanProof.FIELDS:
        getattr(vp, f)[...] = getattr(proof, f).reshape(getattr(vp, f).shape)
    assert orc.spartan_verify(O, keys, vk, X, vp) == 0
    sz = S.sizes()
    assert sz["num_cons"] == 1 << 16 and sz["long_rows"] > 0 and sz["long_cols"] > 0


@pytest.mark.parametrize("msg_len", [2048])
def test_full_size_config_proof_is_bit_exact_and_accepted_by_the_oracle_verifier(ctx, orc, msg_len):
    """BASELINE config 2 at full size (2 KiB message, N = M = 2^20): EVERY field of the device-made proof equals the
    oracle prover's (same keys, witness and prover randomness; restatement of src/spartan.rs:219-466), the oracle's
    restatement of SpartanSNARK::verify (src/spartan.rs:469-578) accepts it, and rejects it after tampering."""
    import hashlib
    import spartan2_b200 as sp
    from spartan2_b200.frontend import Sha256Circuit
    msg = b"\x00" * msg_len
    circ = Sha256Circuit(msg)
    assert circ.digest == hashlib.sha256(msg).digest()
    width = 2048
    pts = ctx.test_points(width + 3, seed=5)
    ck, h, ck_s, h_s = pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3]
    A, B, Cm = circ.matrices(); W, X = circ.witness()
    nv = circ.num_vars; rows = nv // width; cl = circ.num_precommitted; cr = cl // width
    rng = np.random.default_rng(msg_len)
    blinds = rand_fe(rng, rows); be = rand_fe(rng, 1); dv = rand_fe(rng, width); rd = rand_fe(rng, 1); rb = rand_fe(rng, 1)
    vk = bytes(range(32))
    S = sp.SplitR1CSShape(ctx, *circ.dims(), A, B, Cm)
    K = sp.CommitmentKey(ctx, ck, h, ck_s, h_s)
    prep = sp.SpartanSNARK.prep_prove(ctx, S, K, W[:cl], blinds[:cr], is_small=True)
    proof = sp.SpartanSNARK.prove(ctx, S, K, prep, vk, X, None, blinds, be, dv, rd, rb)
    assert proof.l == 20 and proof.nry == 21 and proof.rows == 512
    O = orc.Shape(*circ.dims(), A, B, Cm)
    keys = orc.Keys(ck, h, ck_s, h_s)
    orc.set_threads(orc.max_threads())
    comm_pre = orc.hyrax_commit(ck, h, W[:cl], blinds[:cr], is_small=True)
    assert np.array_equal(prep.comm, comm_pre)
    oproof = orc.spartan_prove(O, keys, vk, X, W, comm_pre, orc.Rand(blinds, be, dv, rd, rb))
    for f in sp.SpartanProof.FIELDS:
        assert np.array_equal(getattr(proof, f).reshape(-1), getattr(oproof, f).reshape(-1)), "full-size parity: " + f
    vp = orc.Proof(proof.l, proof.nry, proof.rows, proof.num_cols)
    for f in sp.SpartanProof.FIELDS:
        getattr(vp, f)[...] = getattr(proof, f).reshape(getattr(vp, f).shape)
    assert orc.spartan_verify(O, keys, vk, X, vp) == 0
    vp.z_vec[7, 1] ^= np.uint64(4)
    assert orc.spartan_verify(O, keys, vk, X, vp) != 0
    orc.set_threads(1)


@pytest.mark.parametrize("env", [{"SP2_NO_GATES": "1"}, {"SP2_NO_DERIVE": "1"}, {"SP2_NO_GATES": "1", "SP2_NO_DERIVE": "1"}])
def test_measurement_switches_keep_the_proof(env):
    """SP2_NO_GATES / SP2_NO_DERIVE are read once per process: a fresh interpreter proves the 2^18 / 2^17 instance (streaming rounds
    of the persistent kernel: t(1) derived from the claim by default, gate kernels opened by the helper thread) with the switches set,
    and compares every field with the oracle's proof — the default path does the same in test_prove_bit_exact_and_verifies."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "import tests.conftest\n"
        "import spartan2_b200 as sp\n"
        "from oracle import pyoracle as orc\n"
        "from tests.test_gpu_spartan import _run\n"
        "c = sp.Context(0)\n"
        "_run(c, orc, 11, 18, 17, 256, 3, 0.25)\n"
        "c.close()\n"
        "print('switches ok')\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "switches ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
