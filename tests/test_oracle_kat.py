"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the
path.  CPU only."""
import random

import numpy as np

from oracle import pyref


def test_keccak256_kat(orc):
    # reference src/provider/keccak.rs:155-163
    data = (0xffffffff).to_bytes(4, "little")
    want = "29045a592007d0c246ef02c2223570da9522d0cf0f73282c79a1bc8f0bb2c238"
    assert orc.keccak256(data).hex() == want
    assert pyref.keccak256(data).hex() == want


def test_keccak256_c_vs_python_lengths(orc):
    rng = random.Random(1)
    for n in [0, 1, 31, 135, 136, 137, 200, 271, 272, 273, 1000]:
        d = bytes(rng.getrandbits(8) for _ in range(n))
        assert orc.keccak256(d) == pyref.keccak256(d)


def test_transcript_kat_pallas(orc):
    # reference src/provider/keccak.rs:120-152 (PallasHyraxEngine)
    p = pyref.P_PALLAS_SCALAR
    ts = pyref.Transcript(b"test", p)
    ts.absorb_scalar(b"s1", 2); ts.absorb_scalar(b"s2", 5)
    c1 = ts.squeeze(b"c1")
    assert c1.to_bytes(32, "little").hex() == "b67339da79ce5f6dc72ad23c8c3b4179f49655cadf92d47e79c3e7788f00f125"
    ts.absorb_scalar(b"s3", 128)
    c2 = ts.squeeze(b"c2")
    assert c2.to_bytes(32, "little").hex() == "b7f033d47b3519dd6efe320b995eaad1dc11712cb9b655d2e7006ed5f86bd321"
    # the C oracle, instantiated on the same field, reproduces the same challenges
    t = orc.Transcript(b"test")
    t.absorb_scalars(b"s1", orc.to_mont([2], orc.FPALLAS), orc.FPALLAS)
    t.absorb_scalars(b"s2", orc.to_mont([5], orc.FPALLAS), orc.FPALLAS)
    assert orc.from_mont(t.squeeze(b"c1", orc.FPALLAS), orc.FPALLAS)[0] == c1
    t.absorb_scalars(b"s3", orc.to_mont([128], orc.FPALLAS), orc.FPALLAS)
    assert orc.from_mont(t.squeeze(b"c2", orc.FPALLAS), orc.FPALLAS)[0] == c2


def test_transcript_t256_c_vs_python(orc):
    # T256 has no KAT in the reference (SURVEY §8c "parity unpinned" item 3): C and Python
    # restatements must at least agree, incl. dom_sep, multi-absorb and round counter.
    p = pyref.P_T256_SCALAR
    rng = random.Random(7)
    ts = pyref.Transcript(b"SpartanSNARK", p); t = orc.Transcript(b"SpartanSNARK")
    for rnd in range(5):
        vals = [rng.randrange(p) for _ in range(rnd + 1)]
        ts.absorb_scalars(b"x", vals); t.absorb_scalars(b"x", orc.to_mont(vals))
        if rnd == 2:
            ts.dom_sep(b"inner product argument (linear)"); t.dom_sep(b"inner product argument (linear)")
        blob = bytes(rng.getrandbits(8) for _ in range(40 * rnd))
        ts.absorb_bytes(b"vk", blob); t.absorb_bytes(b"vk", blob)
        assert orc.from_mont(t.squeeze(b"c"))[0] == ts.squeeze(b"c")


def test_unipoly_kats(orc):
    # reference src/polys/univariate.rs:298-362 (quadratic 2x^2+3x+1, cubic x^3+2x^2+3x+1)
    c = orc.from_mont(orc.unipoly_from_evals(orc.to_mont([1, 6, 15])))
    assert c == [1, 3, 2]
    assert orc.from_mont(orc.unipoly_eval(orc.to_mont([1, 3, 2]), orc.to_mont([3])))[0] == 28
    c = orc.from_mont(orc.unipoly_from_evals(orc.to_mont([1, 7, 23, 55])))
    assert c == [1, 3, 2, 1]
    assert orc.from_mont(orc.unipoly_eval(orc.to_mont([1, 3, 2, 1]), orc.to_mont([4])))[0] == 109


def test_spmv_kat(orc):
    # reference src/r1cs/sparse.rs:637-653: [[0,2,7],[0,0,3],[4,0,0]] * [1,2,3] = [25,9,4]
    data = orc.to_mont([2, 7, 3, 4]); idx = np.array([1, 2, 2, 0], dtype=np.uint32); ptr = np.array([0, 2, 3, 4], dtype=np.uint32)
    z = orc.to_mont([1, 2, 3])
    assert orc.from_mont(orc.csr_multiply_vec(3, data, idx, ptr, z)) == [25, 9, 4]
    # same through the classified path (PrecomputedSparseMatrix): pad to the shape interface
    empty = (orc.fe_array(0), np.zeros(0, dtype=np.uint32), np.zeros(5, dtype=np.uint32))
    ptr4 = np.array([0, 2, 3, 4, 4], dtype=np.uint32)
    S = orc.Shape(4, 3, 0, 0, 2, 0, 0, (data, idx, ptr4), empty, empty)
    az, bz, cz = S.multiply_vec(z)
    assert orc.from_mont(az) == [25, 9, 4, 0] and orc.from_mont(bz) == [0, 0, 0, 0]


def test_eq_kat(orc):
    # reference src/polys/eq.rs:132-148: eq(r=[1,0,1]) is 1 at index 5 only
    ev = orc.from_mont(orc.eq_evals(orc.to_mont([1, 0, 1])))
    assert ev == [1 if i == 5 else 0 for i in range(8)]


def test_mle_bind_order(orc):
    # reference src/polys/multilinear.rs:247-379: Z = [0,0,0,1,0,1,0,2] is (x1+x2)*x3, first
    # challenge binds the top (MSB) variable; binding all variables == evaluate
    p = pyref.P_T256_SCALAR
    Z = [0, 0, 0, 1, 0, 1, 0, 2]
    r = [5, 7, 11]
    assert pyref.mle_eval(Z, r, p) == (5 + 7) * 11 % p
    Zm = orc.to_mont(Z)
    for rv in r:
        Zm = orc.bind_top(Zm, orc.to_mont([rv]))
    assert orc.from_mont(Zm)[0] == (5 + 7) * 11 % p
    rng = random.Random(3)
    Z = [rng.randrange(p) for _ in range(32)]; r = [rng.randrange(p) for _ in range(5)]
    Zm = orc.to_mont(Z)
    for rv in r:
        Zm = orc.bind_top(Zm, orc.to_mont([rv]))
    assert orc.from_mont(Zm)[0] == pyref.mle_eval(Z, r, p)
    assert orc.from_mont(orc.eq_evals(orc.to_mont(r))) == pyref.eq_evals(r, p)
