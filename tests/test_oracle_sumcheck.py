"""The oracle's restatement of the reference's optimised sum-check provers (split-eq + BDDT
claim derivation, sumcheck.rs:502-571/920-1429; quadratic, sumcheck.rs:128-247) against the
definition-based Python prover, and prove->verify self-consistency as in the reference's own
hot-path tests (sumcheck.rs:1431-1573)."""
import random

import numpy as np
import pytest

from oracle import pyref

P = pyref.P_T256_SCALAR


def _rand_tables(rng, n, small=False):
    if small:
        return [rng.randrange(4) for _ in range(n)]
    return [rng.randrange(P) for _ in range(n)]


@pytest.mark.parametrize("l", [1, 2, 3, 4, 5, 6, 7])
def test_cubic_vs_definition(orc, l):
    rng = random.Random(0xDEADBEEF + l)
    n = 1 << l
    A, B, C = (_rand_tables(rng, n) for _ in range(3))
    taus = [rng.randrange(P) for _ in range(l)]
    claim = sum(e * (a * b - c) for e, a, b, c in zip(pyref.eq_evals(taus, P), A, B, C)) % P
    tp = pyref.Transcript(b"test", P)
    polys, rs, finals = pyref.sumcheck_cubic_naive(taus, A, B, C, tp, P)
    tc = orc.Transcript(b"test")
    cpolys, cr, cclaims, _ = orc.sumcheck_cubic_prove(orc.to_mont([claim]), orc.to_mont(taus), orc.to_mont(A), orc.to_mont(B), orc.to_mont(C), tc)
    assert [orc.from_mont(cp) for cp in cpolys] == polys
    assert orc.from_mont(cr) == rs
    assert orc.from_mont(cclaims) == list(finals)
    # transcripts end in the same state
    assert orc.from_mont(tc.squeeze(b"x"))[0] == tp.squeeze(b"x")


def test_cubic_satisfied_r1cs_claim_zero(orc):
    # Spartan's actual use: C = A*B pointwise, claim = 0, small witness-like values
    l = 6; rng = random.Random(4); n = 1 << l
    A = _rand_tables(rng, n, True); B = _rand_tables(rng, n, True); C = [a * b % P for a, b in zip(A, B)]
    taus = [rng.randrange(P) for _ in range(l)]
    tp = pyref.Transcript(b"t", P); tc = orc.Transcript(b"t")
    polys, rs, finals = pyref.sumcheck_cubic_naive(taus, A, B, C, tp, P)
    cpolys, cr, cclaims, _ = orc.sumcheck_cubic_prove(orc.to_mont([0]), orc.to_mont(taus), orc.to_mont(A), orc.to_mont(B), orc.to_mont(C), tc)
    assert [orc.from_mont(cp) for cp in cpolys] == polys and orc.from_mont(cclaims) == list(finals)


def test_cubic_fallback_tau_zero(orc):
    # tau_i = 0 makes l(1)*p = 0: derive_from_claim returns None and the third sum is used
    # (sumcheck.rs:1290-1292, 1327-1396).  Must still match the definition.
    l = 5; rng = random.Random(8); n = 1 << l
    A, B, C = (_rand_tables(rng, n) for _ in range(3))
    taus = [rng.randrange(P) for _ in range(l)]; taus[0] = 0; taus[3] = 0
    claim = sum(e * (a * b - c) for e, a, b, c in zip(pyref.eq_evals(taus, P), A, B, C)) % P
    tp = pyref.Transcript(b"t", P); tc = orc.Transcript(b"t")
    polys, rs, finals = pyref.sumcheck_cubic_naive(taus, A, B, C, tp, P)
    cpolys, cr, cclaims, _ = orc.sumcheck_cubic_prove(orc.to_mont([claim]), orc.to_mont(taus), orc.to_mont(A), orc.to_mont(B), orc.to_mont(C), tc)
    assert [orc.from_mont(cp) for cp in cpolys] == polys and orc.from_mont(cclaims) == list(finals)


@pytest.mark.parametrize("l", [1, 2, 5, 8])
def test_quad_vs_definition(orc, l):
    rng = random.Random(77 + l); n = 1 << l
    A, B = _rand_tables(rng, n), _rand_tables(rng, n)
    claim = sum(a * b for a, b in zip(A, B)) % P
    tp = pyref.Transcript(b"q", P); tc = orc.Transcript(b"q")
    polys, rs, finals = pyref.sumcheck_quad_naive(l, A, B, tp, P)
    cpolys, cr, cclaims = orc.sumcheck_quad_prove(orc.to_mont([claim]), l, orc.to_mont(A), orc.to_mont(B), tc)
    assert [orc.from_mont(cp) for cp in cpolys] == polys
    assert orc.from_mont(cr) == rs and orc.from_mont(cclaims) == list(finals)


@pytest.mark.parametrize("l", [10, 14])
def test_cubic_prove_verify_roundtrip(orc, l):
    # mirrors sumcheck.rs:1450-1553: random tables, seed 0xDEADBEEF, prove then verify with a fresh transcript
    rng = np.random.default_rng(0xDEADBEEF)
    n = 1 << l
    def rnd(k):
        a = rng.integers(0, 2**63, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffff); return a   # < p
    A, B, C, taus = rnd(n), rnd(n), rnd(n), rnd(l)
    e = orc.eq_evals(taus)
    ab = orc.f_sub(orc.f_mul(A, B), C)
    claim = orc.f_dot_delayed(e, ab)
    # f_dot_delayed multiplies Montgomery limbs and REDCs once: result is (sum e*ab) in Montgomery form
    tc = orc.Transcript(b"rt")
    polys, r, claims, _ = orc.sumcheck_cubic_prove(claim, taus, A, B, C, tc)
    tv = orc.Transcript(b"rt")
    e_final, rv = orc.sumcheck_verify(polys, 3, claim, tv)
    assert np.array_equal(rv, r)
    # final check: e == eq(tau, r) * (A(r) B(r) - C(r))
    pi = orc.from_mont
    tb = 1
    for t, x in zip(pi(taus), pi(r)):
        tb = tb * (t * x + (1 - t) * (1 - x)) % P
    a, b, c = pi(claims)
    assert pi(e_final)[0] == tb * (a * b - c) % P


@pytest.mark.parametrize("l", [1, 2, 5, 9])
def test_zero_check_round0_is_the_first_round_of_a_satisfied_cubic(orc, l):
    """evaluation_points_zero_check_round0 (sumcheck.rs:1163-1271): on Az o Bz = Cz the first round polynomial of
    prove_cubic_with_three_inputs, evaluated at 0, 2, 3, equals the shortcut that only sums t(inf)."""
    rng = np.random.default_rng(l)

    def rf(n):
        a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a
    n = 1 << l
    A, B, taus = rf(n), rf(n), rf(l)
    polys = orc.sumcheck_cubic_prove(np.zeros((1, 4), dtype=np.uint64), taus, A.copy(), B.copy(), orc.f_mul(A, B), orc.Transcript(b"x"))[0]
    c = orc.from_mont(np.asarray(polys).reshape(-1, 4)[:4]); P = orc.P_T256_SCALAR
    assert orc.from_mont(orc.zero_check_round0(taus, A, B)) == [sum(ci * pow(x, k, P) for k, ci in enumerate(c)) % P for x in (0, 2, 3)]
