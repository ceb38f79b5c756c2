"""Committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py).
CPU: the oracle reproduces the reference's own known answers and its frozen small sum-check vectors.
GPU (-m gpu): the CUDA provers reproduce the same frozen vectors through the C ABI."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    return json.load(open(os.path.join(HERE, "golden", name)))


def _u(hexlist, shape):
    return np.array([int(x, 16) for x in hexlist], dtype=np.uint64).reshape(shape)


def test_reference_kats(orc):
    from oracle import pyref
    k = _load("reference_kats.json")
    for v in k["keccak256"]:
        assert orc.keccak256(bytes.fromhex(v["input_hex"])).hex() == v["digest_hex"]
    tp = k["transcript_pallas"]
    t = orc.Transcript(tp["label"].encode())
    for op, label, val in tp["steps"]:
        if op == "absorb":
            t.absorb_scalars(label.encode(), orc.to_mont([val], orc.FPALLAS), orc.FPALLAS)
        else:
            c = orc.from_mont(t.squeeze(label.encode(), orc.FPALLAS), orc.FPALLAS)[0]
            assert c.to_bytes(32, "little").hex() == val
    s = k["spmv"]
    rows = len(s["matrix"]); data, idx, ptr = [], [], [0]
    for row in s["matrix"]:
        for j, v in enumerate(row):
            if v:
                data.append(v); idx.append(j)
        ptr.append(len(idx))
    out = orc.csr_multiply_vec(rows, orc.to_mont(data), np.array(idx, dtype=np.uint32), np.array(ptr, dtype=np.uint32), orc.to_mont(s["z"]))
    assert orc.from_mont(out) == s["out"]
    from spartan2_b200.frontend import Sha256Circuit
    c = Sha256Circuit(bytes(64), kind="compression")
    assert c.num_cons_unpadded == k["sha256_compression_constraints"]["total"]
    assert pyref.keccak256(b"") .hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"


def _golden_sc():
    g = _load("sumcheck_small.json"); l = g["l"]; n = 1 << l
    return g, l, _u(g["A"], (n, 4)), _u(g["B"], (n, 4)), _u(g["C"], (n, 4)), _u(g["taus"], (l, 4)), _u(g["claim"], (1, 4)), _u(g["quad_claim"], (1, 4))


def test_oracle_reproduces_frozen_sumcheck_vectors(orc):
    g, l, A, B, Cz, taus, claim, qclaim = _golden_sc()
    t = orc.Transcript(b"golden"); t.squeeze(b"s")
    assert [t.state()[0].hex(), t.state()[1]] == g["ts0"]
    polys, r, claims, _ = orc.sumcheck_cubic_prove(claim, taus, A, B, Cz, t)
    assert np.array_equal(polys, _u(g["cubic"]["polys"], polys.shape)) and np.array_equal(r, _u(g["cubic"]["r"], r.shape))
    assert np.array_equal(claims, _u(g["cubic"]["claims"], claims.shape)) and [t.state()[0].hex(), t.state()[1]] == g["cubic"]["ts"]
    qp, qr, qc = orc.sumcheck_quad_prove(qclaim, l, A, B, t)
    assert np.array_equal(qp, _u(g["quad"]["polys"], qp.shape)) and np.array_equal(qr, _u(g["quad"]["r"], qr.shape))
    assert np.array_equal(qc, _u(g["quad"]["claims"], qc.shape)) and [t.state()[0].hex(), t.state()[1]] == g["quad"]["ts"]


@pytest.mark.gpu
def test_cuda_reproduces_frozen_sumcheck_vectors():
    import spartan2_b200 as sp
    g, l, A, B, Cz, taus, claim, qclaim = _golden_sc()
    ctx = sp.Context(0)
    ts = sp.TranscriptState.make(bytes.fromhex(g["ts0"][0]), g["ts0"][1])
    polys, r, claims = sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, claim, taus, A, B, Cz, ts)
    assert np.array_equal(polys, _u(g["cubic"]["polys"], polys.shape)) and np.array_equal(r, _u(g["cubic"]["r"], r.shape))
    assert np.array_equal(claims, _u(g["cubic"]["claims"], claims.shape)) and [ts.get()[0].hex(), ts.get()[1]] == g["cubic"]["ts"]
    qp, qr, qc = sp.SumcheckProof.prove_quad(ctx, qclaim, l, A, B, ts)
    assert np.array_equal(qp, _u(g["quad"]["polys"], qp.shape)) and np.array_equal(qr, _u(g["quad"]["r"], qr.shape))
    assert np.array_equal(qc, _u(g["quad"]["claims"], qc.shape)) and [ts.get()[0].hex(), ts.get()[1]] == g["quad"]["ts"]
    ctx.close()


def _golden_nifs():
    g = _load("nifs_small.json"); n, left, right = g["n"], g["left"], g["right"]; N = left * right
    L = [_u(g[k], (n * N, 4)) for k in ("A", "B", "C")]
    S64 = [np.array([int(x) for x in g[k]], dtype=np.int64) for k in ("A64", "B64", "C64")]
    return g, n, left, right, N, L, S64, np.array(g["large_positions"], dtype=np.uint64), _u(g["tau"], (1, 4)), _u(g["rhos"], (2, 4)), _u(g["r_b"], (1, 4))


def test_oracle_reproduces_frozen_nifs_small_vectors(orc):
    g, n, left, right, N, L, S64, lp, tau, rhos, r_b = _golden_nifs()
    E = orc.pow_split_evals(tau, left, right)
    union = set()
    for k in range(3):
        for b in range(n):
            v, lg = orc.to_small_vec_or_zero(L[k][b * N:(b + 1) * N]); union |= set(int(x) for x in lg)
            keep = np.array([i not in set(int(x) for x in lp) for i in range(N)])
            assert np.array_equal(v[keep], S64[k][b * N:(b + 1) * N][keep])
    assert sorted(union) == [int(x) for x in lp]
    assert np.array_equal(orc.nifs_round0_small(rhos, left, right, E, L[0], L[1], S64[0], S64[1], lp, N, n), _u(g["round0"], (2, 4)))
    assert np.array_equal(orc.nifs_round(0, rhos, left, right, E, L[0], L[1], L[2], N, n), _u(g["round0"], (2, 4)))
    assert np.array_equal(orc.nifs_cvals_small(left, right, E, L[2], S64[2], lp, N, n), _u(g["c_vals"], (n, 4)))
    F = [orc.nifs_fold(L[k], N, n, r_b) for k in range(3)]
    assert np.array_equal(F[0], _u(g["folded_A"], F[0].shape))
    assert np.array_equal(orc.nifs_round(1, rhos, left, right, E, F[0], F[1], F[2], N, n // 2), _u(g["round1"], (2, 4)))


@pytest.mark.gpu
def test_cuda_reproduces_frozen_nifs_small_vectors():
    import spartan2_b200 as sp
    g, n, left, right, N, L, S64, lp, tau, rhos, r_b = _golden_nifs()
    ctx = sp.Context(0)
    E = sp.PowPolynomial.split_evals(ctx, tau, left, right)
    dL = [ctx.upload(x) for x in L]; dE = ctx.upload(E)
    d64, dpos, nl = sp.SmallValue.to_small_layers(ctx, dL, n, N)
    assert nl == len(lp) and np.array_equal(dpos.download((max(nl, 1),), dtype=np.uint64)[:nl], lp)
    for k in range(3):
        assert np.array_equal(d64[k].download((n * N,), dtype=np.int64), S64[k])
    assert np.array_equal(sp.SmallValue.nifs_round0(ctx, rhos, left, right, dE, d64[0], d64[1], dL[0], dL[1], dpos, nl, N, n), _u(g["round0"], (2, 4)))
    assert np.array_equal(sp.SmallValue.cvals(ctx, left, right, dE, dL[2], d64[2], dpos, nl, N, n), _u(g["c_vals"], (n, 4)))
    nifs = sp.NeutronNovaNIFS(ctx, E, left, right, dL[0], dL[1], dL[2], n)
    assert np.array_equal(nifs.round_eval(rhos), _u(g["round0"], (2, 4)))
    nifs.fold(r_b)
    assert np.array_equal(nifs.round_eval(rhos), _u(g["round1"], (2, 4)))
    fa = dL[0].download((n * N, 4)).reshape(n, N, 4)[0::2].reshape(-1, 4)         # folded layers sit at the even slots
    assert np.array_equal(fa, _u(g["folded_A"], fa.shape))
    ctx.close()
