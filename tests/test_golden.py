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
