"""The oracle's non-ZK NeutronNova driver (oracle.c: orc_neutronnova_prove / orc_neutronnova_verify — the checker the
CUDA path's full NeutronNova prove is compared with): prove -> verify accepts, every tampered component is rejected at
the right check, and the Hyrax commitment family it uses (commit_without_blind, commit_incremental,
rerandomize_commitment, fold_blinds, fold_commitments_partial) agrees with its definitions.  CPU only — the reference's
own integration-test pattern (src/neutronnova_zk.rs:2479-2502: prove then verify)."""
import numpy as np
import pytest

from tests.curve_util import points
from tests.r1cs_util import chain_instances


def rf(rng, n):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a


def nn_case(orc, n=4, lc=6, lv=8, width=32, npub=2, seed=11):
    d, mats, zs, zc = chain_instances(n, lc, lv, num_public=npub, seed=seed)
    O = orc.Shape(*d, *mats)
    M = d[2] + d[3] + d[4]; pre = d[3]; rows = M // width; pre_rows = pre // width
    pts = points(orc, width + 3, seed=33)
    keys = orc.Keys(pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3])
    rng = np.random.default_rng(seed)
    zs = np.stack(zs)
    b_old_s = rf(rng, n * pre_rows); b_old_c = rf(rng, pre_rows)
    comm_pre_s = np.concatenate([orc.hyrax_commit(keys.ck, keys.h, zs[i][:pre], b_old_s[i * pre_rows:(i + 1) * pre_rows], is_small=False) for i in range(n)])
    comm_pre_c = orc.hyrax_commit(keys.ck, keys.h, zc[:pre], b_old_c, is_small=False)
    rand = orc.NnRand(rf(rng, n * rows), rf(rng, rows), rf(rng, 2), rf(rng, width), rf(rng, 1), rf(rng, 1))
    vk = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
    return dict(O=O, keys=keys, vk=vk, zs=zs, zc=zc, M=M, npub=npub, rows=rows, pre_rows=pre_rows, width=width, n=n,
                comm_pre_s=comm_pre_s, b_old_s=b_old_s, comm_pre_c=comm_pre_c, b_old_c=b_old_c, rand=rand, dims=d, mats=mats)


def prove(orc, c):
    return orc.neutronnova_prove(c["O"], c["keys"], c["vk"], c["zs"], c["zc"], c["comm_pre_s"], c["b_old_s"], c["comm_pre_c"], c["b_old_c"], c["rand"])


def step_X(c):
    M, npub = c["M"], c["npub"]
    return c["zs"][:, M + 1:M + 1 + npub].reshape(-1, 4), c["zc"][M + 1:M + 1 + npub]


@pytest.mark.parametrize("n,lc,lv,npub", [(2, 5, 7, 0), (4, 6, 8, 2), (8, 6, 8, 1)])
def test_prove_verify_roundtrip_and_tamper(orc, n, lc, lv, npub):
    c = nn_case(orc, n=n, lc=lc, lv=lv, npub=npub)
    P = prove(orc, c)
    sx, cx = step_X(c)
    assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, P) == 0
    # the commitments the proof carries are the rerandomised ones: U_i.comm_W = commit(W_i, new blinds)
    rows, width, M = c["rows"], c["width"], c["M"]
    for i in (0, n - 1):
        want = orc.hyrax_commit(c["keys"].ck, c["keys"].h, c["zs"][i][:M], c["rand"].a[0][i * rows:(i + 1) * rows], is_small=False)
        assert np.array_equal(P.comm_W_steps[i * rows:(i + 1) * rows], want)
    # every tampered component is rejected, at the check that covers it
    for field, idx, code in [("nifs_polys", 1, -2), ("outer_polys", 5, -3), ("claims_outer", 2, -3), ("inner_polys", 7, -4), ("eval_W", 0, -4),
                             ("eval_W", 1, -4), ("blind_eval_W", 0, -5), ("z_vec", 3, -5), ("z_delta", 0, -5), ("z_beta", 0, -5)]:
        a = getattr(P, field); old = a[idx, 0]
        a[idx, 0] = old ^ np.uint64(1)
        rc = orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, P)
        assert rc != 0 and rc == code, (field, rc)
        a[idx, 0] = old
    # a swapped step commitment changes the transcript: the first round check fails
    P.comm_W_steps[[0, rows]] = P.comm_W_steps[[rows, 0]]
    assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, P) != 0
    P.comm_W_steps[[0, rows]] = P.comm_W_steps[[rows, 0]]
    # wrong public IO
    if npub:
        sx2 = sx.copy(); sx2[0, 0] ^= np.uint64(1)
        assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx2, cx, P) != 0
    assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, P) == 0


def test_unsatisfied_instance_is_rejected(orc):
    c = nn_case(orc, n=4)
    c["zs"][2, 3] = orc.to_mont([5])[0]           # break one witness entry of step 2 (its precommitted commitment is stale too)
    P = prove(orc, c)
    sx, cx = step_X(c)
    assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, P) != 0


def test_commitment_family_definitions(orc):
    rng = np.random.default_rng(3); width = 16
    pts = points(orc, width + 3, seed=33); ck, h = pts[:width], pts[width:width + 1]
    v = rf(rng, 3 * width + 5); v[width:2 * width] = 0                     # a zero row in the middle, ragged last row
    blinds = rf(rng, 4)
    raw = orc.hyrax_commit_without_blind(ck, v)
    assert not raw[1].any()                                                # identity for the all-zero row
    full = orc.hyrax_commit(ck, h, v, blinds)
    zero_blind = np.zeros_like(blinds)
    assert np.array_equal(orc.hyrax_commit(ck, h, v, zero_blind)[[0, 2, 3]], raw[[0, 2, 3]])
    # commit_incremental: raw + delta + blind == commit(v + delta)
    delta = np.zeros_like(v); delta[3] = rf(rng, 1); delta[2 * width + 1] = rf(rng, 1)
    inc = orc.hyrax_commit_incremental(ck, h, raw, delta, blinds)
    assert np.array_equal(inc, orc.hyrax_commit(ck, h, orc.f_add(v, delta), blinds))
    # rerandomize: comm + (r_new - r_old) h == commit(v, r_new)
    r_new = rf(rng, 4)
    assert np.array_equal(orc.hyrax_rerandomize(h, full, blinds, r_new), orc.hyrax_commit(ck, h, v, r_new))
    # small scalars through msm_small
    bits = orc.to_mont([int(b) for b in rng.integers(0, 2, size=2 * width)])
    assert np.array_equal(orc.hyrax_commit_without_blind(ck, bits, is_small=True), orc.hyrax_commit_without_blind(ck, bits, is_small=False))
    # folds: partial == full fold when the rest rows are blind * h; fold_blinds is the weighted sum
    n, rows, data_rows = 4, 3, 2
    Ws = rf(rng, n * rows * width).reshape(n, rows * width, 4); Ws[:, data_rows * width:] = 0
    bl = rf(rng, n * rows); w = rf(rng, n)
    comms = np.concatenate([orc.hyrax_commit(ck, h, Ws[i], bl[i * rows:(i + 1) * rows]) for i in range(n)])
    fb = orc.fold_blinds(bl, n, rows, w)
    fm = orc.from_mont; P = orc.P_T256_SCALAR
    assert fm(fb) == [sum(fm(w)[k] * fm(bl)[k * rows + r] for k in range(n)) % P for r in range(rows)]
    part = orc.fold_commitments_partial(comms, n, rows, w, data_rows, fb, h)
    assert np.array_equal(part, orc.fold_commitments(comms, n, rows, w))
    # ... and both equal the commitment of the folded witness with the folded blinds (linearity, what the CUDA path uses)
    Wf = orc.fold_vectors(Ws.reshape(-1, 4), n, rows * width, w)
    assert np.array_equal(part, orc.hyrax_commit(ck, h, Wf, fb))
