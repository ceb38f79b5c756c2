"""The full NeutronNova prove on the device — commitment half included (sp2_neutronnova_prep_commit +
sp2_neutronnova_snark_prove: per-step rerandomisation, commit_zeros, the instance transcript, HOT LOOPS A-C, fold_blinds /
fold_commitments_partial, the c_eval fold, PCS::prove on the folded witness) — against the oracle's C driver of the same
non-ZK protocol (oracle/oracle.c: orc_neutronnova_prove, which follows src/neutronnova_zk.rs:1609-2093 and folds
commitments as GROUP ELEMENTS the way the reference does, while the device re-commits by linearity): every proof field
bit-identical, and the oracle's verifier (orc_neutronnova_verify, after :2096-2330 + zk.rs:600-940) accepts the
device-made proof and rejects it when tampered."""
import numpy as np
import pytest

from tests.gpu_util import ctx, rand_fe  # noqa: F401
from tests.test_oracle_neutronnova_snark import nn_case, prove, step_X

pytestmark = pytest.mark.gpu

FIELDS = ["comm_W_steps", "comm_W_core", "nifs_polys", "outer_polys", "claims_outer", "inner_polys", "eval_W", "blind_eval_W",
          "delta", "beta", "z_vec", "z_delta", "z_beta"]


def _device_prove(ctx, c):
    import spartan2_b200 as sp
    from spartan2_b200 import neutronnova as nn
    K = c["keys"]
    S = sp.SplitR1CSShape(ctx, *c["dims"], *c["mats"])
    ck = sp.CommitmentKey(ctx, K.ck, K.h, K.ck_s, K.h_s)
    prover = nn.NeutronNovaProver(ctx, S, list(c["zs"]), c["zc"])
    comm_s, comm_c = prover.commit(ck, c["b_old_s"], c["b_old_c"])
    assert np.array_equal(comm_s, c["comm_pre_s"]) and np.array_equal(comm_c, c["comm_pre_c"])      # prep_prove's commitments
    r = c["rand"].a
    outs = [prover.snark_prove(c["vk"], *r) for _ in range(2)]                                        # repeatable on one prep state
    for k in FIELDS:
        assert np.array_equal(outs[0][0][k], outs[1][0][k]), k
    prover.free(); S.free(); ck.free()
    return outs[1]


def _check(orc, c, v):
    P = prove(orc, c)
    # probes of the commitment half first (clearer failures than a downstream transcript mismatch)
    n, rows = c["n"], c["rows"]
    one = orc.to_mont([1])
    for k in ("comm_W_steps", "comm_W_core", "eval_W"):
        assert np.array_equal(np.asarray(v[k]).reshape(-1), getattr(P, k).reshape(-1)), "parity: " + k
    ce = np.concatenate([orc.point_add(orc.scalar_mul(c["keys"].ck_s, P.eval_W[b:b + 1]), orc.scalar_mul(c["keys"].h_s, c["rand"].a[2][b:b + 1])) for b in range(2)])
    assert np.array_equal(v["comm_eval_W"], ce), "comm_eval_W"
    assert np.array_equal(v["c_eval"].reshape(-1), P.debug["c_eval"].reshape(-1)), "c_eval"
    w = orc.weights_from_r(v["r_b"].reshape(-1, 4), n)
    folded = orc.fold_commitments(P.comm_W_steps, n, rows, w)
    comm = orc.fold_commitments(np.concatenate([folded, P.comm_W_core]), 2, rows, np.concatenate([one, P.debug["c_eval"]]))
    assert np.array_equal(v["comm_fold"], comm), "comm_fold (fold by linearity vs fold of group elements)"
    for k in FIELDS:
        assert np.array_equal(np.asarray(v[k]).reshape(-1), getattr(P, k).reshape(-1)), "parity: " + k
    assert np.array_equal(v["c_eval"].reshape(-1), P.debug["c_eval"].reshape(-1))
    assert np.array_equal(v["T_out"].reshape(-1), P.debug["T_out"].reshape(-1)) and np.array_equal(v["tau_at_rx"].reshape(-1), P.debug["tau_at_rx"].reshape(-1))
    assert v["outer_ok"] and v["inner_ok"]
    # the oracle verifier on the DEVICE-made proof
    V = orc.NnProof(c["n"], c["dims"][0], c["M"], c["width"])
    for k in FIELDS:
        getattr(V, k)[...] = np.asarray(v[k]).reshape(getattr(V, k).shape)
    sx, cx = step_X(c)
    assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, V) == 0
    V.z_vec[1, 2] ^= np.uint64(8)
    assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, V) == -5
    V.z_vec[1, 2] ^= np.uint64(8)
    V.comm_W_core[0, 0] ^= np.uint64(1)
    assert orc.neutronnova_verify(c["O"], c["keys"], c["vk"], sx, cx, V) != 0


@pytest.mark.parametrize("n,lc,lv,width,npub", [(2, 5, 7, 32, 0), (4, 6, 8, 32, 2), (8, 6, 8, 64, 1), (4, 7, 9, 256, 3)])
def test_snark_bit_exact_small_chains(ctx, orc, n, lc, lv, width, npub):
    """random R1CS chains incl. public IO (X is folded too), 2 to 8 commitment rows"""
    c = nn_case(orc, n=n, lc=lc, lv=lv, width=width, npub=npub, seed=20 + n)
    v, ph = _device_prove(ctx, c)
    _check(orc, c, v)
    assert ph["total"] > 0


@pytest.mark.parametrize("n,lc,lv,width,npub", [(2, 5, 7, 32, 0), (4, 7, 9, 256, 3)])
def test_snark_bit_exact_one_launch_per_round(ctx, orc, monkeypatch, n, lc, lv, width, npub):
    """SP2_NN_PIPE=0 (measurement switch, read per prove): the batched sum-checks as one bind + evaluate launch per round with the host
    waiting for each (nifs.cu: k_nn_outer_round / k_nn_inner_round) instead of the default coefficient kernels one launch ahead of the
    host — the same proof, bit for bit"""
    monkeypatch.setenv("SP2_NN_PIPE", "0")
    c = nn_case(orc, n=n, lc=lc, lv=lv, width=width, npub=npub, seed=20 + n)
    v, ph = _device_prove(ctx, c)
    _check(orc, c, v)


def sha_case(orc, ctx, n, seed=5):
    from tests.neutronnova_ops import sha_chain_instances
    c0, zs, Ws, zc, Wc = sha_chain_instances(n)
    A, B, Cm = c0.matrices(); d = c0.dims()
    width = 2048
    pts = ctx.test_points(width + 3, seed=9)
    keys = orc.Keys(pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3])
    M = d[2] + d[3] + d[4]; pre = d[3]; rows = M // width; pre_rows = pre // width
    rng = np.random.default_rng(seed)
    zs = np.stack(zs)
    b_old_s = rand_fe(rng, n * pre_rows); b_old_c = rand_fe(rng, pre_rows)
    orc.set_threads(orc.max_threads())
    comm_pre_s = np.concatenate([orc.hyrax_commit(keys.ck, keys.h, zs[i][:pre], b_old_s[i * pre_rows:(i + 1) * pre_rows], is_small=True) for i in range(n)])
    comm_pre_c = orc.hyrax_commit(keys.ck, keys.h, zc[:pre], b_old_c, is_small=True)
    rand = orc.NnRand(rand_fe(rng, n * rows), rand_fe(rng, rows), rand_fe(rng, 2), rand_fe(rng, width), rand_fe(rng, 1), rand_fe(rng, 1))
    return dict(O=orc.Shape(*d, A, B, Cm), keys=keys, vk=bytes(range(32)), zs=zs, zc=zc, M=M, npub=d[5], rows=rows, pre_rows=pre_rows, width=width, n=n,
                comm_pre_s=comm_pre_s, b_old_s=b_old_s, comm_pre_c=comm_pre_c, b_old_c=b_old_c, rand=rand, dims=d, mats=(A, B, Cm))


@pytest.mark.parametrize("n", [4, 32])
def test_snark_bit_exact_sha256_chain(ctx, orc, n):
    """the benchmark circuits (benches/sha256_neutronnova.rs: step i hashes [i; 64], core one zero block; N = M = 2^15, 13
    precommitted + 3 rest rows of 2048); n = 32 is BASELINE config 3"""
    c = sha_case(orc, ctx, n)
    v, ph = _device_prove(ctx, c)
    _check(orc, c, v)
    orc.set_threads(1)
