"""NeutronNova building blocks on the device vs the oracle (which tests/test_oracle_neutronnova.py pins to the
definitions): split power table, a complete multi-round NIFS (evaluate -> challenge -> fold layers), witness folding,
the pow-weighted cubic / quadratic per-round evaluation points with binding across all rounds, commitment folding."""
import numpy as np
import pytest

from tests.gpu_util import ctx, rand_fe  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("left,right", [(1, 1), (2, 1), (8, 4), (256, 128)])
def test_pow_split_evals(ctx, orc, left, right):
    import spartan2_b200 as sp
    t = rand_fe(np.random.default_rng(left), 1)
    assert np.array_equal(sp.PowPolynomial.split_evals(ctx, t, left, right), orc.pow_split_evals(t, left, right))


@pytest.mark.parametrize("n,left,right", [(2, 8, 4), (8, 32, 32), (32, 256, 128)])
def test_nifs_rounds_and_folds(ctx, orc, n, left, right):
    import spartan2_b200 as sp
    rng = np.random.default_rng(n)
    N = left * right; ell_b = n.bit_length() - 1
    A, B, Cm = rand_fe(rng, n * N), rand_fe(rng, n * N), rand_fe(rng, n * N)
    tau = rand_fe(rng, 1); rhos = rand_fe(rng, ell_b); r_bs = rand_fe(rng, ell_b)
    E = orc.pow_split_evals(tau, left, right)
    nifs = sp.NeutronNovaNIFS(ctx, E, left, right, ctx.upload(A), ctx.upload(B), ctx.upload(Cm), n)
    oA, oB, oC, m = A, B, Cm, n
    for t in range(ell_b):
        got = nifs.round_eval(rhos)
        want = orc.nifs_round(t, rhos, left, right, E, oA, oB, oC, N, m)
        assert np.array_equal(got, want), t
        nifs.fold(r_bs[t:t + 1])
        oA, oB, oC = (orc.nifs_fold(x, N, m, r_bs[t:t + 1]) for x in (oA, oB, oC))
        m //= 2
    fa, fb, fc = nifs.layer0()
    assert np.array_equal(fa, oA) and np.array_equal(fb, oB) and np.array_equal(fc, oC)
    # the folded layer equals the eq-weighted sum of the original layers (fold_multiple's weights, LSB-first)
    w = sp.weights_from_r(ctx, r_bs, n)
    assert np.array_equal(w, orc.weights_from_r(r_bs, n))
    assert np.array_equal(orc.fold_vectors(A, n, N, w), fa)


def test_fold_multiple(ctx, orc):
    import spartan2_b200 as sp
    rng = np.random.default_rng(4)
    n, dim = 8, 5000
    Ws = rand_fe(rng, n * dim).reshape(n, dim, 4)
    Ws[1] = orc.to_mont([int(x) for x in rng.integers(0, 2, size=dim)])       # a boolean ("small") witness among them
    r_bs = rand_fe(rng, 3)
    got = sp.R1CSWitness.fold_multiple(ctx, r_bs, Ws)
    assert np.array_equal(got, orc.fold_vectors(Ws.reshape(-1, 4), n, dim, orc.weights_from_r(r_bs, n)))


def test_pow_cubic_rounds_with_binding(ctx, orc):
    """All 10 rounds of the NeutronNova outer sum-check shape (one branch): evaluation points, then bind A, B, C."""
    import spartan2_b200 as sp
    rng = np.random.default_rng(8)
    left, right = 32, 32; n = left * right
    tau = rand_fe(rng, 1)
    E = orc.pow_split_evals(tau, left, right)
    A, B, Cm = rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, n)
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cm)
    dpl, dpr = ctx.upload(E[:left]), ctx.upload(E[left:])
    oA, oB, oC = A, B, Cm
    tl = n
    while tl >= 2:
        got = sp.SumcheckRounds.eval_points_cubic_with_outer_pow(ctx, dpl, left, dpr, dA, dB, dC, tl)
        assert np.array_equal(got, orc.pow_cubic_eval(E[:left], E[left:], oA[:tl], oB[:tl], oC[:tl])), tl
        r = rand_fe(rng, 1)
        sp.SumcheckRounds.bind_poly_var_top(ctx, [dA, dB, dC], tl, r)
        oA, oB, oC = (orc.bind_top(x[:tl], r) for x in (oA, oB, oC))
        tl //= 2
        assert np.array_equal(dA.download((tl, 4)), oA)


def test_quad_rounds_with_binding(ctx, orc):
    import spartan2_b200 as sp
    rng = np.random.default_rng(9)
    n = 1 << 12
    A, B = rand_fe(rng, n), rand_fe(rng, n)
    dA, dB = ctx.upload(A), ctx.upload(B)
    oA, oB, tl = A, B, n
    while tl >= 2:
        assert np.array_equal(sp.SumcheckRounds.eval_points_quad(ctx, dA, dB, tl), orc.quad_eval(oA[:tl], oB[:tl]))
        r = rand_fe(rng, 1)
        sp.SumcheckRounds.bind_poly_var_top(ctx, [dA, dB], tl, r)
        oA, oB = orc.bind_top(oA[:tl], r), orc.bind_top(oB[:tl], r)
        tl //= 2


def test_fold_commitments(ctx, orc):
    import spartan2_b200 as sp
    rng = np.random.default_rng(10)
    n, rows = 8, 5
    pts = ctx.test_points(n * rows, seed=3)
    w = rand_fe(rng, n)
    w[0] = orc.to_mont([1])[0]; w[1] = 0                  # unit and zero weights
    assert np.array_equal(sp.fold_commitments(ctx, pts, n, rows, w), orc.fold_commitments(pts, n, rows, w))


@pytest.mark.parametrize("n", [4, 32])
def test_hot_path_device_vs_oracle(ctx, orc, n):
    """HOT LOOPS A-C of NeutronNovaZkSNARK::prove (spartan2_b200.neutronnova.run) on the SHA-256 chain — n = 32 is
    BASELINE config 3 (32 step circuits of one compression each + the core circuit, N = M = 2^15 per instance): the
    same driver over the CUDA backend and over the oracle backend, with identical transcripts; every recorded
    intermediate value (NIFS rounds and polynomials, folded layers and witness, 15 x 2 outer and 16 x 2 inner round
    evaluations, claims, poly_ABC, final evaluations) must be bit-identical, and the verifier's final equations hold."""
    import spartan2_b200 as sp
    from spartan2_b200 import neutronnova as nn
    from tests.neutronnova_ops import OracleOps, sha_chain_instances
    c0, zs, Ws, zc, Wc = sha_chain_instances(n)
    A, B, Cm = c0.matrices()
    S = sp.SplitR1CSShape(ctx, *c0.dims(), A, B, Cm)
    tr_d, tr_o = [], []
    out_d = nn.run(nn.DeviceOps(ctx, S), sp.Keccak256Transcript(b"neutronnova_prove"), c0.num_cons, zs, Ws, zc, Wc, trace=tr_d)
    out_o = nn.run(OracleOps(orc.Shape(*c0.dims(), A, B, Cm), c0.dims()), orc.Transcript(b"neutronnova_prove"), c0.num_cons, zs, Ws, zc, Wc, trace=tr_o)
    assert [t[0] for t in tr_d] == [t[0] for t in tr_o]
    for (name, a), (_, b) in zip(tr_d, tr_o):
        assert np.array_equal(a, b), name
    assert out_d["outer_ok"] and out_d["inner_ok"] and out_o["outer_ok"] and out_o["inner_ok"]
    for k in ("eval_W_step", "eval_W_core", "T_out"):
        assert out_d[k] == out_o[k]
    S.free()


@pytest.mark.parametrize("n", [2, 32])
def test_fused_prover_vs_oracle(ctx, orc, n):
    """The fused path of the library (sp2_neutronnova_prep_prove + sp2_neutronnova_prove: round loop, scalar algebra and
    transcript in C++, tables device-resident) against the per-round driver over the ORACLE backend with an identical
    transcript: every value a verifier would see — NIFS sums and polynomials, both branches' 15 outer and 16 inner
    round evaluations and polynomials, claims, tau(r_x), the folded layers / witness / poly_ABC probes, the final
    evaluations — bit-identical; the challenges r_b, r_x, r_y and the transcript end state agree."""
    import spartan2_b200 as sp
    from spartan2_b200 import neutronnova as nn
    from tests.neutronnova_ops import OracleOps, sha_chain_instances
    c0, zs, Ws, zc, Wc = sha_chain_instances(n)
    A, B, Cm = c0.matrices()
    S = sp.SplitR1CSShape(ctx, *c0.dims(), A, B, Cm)
    prover = nn.NeutronNovaProver(ctx, S, zs, zc)
    for rep in range(2):                                   # prove twice from the same prep state: the cached layers are not consumed
        ts_d = sp.Keccak256Transcript(b"neutronnova_prove")
        v, ph = prover.prove(ts_d)
    tr_o = []
    ts_o = orc.Transcript(b"neutronnova_prove")
    out_o = nn.run(OracleOps(orc.Shape(*c0.dims(), A, B, Cm), c0.dims()), ts_o, c0.num_cons, zs, Ws, zc, Wc, trace=tr_o)
    got = nn.NeutronNovaProver.as_trace(v)
    names = [t[0] for t in tr_o if t[0] != "E"]
    assert sorted(names) == sorted(got.keys())
    for name, b in tr_o:
        if name != "E":
            assert np.array_equal(got[name].reshape(-1), b.reshape(-1)), name
    assert v["outer_ok"] and v["inner_ok"] and out_o["outer_ok"] and out_o["inner_ok"]
    for k in ("r_b", "r_x", "r_y"):
        assert np.array_equal(v[k].reshape(-1), np.asarray(out_o[k], dtype=np.uint64).reshape(-1)), k
    from spartan2_b200 import _fq as fq
    assert fq.to_int(v["T_out"]) == out_o["T_out"]
    assert fq.to_ints(v["eval_W"]) == [out_o["eval_W_step"], out_o["eval_W_core"]]
    st, rnd = ts_o.state()
    assert ts_d.state().get() == (st, rnd)
    assert ph["total"] > 0
    prover.free(); S.free()


def test_fused_prover_argument_errors(ctx, orc):
    """SpartanError mirroring (src/errors.rs:12-110) at the fused entry points: a step count that is not a power of two
    (the reference pads to one, neutronnova_zk.rs:529-552 — the caller's job here) and z vectors of the wrong length."""
    import spartan2_b200 as sp
    from spartan2_b200 import neutronnova as nn
    from tests.neutronnova_ops import sha_chain_instances
    c0, zs, Ws, zc, Wc = sha_chain_instances(2)
    A, B, Cm = c0.matrices()
    S = sp.SplitR1CSShape(ctx, *c0.dims(), A, B, Cm)
    with pytest.raises(sp.SpartanError) as ei:
        nn.NeutronNovaProver(ctx, S, [zs[0], zs[1], zs[0]], zc)
    assert ei.value.kind == "InvalidInputLength"
    with pytest.raises(sp.SpartanError) as ei:
        nn.NeutronNovaProver(ctx, S, [zs[0][:-1], zs[1][:-1]], zc)
    assert ei.value.kind == "InvalidWitnessLength"
    S.free()


@pytest.mark.parametrize("n,world", [(4, 2), (8, 4), (2, 2)])
def test_sharded_prover_two_ranks_on_one_gpu(ctx, orc, n, world):
    """The instance-sharded prove (sp2_neutronnova_prep_prove_sharded / _prove_sharded) with `world` ranks
    emulated on ONE GPU: one context + prep state per rank, one host thread per rank, and an in-process all-gather
    (threading.Barrier + device copies) as the `allgather` callback — the callback variant of every exchange (round sums
    on the host, surviving layers and witness partials on the device).  Every rank must return exactly what the
    single-GPU fused prove of all n instances returns (which test_fused_prover_vs_oracle pins to the oracle)."""
    import ctypes as C
    import threading
    import spartan2_b200 as sp
    from spartan2_b200 import neutronnova as nn
    from tests.neutronnova_ops import sha_chain_instances
    c0, zs, Ws, zc, Wc = sha_chain_instances(n)
    A, B, Cm = c0.matrices()
    S = sp.SplitR1CSShape(ctx, *c0.dims(), A, B, Cm)
    single = nn.NeutronNovaProver(ctx, S, zs, zc)
    want, _ = single.prove(sp.Keccak256Transcript(b"neutronnova_prove"))
    single.free()
    nl = n // world
    ctxs = [sp.Context(0) for _ in range(world)]
    shapes = [sp.SplitR1CSShape(c, *c0.dims(), A, B, Cm) for c in ctxs]
    bar = threading.Barrier(world)
    slots = [None] * world

    def make_allgather(rank):
        def allgather(send, nbytes, recv, on_device):
            slots[rank] = (send, nbytes)
            bar.wait()
            for q in range(world):
                src, nb = slots[q]
                assert nb == nbytes
                if on_device:       # same GPU: a device-to-device copy stands in for the peer transfer
                    if recv + q * nbytes != src:
                        ctxs[rank].check(ctxs[rank].L.sp2_dev_copy(ctxs[rank].h, C.c_void_p(recv + q * nbytes), C.c_void_p(src), C.c_uint64(nbytes)))
                else:
                    C.memmove(recv + q * nbytes, src, nbytes)
            if on_device:
                ctxs[rank].synchronize()
            bar.wait()              # nobody reuses its send buffer before every rank has copied it
        return allgather
    provers = [nn.NeutronNovaProver(ctxs[r], shapes[r], zs[r * nl:(r + 1) * nl], zc, rank=r, nranks=world, allgather=make_allgather(r)) for r in range(world)]
    outs, errs = [None] * world, []

    def run(r):
        try:
            outs[r], _ = provers[r].prove(sp.Keccak256Transcript(b"neutronnova_prove"))
        except Exception as e:      # a failing rank must not leave the others waiting in the barrier
            errs.append(e); bar.abort()
    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ths]
    [t.join(timeout=120) for t in ths]
    assert not errs, errs
    for r in range(world):
        for k, v in want.items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(outs[r][k], v), (r, k)
        assert outs[r]["outer_ok"] and outs[r]["inner_ok"]
    for p in provers:
        p.free()
    for s_ in shapes + [S]:
        s_.free()
    for c in ctxs:
        c.close()
