"""The multi-GPU data path with >= 2 ranks on ONE GPU: every rank is its own context + stream + host
thread, and the ranks exchange raw device pointers instead of CUDA-IPC handles (sp2_comm_connect_ptrs,
sp2_neutronnova_prep_connect_ptrs) — so the kernels that cross GPUs over NVLink in production run here unchanged over
local memory: sumcheck.cu exchange_sums (per-round sums through the peer mailboxes, dc.n > 1), k_shard_gather /
k_shard_barrier (hand-off of the small rounds), the sharded SpMV / poly_ABC of sp2_spartan_prove_sharded, and
nifs.cu k_publish_xchg / k_nn_scatter / k_nn_barrier.  Every rank's output must equal the single-GPU result, which the
other -m gpu tests pin to the oracle (and which is compared with the oracle again here).  Also: a rank whose peer never
arrives gets SP2_ERR_INTERNAL after the bounded device wait instead of a wedged GPU, and sp2_comm_reset recovers."""
import threading
import time

import numpy as np
import pytest

from tests.curve_util import points
from tests.gpu_util import ctx, rand_fe, ts_pair  # noqa: F401
from tests.r1cs_util import dims, random_r1cs

pytestmark = pytest.mark.gpu


def _threads(world, fn):
    """fn(rank) on one host thread per rank (the C-ABI calls block; ctypes drops the GIL); returns the results."""
    outs, errs = [None] * world, []

    def run(r):
        try:
            outs[r] = fn(r)
        except Exception as e:      # noqa: BLE001
            errs.append((r, e))
    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ths]
    [t.join(timeout=300) for t in ths]
    assert not any(t.is_alive() for t in ths), "a rank is still running"
    assert not errs, errs
    return outs


@pytest.fixture()
def ranks():
    import spartan2_b200 as sp
    made = []

    def make(world):
        # device objects of earlier tests that sit in reference cycles (pytest.raises keeps frames alive) would otherwise be freed by
        # the cyclic collector at an arbitrary later allocation — a cudaFree in the middle of kernels that spin on a peer
        import gc
        gc.collect()
        cs = [sp.Context(0) for _ in range(world)]
        made.extend(cs)
        return cs
    yield make
    for c in made:
        c.close()


@pytest.mark.parametrize("world,l", [(2, 17), (4, 18), (2, 16), (8, 19)])
def test_sharded_sumchecks_exchange_in_kernel(ctx, orc, ranks, world, l):
    """Cubic and quadratic sum-checks over tables split cyclically on the low index bits: rounds with > 2^15 global entries
    run sharded (local pairs, the <= 3 partial sums cross ranks inside the round kernel's last CTA), then the shards are
    gathered by peer stores and every rank finishes redundantly.  All ranks == single GPU == oracle."""
    import spartan2_b200 as sp
    rng = np.random.default_rng(1000 + l); n = 1 << l
    A, B, Cz, taus = rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, l)
    claim = orc.f_dot_delayed(orc.eq_evals(taus), orc.f_sub(orc.f_mul(A, B), Cz))
    qclaim = orc.f_dot_delayed(A, B)
    t_or, ts = ts_pair(orc)
    st0 = ts.get()
    want_c = orc.sumcheck_cubic_prove(claim, taus, A, B, Cz, t_or)[:3]
    t_q, tsq = ts_pair(orc, b"q")
    stq = tsq.get()
    want_q = orc.sumcheck_quad_prove(qclaim, l, A, B, t_q)
    cs = ranks(world)
    # warm every rank's scratch with the single-GPU provers (cudaMalloc / cudaFree must not happen while a peer's kernel
    # spins on this rank's next launch — in production every rank is its own process)
    for c in cs:
        t1 = sp.TranscriptState.make(*st0)
        got = sp.SumcheckProof.prove_cubic_with_three_inputs(c, claim, taus, A, B, Cz, t1)
        for a, b in zip(got, want_c):
            assert np.array_equal(a, b)
        t2 = sp.TranscriptState.make(*stq)
        got = sp.SumcheckProof.prove_quad(c, qclaim, l, A, B, t2)
        for a, b in zip(got, want_q):
            assert np.array_equal(a, b)
    comms = sp.Comm.in_process(cs)
    shards = [[c.upload(sp.shard_cyclic(T, world, r)) for T in (A, B, Cz)] for r, c in enumerate(cs)]

    def cubic(r):
        t = sp.TranscriptState.make(*st0)
        out = comms[r].prove_cubic_with_three_inputs(claim, taus, *shards[r], t)
        return out, t.get()
    for (polys, rr, claims), st in _threads(world, cubic):
        assert np.array_equal(polys, want_c[0]) and np.array_equal(rr, want_c[1]) and np.array_equal(claims, want_c[2])
        assert st == t_or.state()
    shards = [[c.upload(sp.shard_cyclic(T, world, r)) for T in (A, B)] for r, c in enumerate(cs)]

    def quad(r):
        t = sp.TranscriptState.make(*stq)
        out = comms[r].prove_quad(qclaim, l, *shards[r], t)
        return out, t.get()
    for (polys, rr, claims), st in _threads(world, quad):
        assert np.array_equal(polys, want_q[0]) and np.array_equal(rr, want_q[1]) and np.array_equal(claims, want_q[2])
        assert st == t_q.state()
    for cm in comms:
        cm.free()


def _prove_case(orc, seed, lc, lv, width, npub):
    inst = random_r1cs(seed, lc, lv, num_public=npub, rest_frac=0.0, width=width)
    rng = np.random.default_rng(seed + 100)
    pts = points(orc, width + 3, seed=21)
    rows = inst["num_vars"] // width
    rnd = dict(blinds=rand_fe(rng, rows), be=rand_fe(rng, 1), dv=rand_fe(rng, width), rd=rand_fe(rng, 1), rb=rand_fe(rng, 1))
    vk = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
    return inst, pts, rnd, vk


@pytest.mark.parametrize("world,lc,lv", [(2, 13, 13), (2, 16, 15), (4, 16, 16)])
def test_sharded_spartan_prove_ranks_on_one_gpu(ctx, orc, ranks, world, lc, lv):
    """sp2_spartan_prove_sharded (rows / transposed columns i = rank mod G, both sum-checks sharded, poly_ABC columns) with
    G ranks on one GPU: every rank's proof == the single-GPU proof == the oracle's proof, and the oracle verifier accepts."""
    import spartan2_b200 as sp
    width = 64
    inst, pts, rnd, vk = _prove_case(orc, 40 + lc, lc, lv, width, 3)
    ck, h, ck_s, h_s = pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3]
    W, X = inst["W"], inst["X"]
    cl = inst["num_shared"] + inst["num_precommitted"]; cr = cl // width
    d = dims(inst)
    O = orc.Shape(*d, inst["A"], inst["B"], inst["C"]); keys = orc.Keys(ck, h, ck_s, h_s)
    orc.set_threads(orc.max_threads())
    comm_pre = orc.hyrax_commit(ck, h, W[:cl], rnd["blinds"][:cr], is_small=False)
    oproof = orc.spartan_prove(O, keys, vk, X, W, comm_pre, orc.Rand(rnd["blinds"], rnd["be"], rnd["dv"], rnd["rd"], rnd["rb"]))
    orc.set_threads(1)
    cs = ranks(world)
    Ks = [sp.CommitmentKey(c, ck, h, ck_s, h_s) for c in cs]
    args = (vk, X, W[cl:], rnd["blinds"], rnd["be"], rnd["dv"], rnd["rd"], rnd["rb"])
    for c, K in zip(cs, Ks):                                # single-GPU proof on every context (also warms its scratch)
        S1 = sp.SplitR1CSShape(c, *d, inst["A"], inst["B"], inst["C"])
        p1 = sp.SpartanSNARK.prep_prove(c, S1, K, W[:cl], rnd["blinds"][:cr], is_small=False)
        single = sp.SpartanSNARK.prove(c, S1, K, p1, *args)
        for f in sp.SpartanProof.FIELDS:
            assert np.array_equal(getattr(single, f).reshape(-1), getattr(oproof, f).reshape(-1)), f
        p1.free(); S1.free()
    comms = sp.Comm.in_process(cs)
    Ss = [sp.SplitR1CSShape(c, *d, inst["A"], inst["B"], inst["C"], rank=r, nranks=world) for r, c in enumerate(cs)]
    preps = [sp.SpartanSNARK.prep_prove(c, S, K, W[:cl], rnd["blinds"][:cr], is_small=False) for c, S, K in zip(cs, Ss, Ks)]
    for rep in range(2):                                    # twice: epochs / flags advance correctly across proves
        proofs = _threads(world, lambda r: sp.SpartanSNARK.prove(cs[r], Ss[r], Ks[r], preps[r], *args, comm=comms[r]))
        for p in proofs:
            for f in sp.SpartanProof.FIELDS:
                assert np.array_equal(getattr(p, f).reshape(-1), getattr(oproof, f).reshape(-1)), (rep, f)
    vp = orc.Proof(proofs[0].l, proofs[0].nry, proofs[0].rows, proofs[0].num_cols)
    for f in sp.SpartanProof.FIELDS:
        getattr(vp, f)[...] = getattr(proofs[-1], f).reshape(getattr(vp, f).shape)
    assert orc.spartan_verify(O, keys, vk, X, vp) == 0
    for x in preps + Ss + comms + Ks:
        x.free()


@pytest.mark.parametrize("n,world", [(4, 2), (8, 4), (16, 2)])
def test_sharded_neutronnova_peer_stores_on_one_gpu(ctx, orc, ranks, n, world):
    """The instance-sharded NeutronNova prove with the in-kernel exchange: per-round sums through the comm mailboxes
    (k_publish_xchg), surviving layers and witness partials by peer stores + flag barriers (k_nn_scatter / k_nn_barrier);
    no allgather callback at all.  Every rank == the single-GPU fused prove (pinned to the oracle elsewhere)."""
    import spartan2_b200 as sp
    from spartan2_b200 import neutronnova as nn
    from tests.neutronnova_ops import sha_chain_instances
    c0, zs, Ws, zc, Wc = sha_chain_instances(n)
    A, B, Cm = c0.matrices()
    cs = ranks(world)
    shapes = [sp.SplitR1CSShape(c, *c0.dims(), A, B, Cm) for c in cs]
    want = None
    for c, S in zip(cs, shapes):                            # single-GPU prove on every context (reference + scratch warm-up)
        single = nn.NeutronNovaProver(c, S, zs, zc)
        v, _ = single.prove(sp.Keccak256Transcript(b"neutronnova_prove"))
        single.free()
        if want is None:
            want = v
        else:
            assert all(np.array_equal(v[k], want[k]) for k in want if isinstance(want[k], np.ndarray))
    nl = n // world
    comms = sp.Comm.in_process(cs)
    provers = [nn.NeutronNovaProver(cs[r], shapes[r], zs[r * nl:(r + 1) * nl], zc, rank=r, nranks=world, comm=comms[r]) for r in range(world)]
    nn.NeutronNovaProver.connect_in_process(provers)
    for rep in range(2):
        outs = _threads(world, lambda r: provers[r].prove(sp.Keccak256Transcript(b"neutronnova_prove"))[0])
        for r in range(world):
            for k, v in want.items():
                if isinstance(v, np.ndarray):
                    assert np.array_equal(outs[r][k], v), (rep, r, k)
            assert outs[r]["outer_ok"] and outs[r]["inner_ok"]
    for x in provers + comms + shapes:
        x.free()


@pytest.mark.parametrize("n,world", [(4, 2), (8, 4)])
def test_sharded_neutronnova_snark_on_one_gpu(ctx, orc, ranks, n, world):
    """The FULL NeutronNova prove, instance-sharded (sp2_neutronnova_snark_prove_sharded): every rank rerandomises / commits only
    its own instances, the rows are all-gathered for the transcript, per-round sums and bulk exchanges go through the peer
    mailboxes — every rank's proof == the single-GPU proof == the oracle's C driver, and the oracle verifier accepts it."""
    import ctypes as C
    import spartan2_b200 as sp
    from spartan2_b200 import neutronnova as nn
    from tests.test_gpu_neutronnova_snark import FIELDS, sha_case
    from tests.test_oracle_neutronnova_snark import prove, step_X
    c = sha_case(orc, ctx, n)
    orc.set_threads(orc.max_threads())
    P = prove(orc, c)
    orc.set_threads(1)
    K = c["keys"]; rows, pre_rows = c["rows"], c["pre_rows"]; nl = n // world
    cs = ranks(world)
    shapes = [sp.SplitR1CSShape(x, *c["dims"], *c["mats"]) for x in cs]
    cks = [sp.CommitmentKey(x, K.ck, K.h, K.ck_s, K.h_s) for x in cs]
    rnd = c["rand"].a
    for x, S, ck in zip(cs, shapes, cks):                    # single-GPU snark on every context: reference + scratch warm-up
        single = nn.NeutronNovaProver(x, S, list(c["zs"]), c["zc"]); single.commit(ck, c["b_old_s"], c["b_old_c"])
        v1, _ = single.snark_prove(c["vk"], *rnd)
        for k in FIELDS:
            assert np.array_equal(np.asarray(v1[k]).reshape(-1), getattr(P, k).reshape(-1)), k
        single.free()
    bar = threading.Barrier(world); slots = [None] * world

    def make_allgather(rank):
        def allgather(send, nbytes, recv, on_device):
            slots[rank] = (send, nbytes)
            bar.wait()
            for q in range(world):
                src, nb = slots[q]
                if on_device:
                    if recv + q * nbytes != src:
                        cs[rank].check(cs[rank].L.sp2_dev_copy(cs[rank].h, C.c_void_p(recv + q * nbytes), C.c_void_p(src), C.c_uint64(nbytes)))
                else:
                    C.memmove(recv + q * nbytes, src, nbytes)
            if on_device:
                cs[rank].synchronize()
            bar.wait()
        return allgather
    comms = sp.Comm.in_process(cs)
    provers = [nn.NeutronNovaProver(cs[r], shapes[r], list(c["zs"][r * nl:(r + 1) * nl]), c["zc"], rank=r, nranks=world, allgather=make_allgather(r), comm=comms[r])
               for r in range(world)]
    nn.NeutronNovaProver.connect_in_process(provers)
    for r, pr in enumerate(provers):
        cs_, cc_ = pr.commit(cks[r], c["b_old_s"][r * nl * pre_rows:(r + 1) * nl * pre_rows], c["b_old_c"])
        assert np.array_equal(cs_, c["comm_pre_s"][r * nl * pre_rows:(r + 1) * nl * pre_rows]) and np.array_equal(cc_, c["comm_pre_c"])

    def run(r):
        try:
            return provers[r].snark_prove(c["vk"], *rnd)[0]
        except Exception:
            bar.abort(); raise
    for rep in range(2):
        outs = _threads(world, run)
        for v in outs:
            for k in FIELDS:
                assert np.array_equal(np.asarray(v[k]).reshape(-1), getattr(P, k).reshape(-1)), (rep, k)
    V = orc.NnProof(n, c["dims"][0], c["M"], c["width"])
    for k in FIELDS:
        getattr(V, k)[...] = np.asarray(outs[-1][k]).reshape(getattr(V, k).shape)
    sx, cx = step_X(c)
    assert orc.neutronnova_verify(c["O"], K, c["vk"], sx, cx, V) == 0
    for x in provers + comms + shapes + cks:
        x.free()


def test_missing_peer_times_out_and_comm_recovers(ctx, orc, ranks):
    """Rank 0 enters a sharded sum-check alone: its kernels' waits on the peer mailbox are bounded (%globaltimer), so the
    call returns InternalError after ~2 s instead of hanging the GPU; after sp2_comm_reset on both ranks the same call
    succeeds and is bit-exact."""
    import spartan2_b200 as sp
    l = 17; world = 2
    rng = np.random.default_rng(5); n = 1 << l
    A, B, Cz, taus = rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, l)
    claim = orc.f_dot_delayed(orc.eq_evals(taus), orc.f_sub(orc.f_mul(A, B), Cz))
    t_or, ts = ts_pair(orc)
    st0 = ts.get()
    want = orc.sumcheck_cubic_prove(claim, taus, A, B, Cz, t_or)[:3]
    cs = ranks(world)
    for c in cs:
        sp.SumcheckProof.prove_cubic_with_three_inputs(c, claim, taus, A, B, Cz, sp.TranscriptState.make(*st0))
    comms = sp.Comm.in_process(cs)
    up = lambda: [[c.upload(sp.shard_cyclic(T, world, r)) for T in (A, B, Cz)] for r, c in enumerate(cs)]   # noqa: E731
    shards = up()
    t0 = time.perf_counter()
    with pytest.raises(sp.SpartanError) as ei:
        comms[0].prove_cubic_with_three_inputs(claim, taus, *shards[0], sp.TranscriptState.make(*st0))
    assert ei.value.kind == "InternalError" and time.perf_counter() - t0 < 30
    cs[0].synchronize()                                     # the GPU is alive
    for cm in comms:
        cm.reset()
    shards = up()

    def cubic(r):
        return comms[r].prove_cubic_with_three_inputs(claim, taus, *shards[r], sp.TranscriptState.make(*st0))
    for polys, rr, claims in _threads(world, cubic):
        assert np.array_equal(polys, want[0]) and np.array_equal(rr, want[1]) and np.array_equal(claims, want[2])
    for cm in comms:
        cm.free()
