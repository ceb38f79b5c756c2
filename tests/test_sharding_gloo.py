"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes.

The CUDA kernels exchange their per-round partial sums through peer mailboxes; what the host side owns is the
SHARDING SCHEME (cyclic split on the low index bits, weights indexed by the global pair id, handle all-gather).
Here two gloo ranks apply that scheme to oracle-computed quantities and must reproduce the full-table results:
  * bind locality:   shard(bind(T, r)) == bind(shard(T), r)          (every bind pair lives on one rank)
  * round sums:      sum over ranks of the local (t(0), t(inf)) == the full-table sums of the first two rounds
  * handle exchange: the all-gather used for the 64-byte IPC handles returns rank-ordered payloads
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2


def _rand_fe(rng, n):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a


def _worker(rank, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from oracle import pyoracle as orc
        from spartan2_b200.host import shard_cyclic
        l = 8; n = 1 << l
        rng = np.random.default_rng(42)                               # same global tables on both ranks
        A, B, Cz, taus, r0 = _rand_fe(rng, n), _rand_fe(rng, n), _rand_fe(rng, n), _rand_fe(rng, l), _rand_fe(rng, 1)
        # 1. handle all-gather is rank ordered
        out = [None] * WORLD
        dist.all_gather_object(out, bytes([rank]) * 64)
        assert out == [bytes([q]) * 64 for q in range(WORLD)]
        # 2. bind locality
        for T in (A, B, Cz):
            assert np.array_equal(shard_cyclic(orc.bind_top(T, r0), WORLD, rank), orc.bind_top(shard_cyclic(T, WORLD, rank), r0))
        # 3. round sums: local sums with weights taken at the GLOBAL index, all-gathered, added mod p
        E = orc.eq_evals(taus[1:])                                    # round-1 weights over the n/2 pairs (tau_0 is the bound variable)
        half = n // 2
        t0e = orc.f_sub(orc.f_mul(A[:half], B[:half]), Cz[:half])
        tie = orc.f_mul(orc.f_sub(A[half:], A[:half]), orc.f_sub(B[half:], B[:half]))
        full = (orc.f_dot_delayed(E, t0e), orc.f_dot_delayed(E, tie))
        sl = slice(rank, None, WORLD)                                 # this rank's pairs = its shard's pairs
        La, Lb, Lc = (shard_cyclic(T, WORLD, rank) for T in (A, B, Cz))
        lh = half // WORLD
        lt0 = orc.f_sub(orc.f_mul(La[:lh], Lb[:lh]), Lc[:lh])
        lti = orc.f_mul(orc.f_sub(La[lh:], La[:lh]), orc.f_sub(Lb[lh:], Lb[:lh]))
        local = (orc.f_dot_delayed(np.ascontiguousarray(E[sl]), lt0), orc.f_dot_delayed(np.ascontiguousarray(E[sl]), lti))
        parts = [None] * WORLD
        dist.all_gather_object(parts, [x.tolist() for x in local])
        tot = [np.zeros((1, 4), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64)]
        for p in parts:
            for k in range(2):
                tot[k] = orc.f_add(tot[k], np.array(p[k], dtype=np.uint64))
        assert np.array_equal(tot[0], full[0]) and np.array_equal(tot[1], full[1])
        # ... and they are the sums the oracle's prover uses in round 1 (claim-consistent instance)
        claim = orc.f_dot_delayed(orc.eq_evals(taus), orc.f_sub(orc.f_mul(A, B), Cz))
        t = orc.Transcript(b"x")
        _, _, _, traw = orc.sumcheck_cubic_prove(claim, taus, A, B, Cz, t)
        assert np.array_equal(traw[0, 0].reshape(1, 4), full[0]) and np.array_equal(traw[0, 1].reshape(1, 4), full[1])
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_cyclic_sharding_scheme_world2():
    mgr = mp.Manager(); ret = mgr.dict()
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(port, ret), nprocs=WORLD, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def _worker_nn(rank, port, ret):
    """NeutronNova instance sharding: rank g holds the contiguous block of instances [g n/G, (g+1) n/G).
    Two gloo ranks run the NIFS on their blocks — per round the local (e0, quad) with suffix weights taken at the GLOBAL
    pair index, all-gathered and added mod p; local folds; then the surviving layers and the witness partials are
    all-gathered — and must reproduce the single-process oracle values."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from oracle import pyoracle as orc
        n, left, right = 8, 8, 4
        N = left * right; ell_b = 3; nl = n // WORLD; M = 16
        rng = np.random.default_rng(7)                                # same global data on both ranks
        A, B, Cm = _rand_fe(rng, n * N), _rand_fe(rng, n * N), _rand_fe(rng, n * N)
        Ws = _rand_fe(rng, n * M)
        E = orc.pow_split_evals(_rand_fe(rng, 1), left, right); rhos = _rand_fe(rng, ell_b); r_bs = _rand_fe(rng, ell_b)
        blk = slice(rank * nl * N, (rank + 1) * nl * N)
        lA, lB, lC = A[blk].copy(), B[blk].copy(), Cm[blk].copy()
        fA, fB, fC, m, ml = A, B, Cm, n, nl
        for t in range(ell_b):
            want = orc.nifs_round(t, rhos, left, right, E, fA, fB, fC, N, m)
            if ml >= 2:                                               # local round
                mine = orc.nifs_round(t, rhos, left, right, E, lA, lB, lC, N, ml, pair_offset=rank * (ml // 2))
                parts = [None] * WORLD
                dist.all_gather_object(parts, mine.tolist())
                got = np.zeros((2, 4), dtype=np.uint64)
                for p in parts:
                    got = orc.f_add(got, np.array(p, dtype=np.uint64))
                lA, lB, lC = (orc.nifs_fold(x, N, ml, r_bs[t:t + 1]) for x in (lA, lB, lC))
                ml //= 2
                if ml == 1:                                           # hand-off: gather the surviving layers in rank order
                    gath = [None] * WORLD
                    dist.all_gather_object(gath, [lA.tolist(), lB.tolist(), lC.tolist()])
                    lA, lB, lC = (np.concatenate([np.array(g[k], dtype=np.uint64) for g in gath], axis=0) for k in range(3))
                    ml = 0; mrep = WORLD
            else:                                                     # replicated rounds on the gathered layers
                got = orc.nifs_round(t, rhos, left, right, E, lA, lB, lC, N, mrep)
                lA, lB, lC = (orc.nifs_fold(x, N, mrep, r_bs[t:t + 1]) for x in (lA, lB, lC))
                mrep //= 2
            assert np.array_equal(got, want), t
            fA, fB, fC = (orc.nifs_fold(x, N, m, r_bs[t:t + 1]) for x in (fA, fB, fC))
            m //= 2
        assert np.array_equal(lA, fA) and np.array_equal(lB, fB) and np.array_equal(lC, fC)
        # witness fold: local partial with the GLOBAL weights of this rank's instances, gathered, added
        w = orc.weights_from_r(r_bs, n)
        part = orc.fold_vectors(Ws[rank * nl * M:(rank + 1) * nl * M], nl, M, np.ascontiguousarray(w[rank * nl:(rank + 1) * nl]))
        parts = [None] * WORLD
        dist.all_gather_object(parts, part.tolist())
        tot = np.zeros((M, 4), dtype=np.uint64)
        for p in parts:
            tot = orc.f_add(tot, np.array(p, dtype=np.uint64))
        assert np.array_equal(tot, orc.fold_vectors(Ws, n, M, w))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_neutronnova_instance_sharding_world2():
    mgr = mp.Manager(); ret = mgr.dict()
    port = 29950 + (os.getpid() % 300)
    mp.spawn(_worker_nn, args=(port, ret), nprocs=WORLD, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
