"""Multi-GPU parity + timing of the sharded sum-checks (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Every rank builds the same seeded global tables, keeps its cyclic shard, proves with the peer-mailbox exchange and
compares with the oracle on the full tables (small sizes) / with the single-GPU prover (large sizes, rank 0)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spartan2_b200 as sp  # noqa: E402


def rand_fe(rng, n):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sp.Context(local)

    def allgather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    comm = sp.Comm(ctx, rank, world, allgather_bytes)
    from oracle import pyoracle as orc
    ok = True
    sizes = [int(x) for x in os.environ.get("SP2_MGPU_SIZES", "6,12,17,18,20,22").split(",")]
    for l in sizes:
        rng = np.random.default_rng(1000 + l)                     # same tables on every rank
        n = 1 << l
        A, B, Cz, taus = rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, l)
        zero = np.zeros((1, 4), dtype=np.uint64)
        if l <= 18:      # true claims (the oracle, like the reference, derives t(1) from the claim)
            claim_c = orc.f_dot_delayed(orc.eq_evals(taus), orc.f_sub(orc.f_mul(A, B), Cz)); claim_q = orc.f_dot_delayed(A, B)
        else:
            claim_c = claim_q = zero
        t = orc.Transcript(b"mgpu"); t.squeeze(b"s"); st, rd = t.state()
        # sharded
        dA, dB, dC = (ctx.upload(sp.shard_cyclic(x, world, rank)) for x in (A, B, Cz))
        ts = sp.TranscriptState.make(st, rd)
        dist.barrier(); ctx.synchronize(); t0 = time.perf_counter()
        polys, r, claims = comm.prove_cubic_with_three_inputs(claim_c, taus, dA, dB, dC, ts)
        dt_c = (time.perf_counter() - t0) * 1e3
        dA2, dB2 = ctx.upload(sp.shard_cyclic(A, world, rank)), ctx.upload(sp.shard_cyclic(B, world, rank))
        tsq = sp.TranscriptState.make(st, rd)
        dist.barrier(); ctx.synchronize(); t0 = time.perf_counter()
        qpolys, qr, qclaims = comm.prove_quad(claim_q, l, dA2, dB2, tsq)
        dt_q = (time.perf_counter() - t0) * 1e3
        # reference result: oracle up to 2^18, the single-GPU CUDA prover (itself oracle-checked by tests/) above
        if l <= 18:
            opolys, orr, oclaims, _ = orc.sumcheck_cubic_prove(claim_c, taus, A, B, Cz, t)
            t2 = orc.Transcript(b"mgpu"); t2.squeeze(b"s")
            oq, oqr, oqc = orc.sumcheck_quad_prove(claim_q, l, A, B, t2)
            kind = "oracle"
        else:
            fA, fB, fC = ctx.upload(A), ctx.upload(B), ctx.upload(Cz)
            ts1 = sp.TranscriptState.make(st, rd)
            ctx.synchronize(); t0 = time.perf_counter()
            opolys, orr, oclaims = sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, zero, taus, fA, fB, fC, ts1)
            dt1 = (time.perf_counter() - t0) * 1e3
            fA, fB = ctx.upload(A), ctx.upload(B)
            ts2 = sp.TranscriptState.make(st, rd)
            ctx.synchronize(); t0 = time.perf_counter()
            oq, oqr, oqc = sp.SumcheckProof.prove_quad(ctx, zero, l, fA, fB, ts2)
            dt2 = (time.perf_counter() - t0) * 1e3
            kind = "single-GPU (cubic %.2f ms, quad %.2f ms)" % (dt1, dt2)
        good = (np.array_equal(polys, opolys) and np.array_equal(r, orr) and np.array_equal(claims, oclaims)
                and np.array_equal(qpolys, oq) and np.array_equal(qr, oqr) and np.array_equal(qclaims, oqc))
        flag = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        if rank == 0:
            print("l=%d ranks=%d: sharded == %s: %s | cubic %.2f ms, quad %.2f ms (wall, incl. result download)" % (l, world, kind, bool(flag.item()), dt_c, dt_q), flush=True)
    comm.free(); ctx.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
