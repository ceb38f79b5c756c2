"""Multi-GPU parity + timing of the hypercube-sharded SpartanSNARK::prove (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_prove.py
For each message length (SP2_MGPU_MSGS, bytes; 8192 = BASELINE config 4, N = M = 2^22) every rank builds the same SHA-256
circuit, uploads its shard of the shape (rows / transposed columns i = rank mod N), runs prep_prove + the sharded prove
(sp2_spartan_prove_sharded: Az/Bz/Cz, both sum-checks and poly_ABC sharded, per-round sums exchanged inside the round
kernels over NVLink) and compares every proof field with rank 0's single-GPU proof of the same inputs (itself checked
bit-for-bit against the oracle by tests/test_gpu_spartan.py); for short messages the oracle verifier checks the proof."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spartan2_b200 as sp  # noqa: E402
from spartan2_b200.frontend import Sha256Circuit  # noqa: E402

WIDTH = 2048


def rnd(rng, k):
    a = rng.integers(0, 2**64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a


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
    pts = ctx.test_points(WIDTH + 3, seed=5)
    K = sp.CommitmentKey(ctx, pts[:WIDTH], pts[WIDTH:WIDTH + 1], pts[WIDTH + 1:WIDTH + 2], pts[WIDTH + 2:])
    flush = ctx.alloc(512 << 20)
    ok = True
    reps = int(os.environ.get("SP2_MGPU_REPS", "5"))
    for msg_len in [int(x) for x in os.environ.get("SP2_MGPU_MSGS", "64,2048,8192").split(",")]:
        circ = Sha256Circuit(b"\x00" * msg_len, width=WIDTH)
        A, B, Cm = circ.matrices(); W, X = circ.witness()
        rows = circ.num_vars // WIDTH; cl = circ.num_precommitted; cr = cl // WIDTH
        rng = np.random.default_rng(4)                                  # same randomness on every rank
        blinds, be, dv, rd, rb = rnd(rng, rows), rnd(rng, 1), rnd(rng, WIDTH), rnd(rng, 1), rnd(rng, 1)
        vk = bytes(32)

        def timed(shape, prep, cm):
            dev, wall = [], []
            for i in range(reps + 2):
                ctx.check(ctx.L.sp2_dev_memset(ctx.h, flush.ptr, 0, 512 << 20)); ctx.synchronize()
                if cm is not None:
                    dist.barrier()
                t0 = time.perf_counter()
                p = sp.SpartanSNARK.prove(ctx, shape, K, prep, vk, X, None, blinds, be, dv, rd, rb, comm=cm)
                w = (time.perf_counter() - t0) * 1e3
                if i >= 2:
                    dev.append(p.phase_ms["total"]); wall.append(w)
            return p, float(np.mean(dev)), float(np.mean(wall))

        # sharded: every rank
        Ss = sp.SplitR1CSShape(ctx, *circ.dims(), A, B, Cm, rank=rank, nranks=world)
        preps = sp.SpartanSNARK.prep_prove(ctx, Ss, K, W[:cl], blinds[:cr], is_small=True)
        ps, dev_s, wall_s = timed(Ss, preps, comm)
        t = torch.tensor([dev_s, wall_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                        # max over ranks
        dev_s, wall_s = float(t[0]), float(t[1])
        preps.free(); Ss.free()
        # every rank must hold the same proof: compare against rank 0's bytes
        blob = b"".join(getattr(ps, f).tobytes() for f in sp.SpartanProof.FIELDS)
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        same = all(b == blobs[0] for b in blobs)
        good = same
        msg = ""
        if rank == 0:
            S1 = sp.SplitR1CSShape(ctx, *circ.dims(), A, B, Cm)
            prep1 = sp.SpartanSNARK.prep_prove(ctx, S1, K, W[:cl], blinds[:cr], is_small=True)
            p1, dev_1, wall_1 = timed(S1, prep1, None)
            eq = all(np.array_equal(getattr(ps, f), getattr(p1, f)) for f in sp.SpartanProof.FIELDS)
            good = good and eq
            ver = ""
            if msg_len <= 2048:
                from oracle import pyoracle as orc
                orc.set_threads(orc.max_threads())
                O = orc.Shape(*circ.dims(), A, B, Cm); keys = orc.Keys(pts[:WIDTH], pts[WIDTH:WIDTH + 1], pts[WIDTH + 1:WIDTH + 2], pts[WIDTH + 2:])
                vp = orc.Proof(ps.l, ps.nry, ps.rows, ps.num_cols)
                for f in sp.SpartanProof.FIELDS:
                    getattr(vp, f)[...] = getattr(ps, f).reshape(getattr(vp, f).shape)
                rc = orc.spartan_verify(O, keys, vk, X, vp)
                good = good and rc == 0
                ver = " | oracle verifier: %s" % ("ACCEPT" if rc == 0 else "REJECT %d" % rc)
            msg = ("msg %d B (N=2^%d, M=2^%d) ranks=%d: all ranks identical: %s, sharded == single-GPU proof: %s%s | prove device ms single %.3f -> sharded %.3f "
                   "(wall %.3f -> %.3f)\n    single phases %s\n    sharded phases %s"
                   % (msg_len, ps.l, ps.nry - 1, world, same, eq, ver, dev_1, dev_s, wall_1, wall_s,
                      {k: round(v, 3) for k, v in p1.phase_ms.items()}, {k: round(v, 3) for k, v in ps.phase_ms.items()}))
            prep1.free(); S1.free()
        flag = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        if rank == 0:
            print(msg, flush=True)
    comm.free(); ctx.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
