"""Multi-GPU parity + timing of the instance-sharded NeutronNova hot path (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29514 tools/multi_gpu_neutronnova.py [n ...]
For each total instance count n (default 32 256; 256 = BASELINE config 5) every rank builds the SHA-256 chain, keeps its
n/G step instances, runs sp2_neutronnova_prep_prove_sharded + sp2_neutronnova_prove_sharded (local NIFS rounds with one
64-byte all-gather each, NCCL all-gather of the surviving layers and of the witness partials, replicated sum-checks) and
compares EVERY output with rank 0's single-GPU fused prove of all n instances (itself checked against the oracle by
tests/test_gpu_neutronnova.py) and across ranks."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spartan2_b200 as sp  # noqa: E402
from spartan2_b200 import neutronnova as nn  # noqa: E402
from spartan2_b200 import _fq as fq  # noqa: E402
from spartan2_b200.frontend import Sha256Circuit  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = sp.Context(local)
    def allgather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    comm = sp.Comm(ctx, rank, world, allgather_bytes) if os.environ.get("SP2_NN_HOST_XCHG") != "1" else None
    ns = [int(a) for a in sys.argv[1:]] or [32, 256]
    one = fq.from_int(1)
    ok_all = True
    for n in ns:
        if n % world or (n // world) & (n // world - 1):
            continue
        nl = n // world
        mine = range(rank * nl, (rank + 1) * nl)
        need = range(n) if rank == 0 else mine

        def z_of(c):
            W, X = c.witness()
            return np.concatenate([W, one, X], axis=0)
        circs = {i: Sha256Circuit(bytes([i % 256]) * 64, kind="compression") for i in need}
        core = Sha256Circuit(bytes(64), kind="compression")
        c0 = core
        A, B, Cm = c0.matrices()
        S = sp.SplitR1CSShape(ctx, *c0.dims(), A, B, Cm)
        zc = z_of(core)
        t0 = time.perf_counter()
        prover = nn.NeutronNovaProver(ctx, S, [z_of(circs[i]) for i in mine], zc, rank=rank, nranks=world, allgather=nn.torch_allgather(world, dev), comm=comm,
                                      allgather_bytes=allgather_bytes if os.environ.get("SP2_NN_NCCL_GATHER") != "1" else None)
        ctx.synchronize(); prep_ms = (time.perf_counter() - t0) * 1e3
        best = None
        for it in range(6):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            v, ph = prover.prove(sp.Keccak256Transcript(b"neutronnova_prove"))
            wall = (time.perf_counter() - t0) * 1e3
            if it and (best is None or wall < best[0]):
                best = (wall, ph)
        t = torch.tensor([best[0]], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        keys = [k for k in v if isinstance(v[k], np.ndarray)]
        blob = b"".join(np.ascontiguousarray(v[k]).tobytes() for k in keys)
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        same_ranks = all(b == blobs[0] for b in blobs)
        msg = ""
        if rank == 0:
            single = nn.NeutronNovaProver(ctx, S, [z_of(circs[i]) for i in range(n)], zc)
            bs = None
            for it in range(4):
                t0 = time.perf_counter()
                v1, ph1 = single.prove(sp.Keccak256Transcript(b"neutronnova_prove"))
                w1 = (time.perf_counter() - t0) * 1e3
                if it and (bs is None or w1 < bs[0]):
                    bs = (w1, ph1)
            diff = [k for k in keys if not np.array_equal(v[k], v1[k])]
            same_single = not diff and v["outer_ok"] and v["inner_ok"]
            if diff:
                print("    differing outputs: %s (first differing NIFS round: %s)" % (diff, [i for i in range(v["nifs_evals"].shape[0]) if not np.array_equal(v["nifs_evals"][i], v1["nifs_evals"][i])][:1]), flush=True)
            ok_all &= same_ranks and same_single
            msg = ("n=%d instances over %d GPUs (%d per rank): all ranks identical: %s, sharded == single-GPU: %s | prove wall ms single %.3f -> sharded %.3f "
                   "(max over ranks); prep_prove (rank 0) %.1f ms\n    single phases %s\n    sharded phases %s"
                   % (n, world, nl, same_ranks, same_single, bs[0], float(t[0]), prep_ms, {k: round(x, 3) for k, x in bs[1].items()},
                      {k: round(x, 3) for k, x in best[1].items()}))
            single.free()
            print(msg, flush=True)
        prover.free(); S.free()
        dist.barrier()
    if comm is not None:
        comm.free()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
