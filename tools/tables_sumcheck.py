"""The reference's pure-table sum-check benchmark (src/sumcheck.rs:1450-1553) on device-resident uniform random tables:
python tools/tables_sumcheck.py [num_vars=24] [reps=3] — prints total and persistent-kernel times; used under ncu to
profile the streaming rounds of k_cubic_persist / k_quad_persist in isolation."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan2_b200 as sp  # noqa: E402

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = sp.Context(0)
rng = np.random.default_rng(0xDEADBEEF)
n = 1 << nv; chunk = min(n, 1 << 20)


def rnd(k):
    a = rng.integers(0, 2**64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a


small = [ctx.upload(rnd(chunk)) for _ in range(3)]
tabs = [ctx.alloc(n * 32) for _ in range(3)]
taus = rnd(nv); zero = np.zeros((1, 4), dtype=np.uint64)
for kind in ("cubic", "quad"):
    for it in range(reps):
        for t, sbuf in zip(tabs, small):
            for i in range(n // chunk):
                ctx.check(ctx.L.sp2_dev_copy(ctx.h, C.c_void_p(t.ptr.value + i * chunk * 32), sbuf.ptr, C.c_uint64(chunk * 32)))
        ctx.synchronize()
        ts = sp.TranscriptState()
        ctx.timer_start()
        if kind == "cubic":
            sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, zero, taus, tabs[0], tabs[1], tabs[2], ts)
        else:
            sp.SumcheckProof.prove_quad(ctx, zero, nv, tabs[0], tabs[1], ts)
        ms = ctx.timer_stop()
        k = C.c_float(0)
        if kind == "cubic" and ctx.L.sp2_last_cubic_persist_ms(ctx.h, C.byref(k)) != 0:
            k = C.c_float(float("nan"))          # SP2_NO_PERSIST=1: one launch per round, no persistent kernel
        bytes_alg = (368 if kind == "cubic" else 256) * n
        print("%s 2^%d: total %.3f ms (%.0f GB/s algorithmic)%s" % (kind, nv, ms, bytes_alg / ms / 1e6, "; k_cubic_persist %.3f ms" % k.value if kind == "cubic" else ""), flush=True)
