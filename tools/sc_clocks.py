import sys, os, ctypes as C, numpy as np
sys.path.insert(0, '.')
import spartan2_b200 as sp
ctx = sp.Context(0)
rng = np.random.default_rng(1)
def rnd(k):
    a = rng.integers(0, 2**64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a
for l in (10, 14, 16):
    A, B, Cc, taus = rnd(1 << l), rnd(1 << l), rnd(1 << l), rnd(l)
    zero = np.zeros((1, 4), dtype=np.uint64)
    for kind in ("cubic", "quad"):
        for rep in range(3):
            ts = sp.TranscriptState()
            ctx.timer_start()
            if kind == "cubic":
                sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, zero, taus, A, B, Cc, ts)
            else:
                sp.SumcheckProof.prove_quad(ctx, zero, l, A, B, ts)
            ms = ctx.timer_stop()
        out = np.zeros(13, dtype=np.uint64)
        ctx.check(ctx.L.sp2_debug_sc_clocks(ctx.h, out.ctypes.data_as(C.c_void_p)))
        d = [int(out[i + 1]) - int(out[i]) for i in range(6)]
        print(kind, "l=%d" % l, "total %.1f us (%.1f us/round)" % (ms * 1e3, ms * 1e3 / l), "cycles: round-body %d, scalar %d, build-msg %d, keccak %d, from_uniform %d, post %d" % tuple(d), "| previous tail round period %d cycles" % (int(out[0]) - int(out[11])), "| last multi-CTA cubic round: entry->election %.1f us, finalize %.1f us, gap after the previous round's end %.1f us" % ((int(out[8]) - int(out[10])) / 1e3, (int(out[9]) - int(out[8])) / 1e3, (int(out[10]) - int(out[12])) / 1e3))
