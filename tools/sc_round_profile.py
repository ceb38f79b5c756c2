"""Per-round timeline of the persistent cubic sum-check kernel at 2^l (debug instrumentation, %globaltimer on CTA 0)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan2_b200 as sp
ctx = sp.Context(0)
rng = np.random.default_rng(1)
def rnd(k):
    a = rng.integers(0, 2**64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a
l = int(sys.argv[1]) if len(sys.argv) > 1 else 20
A, B, Cc, taus = rnd(1 << l), rnd(1 << l), rnd(1 << l), rnd(l)
zero = np.zeros((1, 4), dtype=np.uint64)
for rep in range(3):
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cc)
    ts = sp.TranscriptState(); ctx.synchronize(); ctx.timer_start()
    sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, zero, taus, dA, dB, dC, ts)
    ms = ctx.timer_stop()
out = np.zeros((l, 4), dtype=np.uint64)
ctx.check(ctx.L.sp2_debug_sc_round_profile(ctx.h, out.ctypes.data_as(C.c_void_p), C.c_uint32(l)))
print("cubic 2^%d: %.1f us total (device-resident tables)" % (l, ms * 1e3))
t0 = int(out[0, 0])
for i in range(l):
    if out[i, 0] == 0: break
    s, c, g, f = [int(x) for x in out[i]]
    nxt = int(out[i + 1, 0]) if i + 1 < l and out[i + 1, 0] else f
    if c == 0: break
    print("round %2d: start +%7.1f us | own compute %6.1f | wait for all CTAs %6.1f | finalize %6.1f | release->next start %5.1f" % (i + 1, (s - t0) / 1e3, (c - s) / 1e3, (g - c) / 1e3, (f - g) / 1e3, (nxt - f) / 1e3))
