"""debug: per-round phase stamps of k_cubic_mid_pipe on CTA 0 (library built with SP2_NVCC_FLAGS=-DSP2_TAIL_TRACE)"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
import spartan2_b200 as sp
from tests.gpu_util import rand_fe
ctx = sp.Context(0)
l = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(1)
n = 1 << l
A, B, Cz, taus = rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, l)
claim = np.zeros((1, 4), dtype=np.uint64)
dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cz)
for it in range(3):
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cz)
    ts = sp.TranscriptState.make(bytes(64), 1)
    sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, claim, taus, dA, dB, dC, ts)
out = np.zeros((2, 40, 10), dtype=np.int64)
L = ctx.L
L.sp2_debug_tail_trace.restype = C.c_int32
L.sp2_debug_tail_trace.argtypes = [C.c_void_p]
print("rc", L.sp2_debug_tail_trace(out.ctypes.data))
t = out[1]
base = None
print("round: role[stage start, r seen, bound, direct published, coef done, arrived] fin[start, before squeeze, released]  (us, %globaltimer, rel. to the first stamp)")
for r1 in range(1, l + 1):
    if t[r1, 0] == 0: continue
    if base is None: base = t[r1, 0]
    print(r1, ["%.1f" % ((t[r1, k] - base) / 1000.0) if t[r1, k] else "-" for k in range(0, 6)], ["%.1f" % ((t[r1, k] - base) / 1000.0) if t[r1, k] else "-" for k in range(6, 9)])
t0 = out[0]
print("fin_cubic_msg (ns): load, eval, products, from_mont")
for r1 in range(1, l + 1):
    if t0[r1, 0] == 0: continue
    print(r1, [int(t0[r1, k] - t0[r1, k - 1]) for k in range(1, 5)])
