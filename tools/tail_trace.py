"""debug: per-round phase stamps of k_cubic_tail_pipe (library built with SP2_NVCC_FLAGS=-DSP2_TAIL_TRACE)"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
import spartan2_b200 as sp
from spartan2_b200 import _lib
from tests.gpu_util import rand_fe

ctx = sp.Context(0)
l = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(1)
n = 1 << l
A, B, Cz, taus = rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, n), rand_fe(rng, l)
claim = np.zeros((1, 4), dtype=np.uint64)
for it in range(3):
    ts = sp.TranscriptState.make(bytes(64), 1)
    sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, claim, taus, A, B, Cz, ts)
out = np.zeros((2, 40, 10), dtype=np.int64)
L = ctx.L
L.sp2_debug_tail_trace.restype = C.c_int32
L.sp2_debug_tail_trace.argtypes = [C.c_void_p]
print("rc", L.sp2_debug_tail_trace(out.ctypes.data))
import os
t = out[0 if os.environ.get("SP2_TAIL_PIPE", "1") != "0" else 1]
print("round: role[start bind coef sums sync] fin[start eval msg squeeze]  (cycles rel. to round start)")
for r1 in range(1, l + 1):
    if t[r1, 0] == 0: continue
    b = t[r1, 0]
    print(r1, [int(t[r1, k] - b) for k in range(1, 5)], [int(t[r1, k] - b) for k in range(5, 9)], "round total", int(t[r1, 4] - b))
