"""BASELINE config 3 / 5 shape on one GPU: HOT LOOPS A-C of NeutronNovaZkSNARK::prove on the SHA-256 chain through the
per-round device seams (spartan2_b200.neutronnova), phase timings, next to the same driver over the oracle's CPU
functions (all host threads) for the same phases.  Usage: python tools/neutronnova_path.py [n_steps=32] [--no-cpu]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan2_b200 as sp  # noqa: E402
from spartan2_b200 import neutronnova as nn  # noqa: E402
from tests.neutronnova_ops import OracleOps, sha_chain_instances  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
t0 = time.time()
c0, zs, Ws, zc, Wc = sha_chain_instances(n)
A, B, Cm = c0.matrices()
print("%d step circuits + core: %d constraints each (N = M = 2^%d), built in %.1f s" % (n, c0.num_cons_unpadded, c0.num_cons.bit_length() - 1, time.time() - t0), flush=True)
ctx = sp.Context(0)
S = sp.SplitR1CSShape(ctx, *c0.dims(), A, B, Cm)
ops = nn.DeviceOps(ctx, S)
best = None
for it in range(4):
    tm = {}
    l0 = ctx.launch_count(); t0 = time.perf_counter()
    out = nn.run(ops, sp.Keccak256Transcript(b"neutronnova_prove"), c0.num_cons, zs, Ws, zc, Wc, timing=tm)
    wall = (time.perf_counter() - t0) * 1e3
    assert out["outer_ok"] and out["inner_ok"]
    if it and (best is None or wall < best[0]):
        best = (wall, tm, ctx.launch_count() - l0)
print("CUDA  (B200, per-round seams, %d launches): total %.2f ms; %s" % (best[2], best[0], {k: round(v, 2) for k, v in best[1].items()}), flush=True)
# the fused path of the library: round loop + scalar algebra + transcript in C++, tables device-resident
t0 = time.perf_counter()
prover = nn.NeutronNovaProver(ctx, S, zs, zc)
prover.free()                                   # first call pays the lazy module load of its kernels; time the second
t0 = time.perf_counter()
prover = nn.NeutronNovaProver(ctx, S, zs, zc)
prep_ms = (time.perf_counter() - t0) * 1e3
bestf = None
for it in range(6):
    l0 = ctx.launch_count(); t0 = time.perf_counter()
    v, ph = prover.prove(sp.Keccak256Transcript(b"neutronnova_prove"))
    wall = (time.perf_counter() - t0) * 1e3
    assert v["outer_ok"] and v["inner_ok"]
    if it and (bestf is None or wall < bestf[0]):
        bestf = (wall, ph, ctx.launch_count() - l0)
from spartan2_b200 import _fq as fq  # noqa: E402
assert fq.to_int(v["T_out"]) == out["T_out"] and fq.to_ints(v["eval_W"]) == [out["eval_W_step"], out["eval_W_core"]]
print("CUDA  (B200, fused sp2_neutronnova_prove, %d launches): prove %.3f ms (prep_prove incl. upload, %d SpMVs, i64 layers: %.2f ms); %s"
      % (bestf[2], bestf[0], n + 1, prep_ms, {k: round(x, 3) for k, x in bestf[1].items()}), flush=True)
if "--no-cpu" not in sys.argv:
    from oracle import pyoracle as orc
    orc.lib(native=True); orc.set_threads(orc.max_threads())
    tm = {}
    t0 = time.perf_counter()
    out_o = nn.run(OracleOps(orc.Shape(*c0.dims(), A, B, Cm), c0.dims()), orc.Transcript(b"neutronnova_prove"), c0.num_cons, zs, Ws, zc, Wc, timing=tm)
    wall = (time.perf_counter() - t0) * 1e3
    same = all(out[k] == out_o[k] for k in ("eval_W_step", "eval_W_core", "T_out"))
    print("oracle (CPU port, %d threads): total %.1f ms; %s | final values identical: %s" % (orc.max_threads(), wall, {k: round(v, 1) for k, v in tm.items()}, same), flush=True)
    sys.exit(0 if same else 1)
