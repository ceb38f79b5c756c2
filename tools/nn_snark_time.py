"""wall clock and phases of the full NeutronNova prove (sp2_neutronnova_snark_prove) for n step circuits (32: BASELINE config 3; 256: config 5
on one GPU); SP2_NN_SIDE_CTAS caps the side-stream MSM of the folded rows."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan2_b200 as sp
from spartan2_b200 import neutronnova as nn
from oracle import pyoracle as orc
from tests.neutronnova_ops import sha_chain_instances
from tests.gpu_util import rand_fe
ctx = sp.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [32]:
    c0, zs, Ws, zc, Wc = sha_chain_instances(n)
    A, B, Cm = c0.matrices(); d = c0.dims(); width = 2048
    pts = ctx.test_points(width + 3, seed=9)
    keys = orc.Keys(pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3])
    M = d[2] + d[3] + d[4]; rows = M // width; pre_rows = d[3] // width
    rng = np.random.default_rng(5)
    rand = orc.NnRand(rand_fe(rng, n * rows), rand_fe(rng, rows), rand_fe(rng, 2), rand_fe(rng, width), rand_fe(rng, 1), rand_fe(rng, 1))
    S = sp.SplitR1CSShape(ctx, *d, A, B, Cm); ck = sp.CommitmentKey(ctx, keys.ck, keys.h, keys.ck_s, keys.h_s)
    t0 = time.perf_counter(); prover = nn.NeutronNovaProver(ctx, S, list(np.stack(zs)), zc); prover.commit(ck, rand_fe(rng, n * pre_rows), rand_fe(rng, pre_rows)); prep = (time.perf_counter() - t0) * 1e3
    walls, phs = [], []
    for it in range(10):
        t0 = time.perf_counter(); v, ph = prover.snark_prove(bytes(32), *rand.a); w = (time.perf_counter() - t0) * 1e3
        if it >= 3: walls.append(w); phs.append(ph)
    print("SP2_NN_SIDE_CTAS=%s n=%d snark_prove %.3f ms (prep %.1f ms) phases %s" % (os.environ.get("SP2_NN_SIDE_CTAS", "default"), n, np.mean(walls), prep,
          {k: round(float(np.mean([p[k] for p in phs])), 3) for k in phs[0]}), flush=True)
    prover.free(); S.free(); ck.free()
