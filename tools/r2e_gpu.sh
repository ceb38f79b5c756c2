set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hyrax_family.py tests/test_gpu_neutronnova_snark.py tests/test_gpu_neutronnova.py -m gpu -q 2>&1 | grep -E "^E   |Error|passed|failed|parity|comm_" | head -40 > gpurun_out/r2e.log
cat gpurun_out/r2e.log
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import spartan2_b200 as sp
from spartan2_b200 import neutronnova as nn
from oracle import pyoracle as orc
from tests.test_gpu_neutronnova_snark import sha_case
ctx = sp.Context(0)
for n in (32, 256):
    c = sha_case(orc, ctx, n) if n == 32 else None
    if c is None:
        # 256 steps: skip the oracle commitments (slow); random "old" blinds, commitments recomputed by the device
        from tests.neutronnova_ops import sha_chain_instances
        from tests.gpu_util import rand_fe
        c0, zs, Ws, zc, Wc = sha_chain_instances(n)
        A, B, Cm = c0.matrices(); d = c0.dims(); width = 2048
        pts = ctx.test_points(width + 3, seed=9)
        keys = orc.Keys(pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3])
        M = d[2] + d[3] + d[4]; rows = M // width; pre_rows = d[3] // width
        rng = np.random.default_rng(5)
        c = dict(keys=keys, vk=bytes(32), zs=np.stack(zs), zc=zc, dims=d, mats=(A, B, Cm), b_old_s=rand_fe(rng, n * pre_rows), b_old_c=rand_fe(rng, pre_rows),
                 rand=orc.NnRand(rand_fe(rng, n * rows), rand_fe(rng, rows), rand_fe(rng, 2), rand_fe(rng, width), rand_fe(rng, 1), rand_fe(rng, 1)))
    K = c["keys"]
    S = sp.SplitR1CSShape(ctx, *c["dims"], *c["mats"]); ck = sp.CommitmentKey(ctx, K.ck, K.h, K.ck_s, K.h_s)
    t0 = time.perf_counter(); prover = nn.NeutronNovaProver(ctx, S, list(c["zs"]), c["zc"]); prover.commit(ck, c["b_old_s"], c["b_old_c"]); prep = (time.perf_counter() - t0) * 1e3
    walls, phs = [], []
    for it in range(8):
        t0 = time.perf_counter(); v, ph = prover.snark_prove(c["vk"], *c["rand"].a); w = (time.perf_counter() - t0) * 1e3
        if it >= 3: walls.append(w); phs.append(ph)
    print("n=%d snark_prove %.3f ms (prep %.1f ms) phases %s" % (n, np.mean(walls), prep, {k: round(float(np.mean([p[k] for p in phs])), 3) for k in phs[0]}), flush=True)
    prover.free(); S.free(); ck.free()
PY
