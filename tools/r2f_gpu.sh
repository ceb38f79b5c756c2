set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_verifier.py tests/test_gpu_neutronnova_snark.py tests/test_gpu_neutronnova.py tests/test_gpu_msm.py -m gpu -q 2>&1 | grep -E "^E   |Error|passed|failed|parity|comm_" | head -40 > gpurun_out/r2f.log
cat gpurun_out/r2f.log
for cap in 48 0 16; do
SP2_NN_SIDE_CTAS=$cap python - <<'PY'
import sys, time, os, numpy as np
sys.path.insert(0, '.')
import spartan2_b200 as sp
from spartan2_b200 import neutronnova as nn
from oracle import pyoracle as orc
from tests.neutronnova_ops import sha_chain_instances
from tests.gpu_util import rand_fe
ctx = sp.Context(0)
for n in (32, 256):
    c0, zs, Ws, zc, Wc = sha_chain_instances(n)
    A, B, Cm = c0.matrices(); d = c0.dims(); width = 2048
    pts = ctx.test_points(width + 3, seed=9)
    M = d[2] + d[3] + d[4]; rows = M // width; pre_rows = d[3] // width
    rng = np.random.default_rng(5)
    S = sp.SplitR1CSShape(ctx, *d, A, B, Cm); ck = sp.CommitmentKey(ctx, pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:width + 3])
    prover = nn.NeutronNovaProver(ctx, S, list(zs), zc); prover.commit(ck, rand_fe(rng, n * pre_rows), rand_fe(rng, pre_rows))
    rnd = (rand_fe(rng, n * rows), rand_fe(rng, rows), rand_fe(rng, 2), rand_fe(rng, width), rand_fe(rng, 1), rand_fe(rng, 1))
    walls, phs = [], []
    for it in range(8):
        t0 = time.perf_counter(); v, ph = prover.snark_prove(bytes(32), *rnd); w = (time.perf_counter() - t0) * 1e3
        if it >= 3: walls.append(w); phs.append(ph)
    hl = []
    for it in range(6):
        t0 = time.perf_counter(); prover.prove(sp.Keccak256Transcript(b"neutronnova_prove")); hl.append((time.perf_counter() - t0) * 1e3)
    print("cap=%s n=%d snark_prove %.3f ms (hot loops only %.3f) phases %s" % (os.environ.get("SP2_NN_SIDE_CTAS"), n, np.mean(walls), np.mean(hl[2:]), {k: round(float(np.mean([p[k] for p in phs])), 3) for k in phs[0]}), flush=True)
    prover.free(); S.free(); ck.free()
PY
done
