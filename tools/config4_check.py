"""BASELINE config 4 shape on ONE GPU: SHA-256 circuit of an 8 KiB message (129 compressions, N = M = 2^22), prep_prove +
prove on the device, proof checked by the oracle verifier.  Prints sizes and timings."""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan2_b200 as sp
from spartan2_b200.frontend import Sha256Circuit
from oracle import pyoracle as orc

msg_len = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
t0 = time.time(); circ = Sha256Circuit(b"\x00" * msg_len); print("circuit: %d constraints, N=%d M=%d nnz=%d (%.1fs)" % (circ.num_cons_unpadded, circ.num_cons, circ.num_vars, sum(circ.nnz), time.time() - t0), flush=True)
assert circ.digest == hashlib.sha256(b"\x00" * msg_len).digest()
ctx = sp.Context(0)
width = 2048
pts = ctx.test_points(width + 3, seed=5)
A, B, Cm = circ.matrices(); W, X = circ.witness()
rows = circ.num_vars // width; cl = circ.num_precommitted; cr = cl // width
rng = np.random.default_rng(4)
def rnd(k):
    a = rng.integers(0, 2**64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff); return a
blinds, be, dv, rd, rb = rnd(rows), rnd(1), rnd(width), rnd(1), rnd(1)
vk = bytes(32)
t0 = time.time(); S = sp.SplitR1CSShape(ctx, *circ.dims(), A, B, Cm); K = sp.CommitmentKey(ctx, pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:]); print("setup (shape + key upload) %.2fs" % (time.time() - t0), flush=True)
t0 = time.time(); prep = sp.SpartanSNARK.prep_prove(ctx, S, K, W[:cl], blinds[:cr], is_small=True); print("prep_prove %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
for _ in range(3):
    t0 = time.time(); proof = sp.SpartanSNARK.prove(ctx, S, K, prep, vk, X, None, blinds, be, dv, rd, rb); wall = (time.time() - t0) * 1e3
print("prove: device %.2f ms, wall %.2f ms; phases %s" % (proof.phase_ms["total"], wall, {k: round(v, 3) for k, v in proof.phase_ms.items()}), flush=True)
O = orc.Shape(*circ.dims(), A, B, Cm); keys = orc.Keys(pts[:width], pts[width:width + 1], pts[width + 1:width + 2], pts[width + 2:])
orc.set_threads(orc.max_threads())
vp = orc.Proof(proof.l, proof.nry, proof.rows, proof.num_cols)
for f in sp.SpartanProof.FIELDS:
    getattr(vp, f)[...] = getattr(proof, f).reshape(getattr(vp, f).shape)
t0 = time.time(); rc = orc.spartan_verify(O, keys, vk, X, vp); print("oracle verifier: %s (%.1fs)" % ("ACCEPT" if rc == 0 else "REJECT %d" % rc, time.time() - t0))
sys.exit(0 if rc == 0 else 1)
