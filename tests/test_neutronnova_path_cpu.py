"""spartan2_b200.neutronnova.run (HOT LOOPS A-C of NeutronNovaZkSNARK::prove with transcript-derived challenges) over the
ORACLE backend on the CPU: checks the driver's host algebra — the final verifier equations of both batched sum-checks
hold, i.e. the NIFS target, the claims and the round polynomials are mutually consistent — without a GPU."""
import numpy as np

from tests.neutronnova_ops import OracleOps, sha_chain_instances


def test_hot_path_consistent_on_oracle(orc):
    from spartan2_b200 import neutronnova as nn
    c0, zs, Ws, zc, Wc = sha_chain_instances(2)
    assert c0.num_cons_unpadded in (25840, 25840 + 512)                      # benches/sha256_neutronnova.rs:159-160
    A, B, Cm = c0.matrices()
    ops = OracleOps(orc.Shape(*c0.dims(), A, B, Cm), c0.dims())
    ts = orc.Transcript(b"neutronnova_prove")
    trace = []
    out = nn.run(ops, ts, c0.num_cons, zs, Ws, zc, Wc, trace=trace)
    assert out["outer_ok"] and out["inner_ok"]
    assert nn.compute_tensor_decomp(c0.num_cons) == (15, 256, 128)
    names = [t[0] for t in trace]
    assert "nifs_round_0" in names and "outer_polys_14" in names and "inner_polys_15" in names


def test_host_transcript_matches_oracle(orc):
    # the library's host Keccak256Transcript (pure host code in the .so) against the oracle's, incl. from_uniform
    import spartan2_b200 as sp
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2**64, size=(5, 4), dtype=np.uint64); a[:, 3] &= np.uint64(0x7fffffffffffffff)
    t1, t2 = sp.Keccak256Transcript(b"tst"), orc.Transcript(b"tst")
    for t in (t1, t2):
        t.absorb_bytes(b"vk", bytes(range(32)))
        t.absorb_scalars(b"x", a)
        t.dom_sep(b"sep")
    assert np.array_equal(t1.squeeze(b"c"), t2.squeeze(b"c"))
    for t in (t1, t2):
        t.absorb_scalars(b"p", a[:3])
    assert np.array_equal(t1.squeeze(b"c"), t2.squeeze(b"c"))
    st, rnd = t2.state()
    assert t1.state().get() == (st, rnd)
