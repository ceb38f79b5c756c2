"""Shared helpers for the -m gpu parity tests: the CUDA path is called through the C ABI
(spartan2_b200 package), the oracle (oracle/) is the checker."""
import numpy as np
import pytest

Q = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff


def rand_fe(rng, n, small=False):
    """n uniformly random canonical Montgomery elements of the T256 scalar field as (n,4) u64."""
    if small:
        # small integers in Montgomery form are full-width, so produce them through the oracle
        raise NotImplementedError
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x7fffffffffffffff)   # < 2^255 < p
    return a


@pytest.fixture(scope="session")
def ctx():
    import spartan2_b200 as sp
    c = sp.Context(0)
    yield c
    c.close()


def ts_pair(orc, label=b"test"):
    """(oracle transcript, product TranscriptState) in the same state, after one squeeze."""
    import spartan2_b200 as sp
    t = orc.Transcript(label)
    t.absorb_bytes(b"seed", b"\x01\x02\x03")
    t.squeeze(b"s")
    st, rnd = t.state()
    return t, sp.TranscriptState.make(st, rnd)
