"""Device field layer vs the oracle (and Python big ints): reference src/big_num property tests
(delayed_reduction.rs:70-94, montgomery.rs:188-229) re-run on the GPU through the C ABI."""
import ctypes as C
import random

import numpy as np
import pytest

from tests.gpu_util import Q, ctx, rand_fe  # noqa: F401

pytestmark = pytest.mark.gpu
P_BASE = 0xffffffff0000000100000000000000017e72b42b30e7317793135661b1c4b117
R = 1 << 256


def _op(ctx, field, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64); out = np.zeros_like(a)
    bp = None if b is None else np.ascontiguousarray(b, dtype=np.uint64).ctypes.data_as(C.c_void_p)
    ctx.check(ctx.L.sp2_test_field_op(ctx.h, C.c_int32(field), C.c_int32(op), a.ctypes.data_as(C.c_void_p), bp,
                                      out.ctypes.data_as(C.c_void_p), C.c_uint64(a.shape[0])))
    return out


def _edge(p, orc, fid):
    vals = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, 2**255 % p, 2**64 - 1, 2**32 - 1, 2**224, 2**96, 2**192 - 1]
    return orc.to_mont([v % p for v in vals], fid)


@pytest.mark.parametrize("fid,p", [(0, Q), (1, P_BASE)])
def test_mul_add_sub_vs_oracle(ctx, orc, fid, p):
    rng = np.random.default_rng(12345 + fid)
    e = _edge(p, orc, fid)
    a = np.concatenate([rand_fe(rng, 4000), np.repeat(e, len(e), axis=0)])
    b = np.concatenate([rand_fe(rng, 4000), np.tile(e, (len(e), 1))])
    assert np.array_equal(_op(ctx, fid, 0, a, b), orc.f_mul(a, b, fid))
    assert np.array_equal(_op(ctx, fid, 1, a, b), orc.f_add(a, b, fid))
    assert np.array_equal(_op(ctx, fid, 2, a, b), orc.f_sub(a, b, fid))


@pytest.mark.parametrize("fid,p", [(0, Q), (1, P_BASE)])
def test_inverse_and_mont_roundtrip(ctx, orc, fid, p):
    rng = np.random.default_rng(5 + fid)
    a = np.concatenate([rand_fe(rng, 64), _edge(p, orc, fid)])
    inv = _op(ctx, fid, 3, a)
    assert np.array_equal(inv, orc.f_inv(a, fid))
    raw = _op(ctx, fid, 4, a)                       # from_mont: canonical integers
    ints = [sum(int(x[k]) << (64 * k) for k in range(4)) for x in raw]
    assert ints == orc.from_mont(a, fid)
    if fid == 0:
        assert np.array_equal(_op(ctx, fid, 5, raw), a)   # to_mont round trip


def test_delayed_reduction_dot(ctx, orc):
    # delayed_reduction.rs:70-94: sum a_i*b_i via the wide accumulator == sum of reduced products (n = 1000, seed 54321)
    rng = np.random.default_rng(54321)
    for n in (1, 7, 1000, 5000):
        a, b = rand_fe(rng, n), rand_fe(rng, n)
        out = np.zeros((1, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_test_dot_delayed(ctx.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_uint64(n),
                                             out.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(out, orc.f_dot_delayed(a, b))
        ai, bi = orc.from_mont(a), orc.from_mont(b)
        assert orc.from_mont(out)[0] == sum(x * y for x, y in zip(ai, bi)) % Q
    # worst case: all operands p-1 (largest products)
    n = 4096
    a = orc.to_mont([Q - 1] * n); out = np.zeros((1, 4), dtype=np.uint64)
    ctx.check(ctx.L.sp2_test_dot_delayed(ctx.h, a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), C.c_uint64(n), out.ctypes.data_as(C.c_void_p)))
    assert orc.from_mont(out)[0] == n % Q


def test_device_transcript_matches_keccak_kat(ctx, orc):
    """Device Keccak256Transcript squeeze == oracle (which is pinned by keccak.rs:146-163 KATs)."""
    import spartan2_b200 as sp
    rng = random.Random(3)
    for plen in (0, 1, 33, 65, 66, 67, 97, 135, 136, 137, 200, 271, 272, 273, 1000):
        t = orc.Transcript(b"SpartanSNARK")
        t.absorb_bytes(b"x", bytes([7] * 5)); t.squeeze(b"t")
        st, rnd = t.state()
        ts = sp.TranscriptState.make(st, rnd)
        pend = bytes(rng.randrange(256) for _ in range(plen))
        for _ in range(3):
            t.absorb_bytes(b"", pend)
            want = t.squeeze(b"c")
            got = np.zeros((1, 4), dtype=np.uint64)
            buf = np.frombuffer(pend, dtype=np.uint8).copy() if plen else np.zeros(1, dtype=np.uint8)
            ctx.check(ctx.L.sp2_test_transcript(ctx.h, C.byref(ts), buf.ctypes.data_as(C.c_void_p), C.c_uint32(plen), b"c",
                                                got.ctypes.data_as(C.c_void_p)))
            assert np.array_equal(got, want), plen
            assert ts.get() == t.state()
