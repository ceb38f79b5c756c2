"""The exact arithmetic the CUDA kernels run (spartan2_b200/csrc/field.cuh), compiled for the host
with the PTX carry-chain primitives emulated (prim.cuh), against Python big ints.  CPU only.
The same checks run on the device in tests/test_gpu_field.py."""
import ctypes as C
import random

import numpy as np
import pytest

from tests.hostlib.build import build

R = 1 << 256
Q = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
P = 0xffffffff0000000100000000000000017e72b42b30e7317793135661b1c4b117


@pytest.fixture(scope="module")
def ht():
    return C.CDLL(build("host_field"))


def arr(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            out[i, k] = (v >> (64 * k)) & (2**64 - 1)
    return out


def ints(a):
    return [sum(int(x[k]) << (64 * k) for k in range(4)) for x in np.asarray(a).reshape(-1, 4)]


def call2(fn, a, b):
    a, b = arr(a), arr(b); o = np.zeros_like(a)
    fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
    return ints(o)


def call1(fn, a):
    a = arr(a); o = np.zeros_like(a)
    fn(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
    return ints(o)


def samples(p, n, seed):
    rng = random.Random(seed)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 2**255 % p, 2**64 - 1, 2**32 - 1, 2**224, 2**96, 2**192 - 1]
    edge = [e % p for e in edge]
    a = edge + [rng.randrange(p) for _ in range(n)]
    b = list(reversed(edge)) + [rng.randrange(p) for _ in range(n)]
    # all edge x edge pairs too
    aa = [x for x in edge for _ in edge]; bb = [y for _ in edge for y in edge]
    return a + aa, b + bb


@pytest.mark.parametrize("name,p", [("fq", Q), ("fp", P)])
def test_mont_mul_add_sub(ht, name, p):
    a, b = samples(p, 3000, 12345)
    rinv = pow(R, -1, p)
    assert call2(getattr(ht, "ht_%s_mul" % name), a, b) == [x * y * rinv % p for x, y in zip(a, b)]
    assert call2(getattr(ht, "ht_%s_add" % name), a, b) == [(x + y) % p for x, y in zip(a, b)]
    assert call2(getattr(ht, "ht_%s_sub" % name), a, b) == [(x - y) % p for x, y in zip(a, b)]


@pytest.mark.parametrize("fn,p", [("ht_fq_mul_il", Q), ("ht_fp_mul_il", P), ("ht_fp_mul_cios", P)])
def test_all_multiplication_forms_agree(ht, fn, p):
    """mont_mul_interleaved<> (the multiplication of the group arithmetic; also instantiated for the scalar field) and the
    wide-product + shaped-CIOS form kept as its cross-check: all equal a * b / R mod p on edge x edge pairs and 20k random pairs."""
    a, b = samples(p, 20000, 4242)
    rinv = pow(R, -1, p)
    assert call2(getattr(ht, fn), a, b) == [x * y * rinv % p for x, y in zip(a, b)]


def test_mul_wide(ht):
    rng = random.Random(2)
    for _ in range(300):
        a = rng.choice([rng.randrange(1 << 256), (1 << 256) - 1, rng.randrange(1 << 64)])
        b = rng.choice([rng.randrange(1 << 256), (1 << 256) - 1, 1])
        o = np.zeros(16, dtype=np.uint32)
        A, B = arr([a]), arr([b])
        ht.ht_mul_wide(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert sum(int(o[i]) << (32 * i) for i in range(16)) == a * b


def test_inverse_and_half(ht):
    rng = random.Random(9)
    for name, p in (("fq", Q), ("fp", P)):
        a = [1, 2, p - 1] + [rng.randrange(1, p) for _ in range(10)]
        am = [x * R % p for x in a]
        got = call1(getattr(ht, "ht_%s_inv" % name), am)
        assert got == [pow(x, -1, p) * R % p for x in a]
    a = [0, 1, 2, 3, Q - 1, Q - 2] + [rng.randrange(Q) for _ in range(100)]
    assert call1(ht.ht_fq_half, a) == [x * pow(2, -1, Q) % Q for x in a]


def test_mont_conversions(ht):
    rng = random.Random(10)
    a = [0, 1, Q - 1] + [rng.randrange(Q) for _ in range(100)]
    assert call1(ht.ht_fq_to_mont, a) == [x * R % Q for x in a]
    assert call1(ht.ht_fq_from_mont, [x * R % Q for x in a]) == a
    b = [0, 1, P - 1] + [rng.randrange(P) for _ in range(100)]
    assert call1(ht.ht_fp_from_mont, [x * R % P for x in b]) == b
    # from_uniform: halves may exceed p
    los = [rng.randrange(1 << 256) for _ in range(50)] + [(1 << 256) - 1]
    his = [rng.randrange(1 << 256) for _ in range(50)] + [(1 << 256) - 1]
    inter = [v for pair in zip(los, his) for v in pair]
    x = arr(inter); o = np.zeros((len(los), 4), dtype=np.uint64)
    ht.ht_fq_from_uniform(x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_size_t(len(los)))
    assert ints(o) == [((lo + (hi << 256)) % Q) * R % Q for lo, hi in zip(los, his)]


def test_delayed_reduction(ht):
    # reference big_num/delayed_reduction.rs:70-94 shape: n = 1000, plus worst-case magnitudes
    rng = random.Random(54321)
    rinv = pow(R, -1, Q)
    for a, b in [([rng.randrange(Q) for _ in range(1000)], [rng.randrange(Q) for _ in range(1000)]),
                 ([Q - 1] * 5000, [Q - 1] * 5000), ([0] * 3, [5] * 3), ([1], [1])]:
        A, B = arr(a), arr(b); o = np.zeros((1, 4), dtype=np.uint64)
        ht.ht_fq_dot(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)), o.ctypes.data_as(C.c_void_p))
        assert ints(o)[0] == sum(x * y for x, y in zip(a, b)) * rinv % Q


def test_acc_reduce_extremes(ht):
    rng = random.Random(4)
    rinv = pow(R, -1, Q)
    vals = [0, 1, (1 << 543) - 1, (1 << 512) - 1, (1 << 512), Q * Q, (1 << 543) - Q] + [rng.randrange(1 << 543) for _ in range(500)]
    for v in vals:
        limbs = np.array([(v >> (32 * i)) & 0xffffffff for i in range(17)], dtype=np.uint32)
        o = np.zeros((1, 4), dtype=np.uint64)
        ht.ht_fq_acc_reduce(limbs.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert ints(o)[0] == v * rinv % Q
