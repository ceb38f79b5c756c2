"""The curve arithmetic the CUDA MSM kernels run (spartan2_b200/csrc/curve.cuh), compiled for the host, against
the oracle's group law.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from tests.curve_util import ORDER, generator, points
from tests.hostlib.build import build


@pytest.fixture(scope="module")
def hc():
    return C.CDLL(build("host_curve"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_generator_on_curve(hc, orc):
    g = generator(orc)
    assert hc.ht_on_curve(_p(g)) == 1 and orc.on_curve(g)
    bad = g.copy(); bad[0, 0] ^= np.uint64(1)
    assert hc.ht_on_curve(_p(bad)) == 0


def test_add_mixed_full_double_vs_oracle(hc, orc):
    pts = points(orc, 40)
    a, b = pts[:20].copy(), pts[20:].copy()
    # special cases: P+P, P+(-P), identity operands
    zero = np.zeros((1, 8), dtype=np.uint64)
    neg = a[:1].copy(); neg[0, 4:] = orc.f_sub(np.zeros((1, 4), dtype=np.uint64), neg[:, 4:], orc.FP)
    a2 = np.concatenate([a, a[:1], a[:1], zero, a[:1], zero])
    b2 = np.concatenate([b, a[:1], neg, a[:1], zero, zero])
    n = a2.shape[0]
    want = np.concatenate([orc.point_add(a2[i:i + 1], b2[i:i + 1]) for i in range(n)])
    for fn in (hc.ht_add_mixed, hc.ht_add_full):
        got = np.zeros_like(a2)
        fn(_p(a2), _p(b2), _p(got), C.c_size_t(n))
        assert np.array_equal(got, want)
    # doubling and (2a)+(2b) with non-trivial Z
    d = np.zeros_like(a); hc.ht_dbl(_p(a), _p(d), C.c_size_t(a.shape[0]))
    wd = np.concatenate([orc.point_add(a[i:i + 1], a[i:i + 1]) for i in range(a.shape[0])])
    assert np.array_equal(d, wd)
    got = np.zeros_like(a); hc.ht_add_full_of_doubles(_p(a), _p(b), _p(got), C.c_size_t(a.shape[0]))
    db = np.concatenate([orc.point_add(b[i:i + 1], b[i:i + 1]) for i in range(b.shape[0])])
    want = np.concatenate([orc.point_add(wd[i:i + 1], db[i:i + 1]) for i in range(a.shape[0])])
    assert np.array_equal(got, want)


def test_scalar_mul_vs_oracle(hc, orc):
    g = generator(orc)
    rng = np.random.default_rng(3)
    for k in [0, 1, 2, 3, ORDER - 1, ORDER, int.from_bytes(rng.bytes(32), "little") % ORDER]:
        kk = np.array([(k >> (32 * i)) & 0xffffffff for i in range(8)], dtype=np.uint32)
        got = np.zeros((1, 8), dtype=np.uint64)
        hc.ht_scalar_mul(_p(g), _p(kk), _p(got))
        want = orc.scalar_mul(g, orc.to_mont([k % ORDER]))
        assert np.array_equal(got, want), hex(k)
    # order * G = identity
    kk = np.array([(ORDER >> (32 * i)) & 0xffffffff for i in range(8)], dtype=np.uint32)
    got = np.ones((1, 8), dtype=np.uint64); hc.ht_scalar_mul(_p(g), _p(kk), _p(got))
    assert not got.any()
