"""Field layer of the oracle vs Python big ints — the reference's big_num property tests
(src/big_num/delayed_reduction.rs:70-94, montgomery.rs:188-229, field_reduction_constants.rs:60-109;
instantiated for T256 at src/provider/pt256.rs:71-81) re-expressed as differential tests."""
import random

import numpy as np
import pytest

R = 1 << 256


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_constants(orc, fid):
    p = orc.MODS[fid]
    mod, r1, r2, inv, max_sub = orc.f_constants(fid)
    assert mod == p
    assert r1 == R % p                       # R_MOD == ONE limbs
    assert r2 == (R * R) % p                 # R512_MOD == 2^512 mod p
    assert (inv * (p & (2**64 - 1)) + 1) % 2**64 == 0   # MONT_INV * p0 == -1 mod 2^64
    assert max_sub == R // p
    if fid == 0:                             # P-256 prime has MONT_INV = 1
        assert inv == 1 and max_sub == 1


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_mul_add_sub_inv(orc, fid):
    p = orc.MODS[fid]
    rng = random.Random(12345)               # seed of montgomery.rs:188-229
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, 2**255 % p, (1 << 64) - 1, (1 << 32) - 1]
    a = edge + [rng.randrange(p) for _ in range(100)]
    b = list(reversed(edge)) + [rng.randrange(p) for _ in range(100)]
    A, B = orc.to_mont(a, fid), orc.to_mont(b, fid)
    assert orc.from_mont(orc.f_mul(A, B, fid), fid) == [x * y % p for x, y in zip(a, b)]
    assert orc.from_mont(orc.f_add(A, B, fid), fid) == [(x + y) % p for x, y in zip(a, b)]
    assert orc.from_mont(orc.f_sub(A, B, fid), fid) == [(x - y) % p for x, y in zip(a, b)]
    nz = [x for x in a if x]
    assert orc.from_mont(orc.f_inv(orc.to_mont(nz, fid), fid), fid) == [pow(x, -1, p) for x in nz]
    # outputs are canonical limbs (< p)
    out = orc.f_mul(A, B, fid)
    assert all(orc.limbs_to_int(o) < p for o in out)


@pytest.mark.parametrize("fid", [0, 2])
def test_delayed_reduction_dot(orc, fid):
    # delayed_reduction.rs:70-94: sum a_i*b_i via wide accumulator == sum of reduced products, n=1000
    p = orc.MODS[fid]
    rng = random.Random(54321)
    a = [rng.randrange(p) for _ in range(1000)]; b = [rng.randrange(p) for _ in range(1000)]
    got = orc.from_mont(orc.f_dot_delayed(orc.to_mont(a, fid), orc.to_mont(b, fid), fid), fid)[0]
    assert got == sum(x * y for x, y in zip(a, b)) % p
    # worst case magnitudes: all (p-1)*(p-1)
    a = [p - 1] * 4096
    got = orc.from_mont(orc.f_dot_delayed(orc.to_mont(a, fid), orc.to_mont(a, fid), fid), fid)[0]
    assert got == 4096 * (p - 1) * (p - 1) % p


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_reduce9_fold_identity(orc, fid):
    # montgomery.rs: value = low8 + h*2^512 for h in {1,2,0xFF,2^32-1,2^64-1}; REDC gives value/R mod p
    p = orc.MODS[fid]
    rng = random.Random(99)
    rinv = pow(R, -1, p)
    for h in [0, 1, 2, 0xFF, 2**32 - 1, 2**64 - 1]:
        low = rng.randrange(1 << 512)
        for lowv in (low, (1 << 512) - 1, 0):
            val = lowv + (h << 512)
            limbs = np.array([(val >> (64 * i)) & (2**64 - 1) for i in range(9)], dtype=np.uint64)
            got = orc.limbs_to_int(orc.f_reduce9(limbs, fid)[0])
            assert got == val * rinv % p


@pytest.mark.parametrize("fid", [0, 2])
def test_from_uniform(orc, fid):
    p = orc.MODS[fid]
    rng = random.Random(5)
    blobs = [bytes(rng.getrandbits(8) for _ in range(64)) for _ in range(20)] + [b"\xff" * 64, b"\x00" * 64]
    got = orc.from_mont(orc.f_from_uniform(b"".join(blobs), fid), fid)
    assert got == [int.from_bytes(b, "little") % p for b in blobs]
