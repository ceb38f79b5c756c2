"""T256 test points: seeded multiples of the generator, through the oracle (numpy (n,8) u64 affine Montgomery)."""
import numpy as np

GX = 3
GY = 0x5a6dd32df58708e64e97345cbe66600decd9d538a351bb3c30b4954925b1f02d
ORDER = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff


def generator(orc):
    return np.concatenate([orc.to_mont([GX], orc.FP), orc.to_mont([GY], orc.FP)], axis=1)


_cache = {}


def points(orc, n, seed=7):
    """n pseudo-random curve points (never the identity): a random walk P_{i+1} = P_i + D_j from a few seeded
    multiples of G — cheap to generate in bulk, still 'random-looking' bases for a commitment key."""
    key = (n, seed)
    if key in _cache:
        return _cache[key]
    rng = np.random.default_rng(seed)
    g = generator(orc)
    deltas = []
    for _ in range(8):
        k = int.from_bytes(rng.bytes(32), "little") % ORDER
        deltas.append(orc.scalar_mul(g, orc.to_mont([k])))
    out = np.zeros((n, 8), dtype=np.uint64)
    cur = orc.scalar_mul(g, orc.to_mont([int.from_bytes(rng.bytes(32), "little") % ORDER]))
    for i in range(n):
        out[i] = cur[0]
        cur = orc.point_add(cur, deltas[int(rng.integers(0, 8))])
    _cache[key] = out
    return out
