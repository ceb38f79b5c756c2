"""ORACLE — TEST INFRASTRUCTURE ONLY.  Pure-Python big-int restatements, small cases only.

Used to pin the C oracle: everything here is written from the *definitions* (not from the
optimised code paths), so agreement with oracle.c checks the reference's optimised algorithms
(split-eq / BDDT two-sum trick, delayed reduction, signed Pippenger) against first principles.

  keccak256 / Transcript   reference src/provider/keccak.rs:18-105 (sha3 0.10 Keccak256)
  eq_evals                 reference src/polys/eq.rs:59-92   (MSB-first)
  sumcheck_cubic_naive     definition of the round polynomial of  sum_x eq(tau,x)(A(x)B(x)-C(x))
  sumcheck_quad_naive      definition of the round polynomial of  sum_x A(x)B(x)
  Curve                    y^2 = x^3 + ax + b affine arithmetic (halo2curves t256 parameters)
"""

P_T256_SCALAR = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
P_T256_BASE = 0xffffffff0000000100000000000000017e72b42b30e7317793135661b1c4b117
P_PALLAS_SCALAR = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
T256_A = P_T256_BASE - 3
T256_B = 0xb441071b12f4a0366fb552f8e21ed4ac36b06aceeb354224863e60f20219fc56
T256_G = (3, 0x5a6dd32df58708e64e97345cbe66600decd9d538a351bb3c30b4954925b1f02d)

_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808a, 0x8000000080008000, 0x000000000000808b,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008a, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000a, 0x000000008000808b, 0x800000000000008b, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800a, 0x800000008000000a,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def _keccak_f(s):
    for rnd in range(24):
        c = [s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20] for x in range(5)]
        d = [c[(x + 4) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        s = [s[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        b[0] = s[0]
        # rho + pi with the standard (t+1)(t+2)/2 offsets
        xx, yy = 1, 0
        for t in range(24):
            r = ((t + 1) * (t + 2) // 2) % 64
            nx, ny = yy, (2 * xx + 3 * yy) % 5
            b[nx + 5 * ny] = _rol(s[xx + 5 * yy], r)
            xx, yy = nx, ny
        s = [b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & _M64 & b[(x + 2) % 5 + 5 * y]) for y in range(5) for x in range(5)]
        s[0] ^= _RC[rnd]
    return s


def keccak256(data):
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    s = [0] * 25
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            s[i] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        s = _keccak_f(s)
    return b"".join(s[i].to_bytes(8, "little") for i in range(4))


class Transcript:
    """Keccak256Transcript (keccak.rs:26-105) over an arbitrary prime modulus."""

    def __init__(self, label, modulus):
        self.p = modulus
        self.round = 0
        self.buf = b""
        self.state = self._upd(b"NoTR" + label)

    @staticmethod
    def _upd(inp):
        return keccak256(inp + b"\x00") + keccak256(inp + b"\x01")

    def absorb_bytes(self, label, data):
        self.buf += label + data

    def absorb_scalar(self, label, v):          # big-endian (provider/traits.rs:282-286)
        self.buf += label + (v % self.p).to_bytes(32, "big")

    def absorb_scalars(self, label, vs):
        self.buf += label + b"".join((v % self.p).to_bytes(32, "big") for v in vs)

    def absorb_unipoly(self, coeffs):           # little-endian, linear term skipped (univariate.rs:182-190)
        self.buf += b"p" + b"".join((c % self.p).to_bytes(32, "little") for i, c in enumerate(coeffs) if i != 1)

    def dom_sep(self, b):
        self.buf += b"NoDS" + b

    def squeeze(self, label):
        inp = self.buf + b"NoDS" + self.round.to_bytes(2, "little") + self.state + label
        out = self._upd(inp)
        self.round += 1
        self.state = out
        self.buf = b""
        return int.from_bytes(out, "little") % self.p   # from_uniform_bytes: LE 512-bit mod p


def eq_evals(r, p):
    ev = [1]
    for rv in reversed(r):
        hi = [x * rv % p for x in ev]
        ev = [(x - y) % p for x, y in zip(ev, hi)] + hi
    return ev


def mle_eval(Z, r, p):
    return sum(z * e for z, e in zip(Z, eq_evals(r, p))) % p


def bind_top(Z, r, p):
    n = len(Z) // 2
    return [(Z[i] + r * (Z[n + i] - Z[i])) % p for i in range(n)]


def interpolate(evals, p):
    """coefficients (low->high) of the polynomial through (0,e0),(1,e1),... by Lagrange"""
    n = len(evals)
    coeffs = [0] * n
    for i, e in enumerate(evals):
        num = [1]; den = 1
        for j in range(n):
            if j == i:
                continue
            num = [((num[k - 1] if k > 0 else 0) - j * (num[k] if k < len(num) else 0)) % p for k in range(len(num) + 1)]
            den = den * (i - j) % p
        s = e * pow(den, -1, p) % p
        for k in range(len(num)):
            coeffs[k] = (coeffs[k] + s * num[k]) % p
    return coeffs


def poly_eval(c, x, p):
    acc = 0
    for ci in reversed(c):
        acc = (acc * x + ci) % p
    return acc


def sumcheck_cubic_naive(taus, A, B, C, ts, p):
    """Round i message = coefficients of s_i(X) = sum_x eq(tau,(r_<i,X,x)) (A B - C)(r_<i,X,x),
    computed at X=0..3 directly from the tables (no split-eq, no claim derivation)."""
    l = len(taus)
    polys, rs = [], []
    E = eq_evals(taus, p)
    for _ in range(l):
        n = len(A) // 2
        evals = []
        for X in range(4):
            tot = 0
            for i in range(n):
                a = (A[i] + X * (A[n + i] - A[i])) % p
                b = (B[i] + X * (B[n + i] - B[i])) % p
                c = (C[i] + X * (C[n + i] - C[i])) % p
                e = (E[i] + X * (E[n + i] - E[i])) % p
                tot += e * (a * b - c)
            evals.append(tot % p)
        co = interpolate(evals, p)
        ts.absorb_unipoly(co)
        r = ts.squeeze(b"c")
        polys.append(co); rs.append(r)
        A, B, C, E = bind_top(A, r, p), bind_top(B, r, p), bind_top(C, r, p), bind_top(E, r, p)
    return polys, rs, (A[0], B[0], C[0])


def sumcheck_quad_naive(rounds, A, B, ts, p):
    polys, rs = [], []
    for _ in range(rounds):
        n = len(A) // 2
        evals = []
        for X in range(3):
            tot = 0
            for i in range(n):
                tot += (A[i] + X * (A[n + i] - A[i])) * (B[i] + X * (B[n + i] - B[i]))
            evals.append(tot % p)
        co = interpolate(evals, p)
        ts.absorb_unipoly(co)
        r = ts.squeeze(b"c")
        polys.append(co); rs.append(r)
        A, B = bind_top(A, r, p), bind_top(B, r, p)
    return polys, rs, (A[0], B[0])


class Curve:
    """Affine short-Weierstrass arithmetic with python ints; None is the identity."""

    def __init__(self, q=P_T256_BASE, a=T256_A, b=T256_B):
        self.q, self.a, self.b = q, a, b

    def on_curve(self, P):
        if P is None:
            return True
        x, y = P
        return (y * y - (x * x * x + self.a * x + self.b)) % self.q == 0

    def add(self, P, Q):
        q = self.q
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if (P[1] + Q[1]) % q == 0:
                return None
            lam = (3 * P[0] * P[0] + self.a) * pow(2 * P[1], -1, q) % q
        else:
            lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, q) % q
        x = (lam * lam - P[0] - Q[0]) % q
        return (x, (lam * (P[0] - x) - P[1]) % q)

    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.q)

    def mul(self, k, P):
        R = None
        while k:
            if k & 1:
                R = self.add(R, P)
            P = self.add(P, P)
            k >>= 1
        return R

    def msm(self, ks, Ps):
        R = None
        for k, P in zip(ks, Ps):
            R = self.add(R, self.mul(k, P))
        return R
