"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Field elements travel as numpy uint64 arrays of shape (n, 4): little-endian
64-bit limbs in Montgomery form (R = 2^256), exactly the layout of the reference's
`MontgomeryLimbs::to_limbs()` (reference src/big_num/montgomery.rs:17-22).  Affine points are
(n, 8) uint64: x limbs then y limbs in the T256 base field, identity = all zero.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FQ, FP, FPALLAS = 0, 1, 2
P_T256_SCALAR = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
P_T256_BASE = 0xffffffff0000000100000000000000017e72b42b30e7317793135661b1c4b117
P_PALLAS_SCALAR = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
MODS = {FQ: P_T256_SCALAR, FP: P_T256_BASE, FPALLAS: P_PALLAS_SCALAR}
R = 1 << 256


def build(native=False):
    """Compile the oracle (gcc).  native=True builds a -march=native copy for CPU timing."""
    out = "_build/native/liboracle.so" if native else "_build/liboracle.so"
    args = ["make", "-C", _HERE, "OUT=" + out]
    if native:
        args.append("MARCH=native")
    subprocess.run(args, check=True, capture_output=True)
    return os.path.join(_HERE, out)


def lib(native=False):
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "_build", "liboracle.so")
    if native:
        try:
            path = build(native=True)
        except Exception:
            pass
    if not os.path.exists(path):
        path = build()
    L = C.CDLL(path)
    L.orc_init()
    L.orc_ts_new.restype = C.c_void_p
    L.orc_shape_new.restype = C.c_void_p
    L.orc_get_max_threads.restype = C.c_int
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fe_array(n):
    return np.zeros((n, 4), dtype=np.uint64)


def int_to_limbs(v):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def limbs_to_int(l):
    return sum(int(l[i]) << (64 * i) for i in range(4))


def to_mont(vals, fid=FQ):
    """python ints (canonical) -> (n,4) Montgomery limbs"""
    p = MODS[fid]
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = int_to_limbs((v % p) * R % p)
    return out


def from_mont(arr, fid=FQ):
    p = MODS[fid]
    rinv = pow(R, -1, p)
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [limbs_to_int(a) * rinv % p for a in arr]


def f_binop(name, a, b, fid=FQ):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    o = np.zeros_like(a)
    getattr(lib(), name)(C.c_int(fid), _p(a), _p(b), _p(o), C.c_size_t(a.shape[0]))
    return o


def f_mul(a, b, fid=FQ): return f_binop("orc_f_mul", a, b, fid)
def f_add(a, b, fid=FQ): return f_binop("orc_f_add", a, b, fid)
def f_sub(a, b, fid=FQ): return f_binop("orc_f_sub", a, b, fid)


def f_inv(a, fid=FQ):
    a = np.ascontiguousarray(a, dtype=np.uint64); o = np.zeros_like(a)
    lib().orc_f_inv(C.c_int(fid), _p(a), _p(o), C.c_size_t(a.shape[0]))
    return o


def f_dot_delayed(a, b, fid=FQ):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    o = fe_array(1)
    lib().orc_f_dot_delayed(C.c_int(fid), _p(a), _p(b), C.c_size_t(a.shape[0]), _p(o))
    return o


def f_reduce9(limbs9, fid=FQ):
    a = np.ascontiguousarray(limbs9, dtype=np.uint64); o = fe_array(1)
    lib().orc_f_reduce9(C.c_int(fid), _p(a), _p(o))
    return o


def f_from_uniform(b, fid=FQ):
    buf = np.frombuffer(bytes(b), dtype=np.uint8).copy(); n = len(buf) // 64
    o = fe_array(n)
    lib().orc_f_from_uniform(C.c_int(fid), _p(buf), _p(o), C.c_size_t(n))
    return o


def f_constants(fid):
    mod, r1, r2 = fe_array(1), fe_array(1), fe_array(1)
    inv = C.c_uint64(); ms = C.c_int()
    lib().orc_f_constants(C.c_int(fid), _p(mod), _p(r1), _p(r2), C.byref(inv), C.byref(ms))
    return limbs_to_int(mod[0]), limbs_to_int(r1[0]), limbs_to_int(r2[0]), inv.value, ms.value


def keccak256(data):
    buf = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_keccak256(_p(buf), C.c_size_t(len(data)), _p(out))
    return bytes(out)


class Transcript:
    def __init__(self, label):
        self.h = C.c_void_p(lib().orc_ts_new(label))

    def __del__(self):
        try:
            lib().orc_ts_free(self.h)
        except Exception:
            pass

    def absorb_bytes(self, label, data):
        buf = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
        lib().orc_ts_absorb_bytes(self.h, label, _p(buf), C.c_size_t(len(data)))

    def absorb_scalars(self, label, arr, fid=FQ):
        arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
        lib().orc_ts_absorb_scalars(self.h, C.c_int(fid), label, _p(arr), C.c_size_t(arr.shape[0]))

    def absorb_commitment(self, label, pts):
        pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 8)
        lib().orc_ts_absorb_commitment(self.h, label, _p(pts), C.c_size_t(pts.shape[0]))

    def dom_sep(self, b):
        lib().orc_ts_dom_sep(self.h, b)

    def squeeze(self, label, fid=FQ):
        o = fe_array(1)
        lib().orc_ts_squeeze(self.h, C.c_int(fid), label, _p(o))
        return o

    def state(self):
        st = np.zeros(64, dtype=np.uint8); rnd = C.c_uint16()
        lib().orc_ts_state(self.h, _p(st), C.byref(rnd))
        return bytes(st), rnd.value


def eq_evals(r):
    r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
    out = fe_array(1 << r.shape[0])
    lib().orc_eq_evals(_p(r), C.c_size_t(r.shape[0]), _p(out))
    return out


def bind_top(Z, r):
    Z = np.ascontiguousarray(Z, dtype=np.uint64).copy(); r = np.ascontiguousarray(r, dtype=np.uint64)
    lib().orc_bind_top(_p(Z), C.c_size_t(Z.shape[0]), _p(r))
    return Z[: Z.shape[0] // 2]


def unipoly_from_evals(evals):
    evals = np.ascontiguousarray(evals, dtype=np.uint64); n = evals.shape[0]
    out = fe_array(n)
    lib().orc_unipoly_from_evals(_p(evals), C.c_int(n), _p(out))
    return out


def unipoly_eval(coeffs, r):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
    o = fe_array(1)
    lib().orc_unipoly_eval(_p(coeffs), C.c_int(coeffs.shape[0]), _p(r), _p(o))
    return o


def sumcheck_cubic_prove(claim, taus, A, B, Cc, ts):
    """returns (polys (l,4,4), r (l,4), claims (3,4), raw (t0,tinf) sums (l,2,4)); A,B,C are copied"""
    taus = np.ascontiguousarray(taus, dtype=np.uint64); l = taus.shape[0]
    A = np.ascontiguousarray(A, dtype=np.uint64).copy(); B = np.ascontiguousarray(B, dtype=np.uint64).copy()
    Cc = np.ascontiguousarray(Cc, dtype=np.uint64).copy(); claim = np.ascontiguousarray(claim, dtype=np.uint64)
    polys = fe_array(4 * l); r = fe_array(l); claims = fe_array(3); traw = fe_array(2 * l)
    lib().orc_sumcheck_cubic_prove(_p(claim), _p(taus), C.c_size_t(l), _p(A), _p(B), _p(Cc), ts.h,
                                   _p(polys), _p(r), _p(claims), _p(traw))
    return polys.reshape(l, 4, 4), r, claims, traw.reshape(l, 2, 4)


def sumcheck_quad_prove(claim, rounds, A, B, ts):
    A = np.ascontiguousarray(A, dtype=np.uint64).copy(); B = np.ascontiguousarray(B, dtype=np.uint64).copy()
    claim = np.ascontiguousarray(claim, dtype=np.uint64)
    polys = fe_array(3 * rounds); r = fe_array(rounds); claims = fe_array(2)
    lib().orc_sumcheck_quad_prove(_p(claim), C.c_size_t(rounds), _p(A), _p(B), ts.h, _p(polys), _p(r), _p(claims))
    return polys.reshape(rounds, 3, 4), r, claims


def sumcheck_verify(polys, degree, claim, ts):
    """polys: (rounds, degree+1, 4) full coefficients; the linear slot is ignored (re-derived)."""
    polys = np.ascontiguousarray(polys, dtype=np.uint64).reshape(-1, degree + 1, 4); rounds = polys.shape[0]
    claim = np.ascontiguousarray(claim, dtype=np.uint64)
    e = fe_array(1); r = fe_array(rounds)
    lib().orc_sumcheck_verify(_p(polys), C.c_size_t(rounds), C.c_int(degree), _p(claim), ts.h, _p(e), _p(r))
    return e, r


class Shape:
    """SplitR1CSShape handle: padded CSR matrices (data (nnz,4) u64, indices u32, indptr u32)."""

    def __init__(self, num_cons, num_cons_unpadded, num_shared, num_precommitted, num_rest, num_public, num_challenges, A, B, Cm):
        self.dims = (num_cons, num_cons_unpadded, num_shared, num_precommitted, num_rest, num_public, num_challenges)
        self.num_cons = num_cons; self.num_vars = num_shared + num_precommitted + num_rest
        self.num_public = num_public; self.num_challenges = num_challenges
        self.num_shared = num_shared; self.num_precommitted = num_precommitted
        self._keep = []
        args = [C.c_size_t(x) for x in self.dims]
        for (d, i, p) in (A, B, Cm):
            d = np.ascontiguousarray(d, dtype=np.uint64).reshape(-1, 4)
            if d.shape[0] == 0:
                d = fe_array(1)
            i = np.ascontiguousarray(i, dtype=np.uint32) if len(i) else np.zeros(1, dtype=np.uint32)
            p = np.ascontiguousarray(p, dtype=np.uint32)
            self._keep += [d, i, p]
            args += [_p(d), _p(i), _p(p)]
        self.h = C.c_void_p(lib().orc_shape_new(*args))

    def __del__(self):
        try:
            lib().orc_shape_free(self.h)
        except Exception:
            pass

    def multiply_vec(self, z):
        z = np.ascontiguousarray(z, dtype=np.uint64)
        az, bz, cz = fe_array(self.num_cons), fe_array(self.num_cons), fe_array(self.num_cons)
        lib().orc_shape_multiply_vec(self.h, _p(z), _p(az), _p(bz), _p(cz))
        return az, bz, cz

    def abc(self, rx, r):
        rx = np.ascontiguousarray(rx, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
        out_len = self.num_vars + 1 + self.num_public + self.num_challenges
        out = fe_array(out_len)
        lib().orc_shape_abc(self.h, _p(rx), _p(r), _p(out), C.c_size_t(out_len))
        return out

    def eval_tables(self, tx, ty):
        tx = np.ascontiguousarray(tx, dtype=np.uint64); ty = np.ascontiguousarray(ty, dtype=np.uint64)
        out = fe_array(3)
        lib().orc_shape_eval_tables(self.h, _p(tx), _p(ty), _p(out))
        return out


def csr_multiply_vec(rows, data, indices, indptr, z):
    data = np.ascontiguousarray(data, dtype=np.uint64); indices = np.ascontiguousarray(indices, dtype=np.uint32)
    indptr = np.ascontiguousarray(indptr, dtype=np.uint32); z = np.ascontiguousarray(z, dtype=np.uint64)
    out = fe_array(rows)
    lib().orc_csr_multiply_vec(C.c_size_t(rows), _p(data), _p(indices), _p(indptr), _p(z), _p(out))
    return out


def msm(scalars, bases):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64); bases = np.ascontiguousarray(bases, dtype=np.uint64)
    out = np.zeros((1, 8), dtype=np.uint64)
    lib().orc_msm(_p(scalars), _p(bases), C.c_size_t(scalars.shape[0]), _p(out))
    return out


def msm_small(scalars_u64, bases):
    s = np.ascontiguousarray(scalars_u64, dtype=np.uint64); bases = np.ascontiguousarray(bases, dtype=np.uint64)
    out = np.zeros((1, 8), dtype=np.uint64)
    lib().orc_msm_small(_p(s), _p(bases), C.c_size_t(s.shape[0]), _p(out))
    return out


def scalar_mul(base, k):
    base = np.ascontiguousarray(base, dtype=np.uint64); k = np.ascontiguousarray(k, dtype=np.uint64)
    out = np.zeros((1, 8), dtype=np.uint64)
    lib().orc_scalar_mul(_p(base), _p(k), _p(out))
    return out


def point_add(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros((1, 8), dtype=np.uint64)
    lib().orc_point_add(_p(a), _p(b), _p(out))
    return out


def on_curve(p):
    p = np.ascontiguousarray(p, dtype=np.uint64)
    return bool(lib().orc_on_curve(_p(p)))


def hyrax_commit(ck, h, v, blinds, is_small=False):
    ck = np.ascontiguousarray(ck, dtype=np.uint64); h = np.ascontiguousarray(h, dtype=np.uint64)
    v = np.ascontiguousarray(v, dtype=np.uint64); blinds = np.ascontiguousarray(blinds, dtype=np.uint64)
    num_cols = ck.shape[0]; n = v.shape[0]; rows = (n + num_cols - 1) // num_cols
    out = np.zeros((rows, 8), dtype=np.uint64)
    lib().orc_hyrax_commit(_p(ck), C.c_size_t(num_cols), _p(h), _p(v), C.c_size_t(n), _p(blinds), C.c_int(int(is_small)), _p(out))
    return out


def hyrax_bind(poly, L, r_len):
    poly = np.ascontiguousarray(poly, dtype=np.uint64); L = np.ascontiguousarray(L, dtype=np.uint64)
    out = fe_array(r_len)
    lib().orc_hyrax_bind(_p(poly), _p(L), C.c_size_t(L.shape[0]), C.c_size_t(r_len), _p(out))
    return out


class _ProofView(C.Structure):
    _fields_ = [("num_rounds_x", C.c_uint64), ("num_rounds_y", C.c_uint64), ("num_comm_rows", C.c_uint64), ("num_cols", C.c_uint64),
                ("comm_W", C.c_void_p), ("outer_polys", C.c_void_p), ("claims_outer", C.c_void_p), ("inner_polys", C.c_void_p),
                ("eval_W", C.c_void_p), ("blind_eval_W", C.c_void_p), ("delta", C.c_void_p), ("beta", C.c_void_p),
                ("z_vec", C.c_void_p), ("z_delta", C.c_void_p), ("z_beta", C.c_void_p)]


class _KeysView(C.Structure):
    _fields_ = [("ck", C.c_void_p), ("num_cols", C.c_size_t), ("h", C.c_void_p), ("ck_s", C.c_void_p), ("h_s", C.c_void_p)]


class _RandView(C.Structure):
    _fields_ = [("blinds_W", C.c_void_p), ("blind_eval_W", C.c_void_p), ("d_vec", C.c_void_p), ("r_delta", C.c_void_p), ("r_beta", C.c_void_p)]


class Proof:
    """Flat proof buffers (same layout the product's C ABI fills)."""
    FIELDS = ["comm_W", "outer_polys", "claims_outer", "inner_polys", "eval_W", "blind_eval_W", "delta", "beta", "z_vec", "z_delta", "z_beta"]

    def __init__(self, l, nry, rows, num_cols):
        self.l, self.nry, self.rows, self.num_cols = l, nry, rows, num_cols
        self.comm_W = np.zeros((rows, 8), dtype=np.uint64)
        self.outer_polys = fe_array(3 * l); self.claims_outer = fe_array(3); self.inner_polys = fe_array(2 * nry)
        self.eval_W = fe_array(1); self.blind_eval_W = fe_array(1)
        self.delta = np.zeros((1, 8), dtype=np.uint64); self.beta = np.zeros((1, 8), dtype=np.uint64)
        self.z_vec = fe_array(num_cols); self.z_delta = fe_array(1); self.z_beta = fe_array(1)

    def view(self):
        v = _ProofView(self.l, self.nry, self.rows, self.num_cols)
        for f in self.FIELDS:
            setattr(v, f, getattr(self, f).ctypes.data)
        return v

    def equal(self, other):
        return all(np.array_equal(getattr(self, f), getattr(other, f)) for f in self.FIELDS)


class Keys:
    def __init__(self, ck, h, ck_s, h_s):
        self.ck = np.ascontiguousarray(ck, dtype=np.uint64); self.h = np.ascontiguousarray(h, dtype=np.uint64).reshape(1, 8)
        self.ck_s = np.ascontiguousarray(ck_s, dtype=np.uint64).reshape(1, 8); self.h_s = np.ascontiguousarray(h_s, dtype=np.uint64).reshape(1, 8)

    def view(self):
        return _KeysView(self.ck.ctypes.data, self.ck.shape[0], self.h.ctypes.data, self.ck_s.ctypes.data, self.h_s.ctypes.data)


class Rand:
    def __init__(self, blinds_W, blind_eval_W, d_vec, r_delta, r_beta):
        self.a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (blinds_W, blind_eval_W, d_vec, r_delta, r_beta)]

    def view(self):
        return _RandView(*[x.ctypes.data for x in self.a])


def spartan_prep_cached(shape, W):
    """prep_prove's cached partial products (spartan.rs:184-187, multiply_vec_precommitted r1cs/mod.rs:1112-1130): Az, Bz, Cz over
    the shared + precommitted columns only (z = [W_cached | 0 ...])."""
    cl = shape.num_shared + shape.num_precommitted
    z = np.zeros((shape.num_vars + 1 + shape.num_public + shape.num_challenges, 4), dtype=np.uint64)
    z[:cl] = np.ascontiguousarray(W, dtype=np.uint64).reshape(-1, 4)[:cl]
    return shape.multiply_vec(z)


def spartan_prove(shape, keys, vk_digest, public_values, W, comm_pre, rand, want_debug=False, cached=None):
    L = lib()
    num_vars = shape.num_vars; N = shape.num_cons
    l = N.bit_length() - 1; m = num_vars.bit_length() - 1; nry = m + 1
    num_cols = keys.ck.shape[0]; rows = num_vars // num_cols
    P = Proof(l, nry, rows, num_cols)
    pv = P.view(); kv = keys.view(); rv = rand.view()
    dig = np.frombuffer(bytes(vk_digest), dtype=np.uint8).copy()
    pub = np.ascontiguousarray(public_values, dtype=np.uint64).reshape(-1, 4)
    if pub.shape[0] == 0:
        pub = fe_array(1)
    W = np.ascontiguousarray(W, dtype=np.uint64); comm_pre = np.ascontiguousarray(comm_pre, dtype=np.uint64).reshape(-1, 8)
    dbg = fe_array(2 * l + nry)
    phases = (C.c_double * 6)()
    cz = [np.ascontiguousarray(x, dtype=np.uint64) for x in cached] if cached is not None else None
    rc = L.orc_spartan_prove_cached(shape.h, C.byref(kv), _p(dig), _p(pub), _p(W), _p(comm_pre), C.c_size_t(comm_pre.shape[0]),
                                    C.byref(rv), C.byref(pv), _p(dbg), phases, *([_p(x) for x in cz] if cz else [None, None, None]))
    if rc != 0:
        raise RuntimeError("orc_spartan_prove failed: %d" % rc)
    P.phase_ms = list(phases)
    if want_debug:
        return P, dbg
    return P


def spartan_verify(shape, keys, vk_digest, public_values, proof):
    kv = keys.view(); pv = proof.view()
    dig = np.frombuffer(bytes(vk_digest), dtype=np.uint8).copy()
    pub = np.ascontiguousarray(public_values, dtype=np.uint64).reshape(-1, 4)
    if pub.shape[0] == 0:
        pub = fe_array(1)
    return lib().orc_spartan_verify(shape.h, C.byref(kv), _p(dig), _p(pub), C.byref(pv))


def set_threads(n):
    lib().orc_set_threads(C.c_int(n))


def max_threads():
    return lib().orc_get_max_threads()


# ---- NeutronNova building blocks -------------------------------------------------------------------
def pow_split_evals(t, left, right):
    t = np.ascontiguousarray(t, dtype=np.uint64); out = fe_array(left + right)
    lib().orc_pow_split_evals(_p(t), C.c_size_t(left), C.c_size_t(right), _p(out))
    return out


def nifs_round(t, rhos, left, right, E, A, B, Cm, N, m, pair_offset=0):
    """pair_offset: first GLOBAL pair index when A, B, Cm are one rank's contiguous block of the live layers"""
    rhos = np.ascontiguousarray(rhos, dtype=np.uint64).reshape(-1, 4)
    E, A, B, Cm = (np.ascontiguousarray(x, dtype=np.uint64) for x in (E, A, B, Cm))
    out = fe_array(2)
    lib().orc_nifs_round_block(C.c_size_t(t), C.c_size_t(rhos.shape[0]), _p(rhos), C.c_size_t(left), C.c_size_t(right), _p(E), _p(A), _p(B), _p(Cm),
                               C.c_size_t(N), C.c_size_t(m), C.c_size_t(pair_offset), _p(out))
    return out


def nifs_fold(L, N, m, r_b):
    L = np.ascontiguousarray(L, dtype=np.uint64).copy(); r_b = np.ascontiguousarray(r_b, dtype=np.uint64)
    lib().orc_nifs_fold(_p(L), C.c_size_t(N), C.c_size_t(m), _p(r_b))
    return L[: (m // 2) * N]


def weights_from_r(r_bs, n):
    r_bs = np.ascontiguousarray(r_bs, dtype=np.uint64).reshape(-1, 4); out = fe_array(n)
    lib().orc_weights_from_r(_p(r_bs), C.c_size_t(r_bs.shape[0]), C.c_size_t(n), _p(out))
    return out


def fold_vectors(Ws, n, dim, w):
    Ws = np.ascontiguousarray(Ws, dtype=np.uint64); w = np.ascontiguousarray(w, dtype=np.uint64); out = fe_array(dim)
    lib().orc_fold_vectors(_p(Ws), C.c_size_t(n), C.c_size_t(dim), _p(w), _p(out))
    return out


def pow_cubic_eval(pl, pr, A, B, Cm):
    pl, pr, A, B, Cm = (np.ascontiguousarray(x, dtype=np.uint64) for x in (pl, pr, A, B, Cm))
    out = fe_array(3)
    lib().orc_pow_cubic_eval(_p(pl), C.c_size_t(pl.shape[0]), _p(pr), _p(A), _p(B), _p(Cm), C.c_size_t(A.shape[0]), _p(out))
    return out


def quad_eval(A, B):
    A, B = (np.ascontiguousarray(x, dtype=np.uint64) for x in (A, B)); out = fe_array(2)
    lib().orc_quad_eval(_p(A), _p(B), C.c_size_t(A.shape[0]), _p(out))
    return out


def fold_commitments(comms, n, rows, w):
    comms = np.ascontiguousarray(comms, dtype=np.uint64); w = np.ascontiguousarray(w, dtype=np.uint64)
    out = np.zeros((rows, 8), dtype=np.uint64)
    lib().orc_fold_commitments(_p(comms), C.c_size_t(n), C.c_size_t(rows), _p(w), _p(out))
    return out


# ---- small-value path (big_num/small_value.rs and its users in neutronnova_zk.rs) ---------------------------------
def to_small_vec_or_zero(poly):
    """-> (i64 values, ascending positions of the values that did not fit and were stored as 0)"""
    poly = np.ascontiguousarray(poly, dtype=np.uint64).reshape(-1, 4); n = poly.shape[0]
    out = np.zeros(n, dtype=np.int64); large = np.zeros(max(n, 1), dtype=np.uint64)
    L = lib(); L.orc_to_small_vec_or_zero.restype = C.c_size_t
    k = L.orc_to_small_vec_or_zero(_p(poly), C.c_size_t(n), _p(out), _p(large))
    return out, large[:k].copy()


def small_acc_dot(f, vals):
    """sum_j f[j] * vals[j] through SmallAccumulator; vals: python ints in the i128 range"""
    f = np.ascontiguousarray(f, dtype=np.uint64).reshape(-1, 4); n = f.shape[0]
    lo = np.array([v & 0xFFFFFFFFFFFFFFFF for v in vals], dtype=np.uint64)
    hi = np.array([(v >> 64) for v in vals], dtype=np.int64)
    out = fe_array(1)
    lib().orc_small_acc_dot(_p(f), _p(lo), _p(hi), C.c_size_t(n), _p(out))
    return out


def nifs_round0_small(rhos, left, right, E, A, B, A64, B64, large, N, m):
    rhos = np.ascontiguousarray(rhos, dtype=np.uint64).reshape(-1, 4)
    E, A, B = (np.ascontiguousarray(x, dtype=np.uint64) for x in (E, A, B))
    A64, B64 = (np.ascontiguousarray(x, dtype=np.int64) for x in (A64, B64))
    large = np.ascontiguousarray(large, dtype=np.uint64); lp = large if large.shape[0] else np.zeros(1, dtype=np.uint64)
    out = fe_array(2)
    lib().orc_nifs_round0_small(C.c_size_t(rhos.shape[0]), _p(rhos), C.c_size_t(left), C.c_size_t(right), _p(E), _p(A), _p(B), _p(A64), _p(B64),
                                _p(lp), C.c_size_t(large.shape[0]), C.c_size_t(N), C.c_size_t(m), _p(out))
    return out


def nifs_cvals_small(left, right, E, Cl, C64, large, N, n):
    E, Cl = (np.ascontiguousarray(x, dtype=np.uint64) for x in (E, Cl)); C64 = np.ascontiguousarray(C64, dtype=np.int64)
    large = np.ascontiguousarray(large, dtype=np.uint64); lp = large if large.shape[0] else np.zeros(1, dtype=np.uint64)
    out = fe_array(n)
    lib().orc_nifs_cvals_small(C.c_size_t(left), C.c_size_t(right), _p(E), _p(Cl), _p(C64), _p(lp), C.c_size_t(large.shape[0]), C.c_size_t(N),
                               C.c_size_t(n), _p(out))
    return out


# ---- Hyrax commitment family on group elements (hyrax_pc.rs:321-344, 533-607, 795-874) ---------------------------------
def _pts(a):
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 8)


def _fes(a):
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)


def hyrax_commit_without_blind(ck, v, is_small=False):
    ck = _pts(ck); v = _fes(v); num_cols = ck.shape[0]; n = v.shape[0]; rows = (n + num_cols - 1) // num_cols
    out = np.zeros((rows, 8), dtype=np.uint64)
    lib().orc_hyrax_commit_without_blind(_p(ck), C.c_size_t(num_cols), _p(v), C.c_size_t(n), C.c_int(int(is_small)), _p(out))
    return out


def hyrax_commit_incremental(ck, h, raw, delta, blinds):
    ck = _pts(ck); h = _pts(h); raw = _pts(raw); delta = _fes(delta); blinds = _fes(blinds)
    num_cols = ck.shape[0]; n = delta.shape[0]; rows = (n + num_cols - 1) // num_cols
    out = np.zeros((rows, 8), dtype=np.uint64)
    lib().orc_hyrax_commit_incremental(_p(ck), C.c_size_t(num_cols), _p(h), _p(raw), C.c_size_t(raw.shape[0]), _p(delta), C.c_size_t(n), _p(blinds), _p(out))
    return out


def hyrax_rerandomize(h, comm, r_old, r_new):
    h = _pts(h); comm = _pts(comm); r_old = _fes(r_old); r_new = _fes(r_new)
    out = np.zeros_like(comm)
    lib().orc_hyrax_rerandomize(_p(h), _p(comm), _p(r_old), _p(r_new), C.c_size_t(comm.shape[0]), _p(out))
    return out


def fold_blinds(blinds, n, rows, w):
    blinds = _fes(blinds); w = _fes(w); out = fe_array(rows)
    lib().orc_fold_blinds(_p(blinds), C.c_size_t(n), C.c_size_t(rows), _p(w), _p(out))
    return out


def fold_commitments_partial(comms, n, rows, w, num_data_rows, folded_blind, h):
    comms = _pts(comms); w = _fes(w); fb = _fes(folded_blind); h = _pts(h)
    out = np.zeros((rows, 8), dtype=np.uint64)
    lib().orc_fold_commitments_partial(_p(comms), C.c_size_t(n), C.c_size_t(rows), _p(w), C.c_size_t(num_data_rows), _p(fb), _p(h), _p(out))
    return out


# ---- NeutronNova, non-ZK variant (oracle.c: orc_neutronnova_prove / _verify) -------------------------------------------
class _NnProofView(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_steps", "ell_b", "ell", "rounds_y", "rows", "num_cols")] + \
               [(k, C.c_void_p) for k in ("comm_W_steps", "comm_W_core", "nifs_polys", "outer_polys", "claims_outer", "inner_polys", "eval_W",
                                          "blind_eval_W", "delta", "beta", "z_vec", "z_delta", "z_beta")]


class _NnRandView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("blinds_steps", "blinds_core", "blind_eval_W", "d_vec", "r_delta", "r_beta")]


class NnProof:
    """Flat buffers of the non-ZK NeutronNova proof (same layout as the product's sp2_nn_snark)."""
    FIELDS = ["comm_W_steps", "comm_W_core", "nifs_polys", "outer_polys", "claims_outer", "inner_polys", "eval_W", "blind_eval_W",
              "delta", "beta", "z_vec", "z_delta", "z_beta"]

    def __init__(self, n, N, M, num_cols):
        self.n, self.ell_b, self.ell, self.my = n, n.bit_length() - 1, N.bit_length() - 1, (2 * M).bit_length() - 1
        self.rows, self.num_cols = M // num_cols, num_cols
        pz = lambda k: np.zeros((k, 8), dtype=np.uint64)   # noqa: E731
        self.comm_W_steps = pz(n * self.rows); self.comm_W_core = pz(self.rows)
        self.nifs_polys = fe_array(4 * self.ell_b); self.outer_polys = fe_array(8 * self.ell); self.claims_outer = fe_array(6)
        self.inner_polys = fe_array(6 * self.my); self.eval_W = fe_array(2); self.blind_eval_W = fe_array(2)
        self.delta = pz(1); self.beta = pz(1); self.z_vec = fe_array(num_cols); self.z_delta = fe_array(1); self.z_beta = fe_array(1)

    def view(self):
        v = _NnProofView(self.n, self.ell_b, self.ell, self.my, self.rows, self.num_cols)
        for f in self.FIELDS:
            setattr(v, f, getattr(self, f).ctypes.data)
        return v


class NnRand:
    def __init__(self, blinds_steps, blinds_core, blind_eval_W, d_vec, r_delta, r_beta):
        self.a = [_fes(x) for x in (blinds_steps, blinds_core, blind_eval_W, d_vec, r_delta, r_beta)]

    def view(self):
        return _NnRandView(*[x.ctypes.data for x in self.a])


def neutronnova_prove(shape, keys, vk_digest, zs, zc, comm_pre_steps, blinds_pre_steps, comm_pre_core, blinds_pre_core, rand, want_debug=False):
    """zs: (n, num_cols, 4) step instances z_i = [W_i | 1 | X_i]; comm_pre_*: precommitted-section commitments of prep_prove
    (rows of the shared+precommitted sections) with their blinds; rand: NnRand (the blinds every U_i.comm_W ends up with)."""
    zs = np.ascontiguousarray(zs, dtype=np.uint64); n = zs.shape[0]; zc = _fes(zc)
    P = NnProof(n, shape.num_cons, shape.num_vars, keys.ck.shape[0])
    pv = P.view(); kv = keys.view(); rv = rand.view()
    dig = np.frombuffer(bytes(vk_digest), dtype=np.uint8).copy()
    cps, bps, cpc, bpc = _pts(comm_pre_steps), _fes(blinds_pre_steps), _pts(comm_pre_core), _fes(blinds_pre_core)
    dbg = fe_array(8); phases = (C.c_double * 8)()
    rc = lib().orc_neutronnova_prove(shape.h, C.byref(kv), _p(dig), C.c_size_t(n), _p(zs), _p(zc), _p(cps), _p(bps), _p(cpc), _p(bpc),
                                     C.byref(rv), C.byref(pv), _p(dbg), phases)
    if rc != 0:
        raise RuntimeError("orc_neutronnova_prove failed: %d" % rc)
    P.phase_ms = dict(zip(["rerandomize+commit_zeros", "transcript_absorb", "nifs", "fold", "outer_sumcheck_batched", "compute_eval_table_sparse",
                           "inner_sumcheck_batched", "pcs_prove"], list(phases)))
    P.debug = dict(zip(["T_out", "tau_at_rx", "r", "c_eval", "eq_rho_at_rb", "eval_X_step", "eval_X_core"], [dbg[i:i + 1].copy() for i in range(7)]))
    return P


def neutronnova_verify(shape, keys, vk_digest, step_X, core_X, proof):
    kv = keys.view(); pv = proof.view()
    dig = np.frombuffer(bytes(vk_digest), dtype=np.uint8).copy()
    sx = _fes(step_X) if np.size(step_X) else fe_array(1); cx = _fes(core_X) if np.size(core_X) else fe_array(1)
    return lib().orc_neutronnova_verify(shape.h, C.byref(kv), _p(dig), _p(sx), _p(cx), C.byref(pv))


def zero_check_round0(taus, A, B):
    """evaluation_points_zero_check_round0 (sumcheck.rs:1163-1271): (eval_0, eval_2, eval_3) of the first zero-check round"""
    taus = _fes(taus); A = _fes(A); B = _fes(B); out = fe_array(3)
    lib().orc_zero_check_round0(_p(taus), C.c_size_t(taus.shape[0]), _p(A), _p(B), _p(out))
    return out
