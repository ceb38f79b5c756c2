"""ctypes wrapper of frontend/sha256_r1cs.cpp: the reference's benchmark circuits as padded CSR matrices."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsha256r1cs.so")
SRC = os.path.join(HERE, "sha256_r1cs.cpp")
Q = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
R = 1 << 256
_lib = None


def build_frontend(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_frontend())
        _lib.sha_r1cs_build.restype = C.c_void_p
    return _lib


def mont_limbs(v):
    v = (v % Q) * R % Q
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


class Sha256Circuit:
    """kind="spartan": Sha256Circuit(preimage) of benches/sha256_spartan.rs:37-152 (digest bits are the 256 public
    inputs).  kind="compression": one compression over 64 input bytes with the constant IV (the unit whose size the
    reference quotes, benches/sha256_neutronnova.rs:159-160)."""

    def __init__(self, preimage, kind="spartan", width=2048):
        L = _load()
        data = bytes(preimage)
        buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, dtype=np.uint8)
        self.h = C.c_void_p(L.sha_r1cs_build(C.c_int(0 if kind == "spartan" else 1), buf.ctypes.data_as(C.c_void_p), C.c_uint64(len(data)), C.c_uint64(width)))
        s = np.zeros(10, dtype=np.uint64)
        L.sha_r1cs_sizes(self.h, s.ctypes.data_as(C.c_void_p))
        (self.num_cons_unpadded, self.num_aux, self.num_public, self.num_cons, self.num_precommitted, self.num_rest) = [int(x) for x in s[:6]]
        self.nnz = [int(x) for x in s[6:9]]
        self.width = width
        self.num_shared = 0
        self.num_vars = self.num_precommitted + self.num_rest
        self.raw = []
        for k in range(3):
            coef = np.zeros(max(self.nnz[k], 1), dtype=np.uint32); idx = np.zeros(max(self.nnz[k], 1), dtype=np.uint32)
            ptr = np.zeros(self.num_cons + 1, dtype=np.uint32)
            L.sha_r1cs_export(self.h, C.c_int(k), coef.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), ptr.ctypes.data_as(C.c_void_p))
            self.raw.append((coef[: self.nnz[k]], idx[: self.nnz[k]], ptr))
        nd = int(s[9])
        sign = np.zeros(nd, dtype=np.uint8); mag = np.zeros((nd, 6), dtype=np.uint64)
        L.sha_r1cs_dict(self.h, sign.ctypes.data_as(C.c_void_p), mag.ctypes.data_as(C.c_void_p))
        self.coef_values = [(-1 if sign[i] else 1) * sum(int(mag[i, j]) << (64 * j) for j in range(6)) for i in range(nd)]
        aux = np.zeros(max(self.num_aux, 1), dtype=np.uint8); pub = np.zeros(max(self.num_public, 1), dtype=np.uint8); dg = np.zeros(256, dtype=np.uint8)
        L.sha_r1cs_witness(self.h, aux.ctypes.data_as(C.c_void_p), pub.ctypes.data_as(C.c_void_p), dg.ctypes.data_as(C.c_void_p))
        self.aux_bits = aux[: self.num_aux]; self.pub_bits = pub[: self.num_public]
        self.digest = bytes(np.packbits(dg))

    def __del__(self):
        try:
            if self.h:
                _load().sha_r1cs_free(self.h); self.h = None
        except Exception:
            pass

    # ---- views in the field ------------------------------------------------------------------------
    def matrices(self):
        """(A, B, C) as (data (nnz,4) u64 Montgomery, indices u32, indptr u32) — the padded SplitR1CSShape matrices."""
        table = np.array([mont_limbs(v) for v in self.coef_values], dtype=np.uint64).reshape(-1, 4)
        return [(table[c], i, p) for (c, i, p) in self.raw]

    def dims(self):
        return (self.num_cons, self.num_cons_unpadded, 0, self.num_precommitted, self.num_rest, self.num_public, 0)

    def witness(self):
        """W (num_vars, 4) Montgomery: the precommitted bits then zero padding; X (num_public, 4)."""
        one = np.array(mont_limbs(1), dtype=np.uint64)
        W = np.zeros((self.num_vars, 4), dtype=np.uint64)
        W[: self.num_aux][self.aux_bits == 1] = one
        X = np.zeros((self.num_public, 4), dtype=np.uint64)
        X[self.pub_bits == 1] = one
        return W, X

    def z_int(self):
        """z = [W | 1 | X] as small integers (every value is a bit)."""
        z = np.zeros(self.num_vars + 1 + self.num_public, dtype=np.int64)
        z[: self.num_aux] = self.aux_bits; z[self.num_vars] = 1; z[self.num_vars + 1:] = self.pub_bits
        return z

    def is_satisfied(self):
        """Az o Bz == Cz over the integers, exactly (C++ big-integer check; every variable is a bit)."""
        aux = np.ascontiguousarray(self.aux_bits); pub = np.ascontiguousarray(self.pub_bits) if self.num_public else np.zeros(1, dtype=np.uint8)
        return bool(_load().sha_r1cs_check(self.h, aux.ctypes.data_as(C.c_void_p), pub.ctypes.data_as(C.c_void_p)))
