"""Host-side mirror of the reference interface over the C ABI (see package docstring)."""
import ctypes as C

import numpy as np

from ._lib import OK, SpartanError, TranscriptState, lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _addr(a):
    """address of a C-contiguous numpy array's data (about half the cost of ndarray.ctypes.data: the per-call overhead of the wrapper
    is part of the end-to-end time of a ~1.3 ms prove)"""
    return C.addressof(C.c_char.from_buffer(a)) if a.flags.writeable and a.size else a.ctypes.data


def _fe(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, 4)


class Context:
    """One per (process, GPU): sp2_ctx."""

    def __init__(self, device=0, stream=None):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.sp2_ctx_create(C.c_int32(device), C.byref(h))
        if rc != OK:
            raise SpartanError(rc, "sp2_ctx_create failed: no usable sm_100 CUDA device %d (no CPU fallback)" % device)
        self.h = h
        self.device = device
        if stream is not None:
            self.check(self.L.sp2_ctx_set_stream(self.h, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pins", []):
                self.L.sp2_host_free(self.h, p)
            self._pins = []
            self.L.sp2_ctx_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != OK:
            raise SpartanError(rc, (self.L.sp2_last_error(self.h) or b"").decode())

    def launch_count(self):
        return int(self.L.sp2_launch_count(self.h))

    def num_sms(self):
        return int(self.L.sp2_num_sms(self.h))

    def synchronize(self):
        self.check(self.L.sp2_synchronize(self.h))

    def timer_start(self):
        self.check(self.L.sp2_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.L.sp2_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def test_points(self, n, seed=1):
        """n pseudo-random T256 points (affine, Montgomery) for test/bench commitment keys."""
        out = np.zeros((n, 8), dtype=np.uint64)
        self.check(self.L.sp2_test_points(self.h, C.c_uint64(seed), C.c_uint32(n), _p(out)))
        return out

    # -- raw memory ---------------------------------------------------------------------------
    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        buf = DeviceBuffer(self, arr.nbytes)
        buf.upload(arr)
        return buf

    def pinned_empty(self, shape, dtype=np.uint64):
        """numpy array backed by page-locked host memory (sp2_host_alloc)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(self.L.sp2_host_alloc(self.h, C.c_uint64(max(n, 32)), C.byref(p)))
        raw = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pins = getattr(self, "_pins", []); self._pins.append(p)
        return arr


class DeviceBuffer:
    def __init__(self, ctx, nbytes):
        self.ctx = ctx; self.nbytes = int(nbytes)
        p = C.c_void_p()
        ctx.check(ctx.L.sp2_dev_alloc(ctx.h, C.c_uint64(self.nbytes), C.byref(p)))
        self.ptr = p

    def upload(self, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        self.ctx.check(self.ctx.L.sp2_dev_upload(self.ctx.h, C.c_void_p(self.ptr.value + offset), _p(arr), C.c_uint64(arr.nbytes)))

    def download(self, shape, dtype=np.uint64, offset=0):
        out = np.zeros(shape, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        self.ctx.check(self.ctx.L.sp2_dev_download(self.ctx.h, _p(out), C.c_void_p(self.ptr.value + offset), C.c_uint64(out.nbytes)))
        return out

    def copy_from(self, other, nbytes=None):
        self.ctx.check(self.ctx.L.sp2_dev_copy(self.ctx.h, self.ptr, other.ptr, C.c_uint64(nbytes or other.nbytes)))

    def free(self):
        if self.ptr and getattr(self.ctx, "h", None):
            self.ctx.L.sp2_dev_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class EqPolynomial:
    """src/polys/eq.rs"""

    @staticmethod
    def evals_from_points(ctx, r):
        r = _fe(r); k = r.shape[0]
        out = np.zeros((1 << k, 4), dtype=np.uint64)
        rr = r if k else np.zeros((1, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_eq_table(ctx.h, _p(rr), C.c_uint32(k), _p(out)))
        return out


class MultilinearPolynomial:
    """src/polys/multilinear.rs"""

    @staticmethod
    def bind_poly_var_top(ctx, Z, r):
        Z = _fe(Z); r = _fe(r)
        out = np.zeros((Z.shape[0] // 2, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_bind_top(ctx.h, _p(Z), C.c_uint64(Z.shape[0]), _p(r), _p(out)))
        return out


class SumcheckProof:
    """src/sumcheck.rs — whole-loop provers.  `ts` is a TranscriptState (round, state) hand-off; it is
    advanced in place exactly as the reference's transcript would be by the l absorb/squeeze pairs."""

    @staticmethod
    def prove_cubic_with_three_inputs(ctx, claim, taus, A, B, Cz, ts):
        """returns (polys (l,4,4) full coefficients low->high, r (l,4), claims (3,4)); A, B, C host arrays
        or DeviceBuffers (bound in place on the device)."""
        taus = _fe(taus); l = taus.shape[0]; claim = _fe(claim)
        polys = np.zeros((l, 4, 4), dtype=np.uint64); r = np.zeros((l, 4), dtype=np.uint64); claims = np.zeros((3, 4), dtype=np.uint64)
        if isinstance(A, DeviceBuffer):
            rc = ctx.L.sp2_sumcheck_cubic_prove_dev(ctx.h, _p(claim), _p(taus), C.c_uint32(l), A.ptr, B.ptr, Cz.ptr,
                                                    C.byref(ts), _p(polys), _p(r), _p(claims))
        else:
            A, B, Cz = _fe(A), _fe(B), _fe(Cz)
            if not (A.shape[0] == B.shape[0] == Cz.shape[0] == (1 << l)):
                raise SpartanError(-2, "tables must have 2^num_rounds entries")
            rc = ctx.L.sp2_sumcheck_cubic_prove(ctx.h, _p(claim), _p(taus), C.c_uint32(l), _p(A), _p(B), _p(Cz),
                                                C.byref(ts), _p(polys), _p(r), _p(claims))
        ctx.check(rc)
        return polys, r, claims

    @staticmethod
    def evaluation_points_zero_check_round0(ctx, taus, A, B):
        """EqSumCheckInstance::evaluation_points_zero_check_round0 (src/sumcheck.rs:1163-1271): (eval_0, eval_2, eval_3); A, B are
        DeviceBuffers or host arrays of 2^len(taus) scalars (not modified)."""
        taus = _fe(taus); l = taus.shape[0]
        if not isinstance(A, DeviceBuffer):
            A = ctx.upload(_fe(A)); B = ctx.upload(_fe(B))
        out = np.zeros((3, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_sc_zero_check_round0_dev(ctx.h, _p(taus), C.c_uint32(l), A.ptr, B.ptr, _p(out)))
        return out

    @staticmethod
    def prove_quad(ctx, claim, num_rounds, A, B, ts):
        """returns (polys (rounds,3,4), r (rounds,4), claims (2,4))"""
        claim = _fe(claim); l = int(num_rounds)
        polys = np.zeros((l, 3, 4), dtype=np.uint64); r = np.zeros((l, 4), dtype=np.uint64); claims = np.zeros((2, 4), dtype=np.uint64)
        if isinstance(A, DeviceBuffer):
            rc = ctx.L.sp2_sumcheck_quad_prove_dev(ctx.h, _p(claim), C.c_uint32(l), A.ptr, B.ptr, C.byref(ts), _p(polys), _p(r), _p(claims))
        else:
            A, B = _fe(A), _fe(B)
            if not (A.shape[0] == B.shape[0] == (1 << l)):
                raise SpartanError(-2, "tables must have 2^num_rounds entries")
            rc = ctx.L.sp2_sumcheck_quad_prove(ctx.h, _p(claim), C.c_uint32(l), _p(A), _p(B), C.byref(ts), _p(polys), _p(r), _p(claims))
        ctx.check(rc)
        return polys, r, claims


class SplitR1CSShape:
    """src/r1cs/mod.rs:743-1398 — device-resident shape (sp2_shape).  A, B, C are padded CSR triples
    (data (nnz,4) u64 Montgomery, indices u32, indptr u32 of num_cons+1 entries)."""

    def __init__(self, ctx, num_cons, num_cons_unpadded, num_shared, num_precommitted, num_rest, num_public, num_challenges, A, B, Cm,
                 rank=0, nranks=1):
        """rank / nranks: this GPU's shard for the multi-GPU prover (rows and transposed columns i = rank mod nranks);
        the same whole matrices are passed on every rank."""
        self.ctx = ctx
        self.rank, self.nranks = rank, nranks
        self.num_cons = num_cons; self.num_vars = num_shared + num_precommitted + num_rest
        self.num_shared, self.num_precommitted, self.num_rest = num_shared, num_precommitted, num_rest
        self.num_public = num_public; self.num_challenges = num_challenges
        self.num_cols = self.num_vars + 1 + num_public + num_challenges
        args = [C.c_uint64(x) for x in (num_cons, num_cons_unpadded, num_shared, num_precommitted, num_rest, num_public, num_challenges)]
        keep = []
        for (d, i, p) in (A, B, Cm):
            d = np.ascontiguousarray(d, dtype=np.uint64).reshape(-1, 4)
            if d.shape[0] == 0:
                d = np.zeros((1, 4), dtype=np.uint64)
            i = np.ascontiguousarray(i, dtype=np.uint32) if len(i) else np.zeros(1, dtype=np.uint32)
            p = np.ascontiguousarray(p, dtype=np.uint32)
            if p.shape[0] != num_cons + 1:
                raise SpartanError(-2, "indptr must have num_cons + 1 entries")
            keep += [d, i, p]
            args += [_p(d), _p(i), _p(p)]
        h = C.c_void_p()
        if nranks == 1:
            ctx.check(ctx.L.sp2_shape_upload(ctx.h, *args, C.byref(h)))
        else:
            ctx.check(ctx.L.sp2_shape_upload_sharded(ctx.h, C.c_int32(rank), C.c_int32(nranks), *args, C.byref(h)))
        self.h = h

    def free(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.sp2_shape_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def sizes(self):
        out = np.zeros(7, dtype=np.uint64)
        self.ctx.L.sp2_shape_sizes(self.h, _p(out))
        return dict(zip(["num_cons", "num_vars", "num_cols", "nnz", "nnz_general", "long_rows", "long_cols"], [int(x) for x in out]))

    def multiply_vec(self, z):
        z = _fe(z); n = self.num_cons // self.nranks     # a shard returns its rows i = rank (mod nranks)
        az, bz, cz = (np.zeros((n, 4), dtype=np.uint64) for _ in range(3))
        self.ctx.check(self.ctx.L.sp2_spmv3(self.ctx.h, self.h, _p(z), C.c_uint64(z.shape[0]), _p(az), _p(bz), _p(cz)))
        return az, bz, cz

    def multiply_vec_incremental(self, z, cached_az, cached_bz, cached_cz):
        z = _fe(z); n = self.num_cons // self.nranks
        ca, cb, cc = _fe(cached_az), _fe(cached_bz), _fe(cached_cz)
        az, bz, cz = (np.zeros((n, 4), dtype=np.uint64) for _ in range(3))
        self.ctx.check(self.ctx.L.sp2_spmv3_incremental(self.ctx.h, self.h, _p(z), C.c_uint64(z.shape[0]), _p(ca), _p(cb), _p(cc),
                                                        _p(az), _p(bz), _p(cz)))
        return az, bz, cz

    def bind_and_prepare_poly_ABC(self, evals_rx, r, full=False):
        rx = _fe(evals_rx); r = _fe(r)
        out_len = 2 * self.num_vars if full else self.num_cols
        if self.nranks > 1:                               # a shard returns its columns j = rank (mod nranks)
            out_len = (out_len - self.rank + self.nranks - 1) // self.nranks
        out = np.zeros((out_len, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.sp2_abc(self.ctx.h, self.h, _p(rx), C.c_uint64(rx.shape[0]), _p(r), _p(out), C.c_uint64(out_len)))
        return out


class CommitmentKey:
    """HyraxPCS commitment key resident on the device (sp2_ck): ck (n row bases), h, and the 1-wide evaluation
    key (ck_s, h_s) — src/provider/pcs/hyrax_pc.rs:152-190.  Points are (k, 8) u64 affine Montgomery."""

    def __init__(self, ctx, ck, h, ck_s, h_s):
        self.ctx = ctx
        ck = np.ascontiguousarray(ck, dtype=np.uint64).reshape(-1, 8)
        h, ck_s, h_s = (np.ascontiguousarray(x, dtype=np.uint64).reshape(1, 8) for x in (h, ck_s, h_s))
        self.n = ck.shape[0]
        hd = C.c_void_p()
        ctx.check(ctx.L.sp2_ck_upload(ctx.h, _p(ck), C.c_uint32(self.n), _p(h), _p(ck_s), _p(h_s), C.byref(hd)))
        self.h = hd

    def free(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.sp2_ck_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DlogGroupExt:
    """src/provider/traits.rs:118-162"""

    @staticmethod
    def vartime_multiscalar_mul(ctx, ck, scalars):
        s = _fe(scalars)
        out = np.zeros((1, 8), dtype=np.uint64)
        ss = s if s.shape[0] else np.zeros((1, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_msm(ctx.h, ck.h, _p(ss), C.c_uint32(s.shape[0]), _p(out)))
        return out


    # ---- arbitrary bases (the trait's actual signatures; Pippenger on the device) ----
    @staticmethod
    def vartime_multiscalar_mul_var(ctx, scalars, bases):
        s = _fe(scalars); b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        out = np.zeros((1, 8), dtype=np.uint64)
        if s.shape[0] != b.shape[0]:
            raise SpartanError(-2, "scalars and bases must have the same length")
        ctx.check(ctx.L.sp2_msm_var(ctx.h, _p(s if s.size else np.zeros((1, 4), dtype=np.uint64)), _p(b if b.size else np.zeros((1, 8), dtype=np.uint64)),
                                    C.c_uint32(s.shape[0]), _p(out)))
        return out

    @staticmethod
    def vartime_multiscalar_mul_small(ctx, scalars_u64, bases):
        s = np.ascontiguousarray(scalars_u64, dtype=np.uint64).reshape(-1); b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        out = np.zeros((1, 8), dtype=np.uint64)
        ctx.check(ctx.L.sp2_msm_small_var(ctx.h, _p(s if s.size else np.zeros(1, dtype=np.uint64)), _p(b if b.size else np.zeros((1, 8), dtype=np.uint64)),
                                          C.c_uint32(s.shape[0]), _p(out)))
        return out

    @staticmethod
    def batch_vartime_multiscalar_mul(ctx, scalar_vecs, bases):
        lens = np.array([len(v) for v in scalar_vecs], dtype=np.uint32)
        cat = np.concatenate([_fe(v) for v in scalar_vecs if len(v)] or [np.zeros((1, 4), dtype=np.uint64)])
        b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        out = np.zeros((len(scalar_vecs), 8), dtype=np.uint64)
        ctx.check(ctx.L.sp2_msm_batch_var(ctx.h, _p(cat), _p(lens), C.c_uint32(len(scalar_vecs)), _p(b), _p(out)))
        return out

    @staticmethod
    def vartime_multiscalar_mul_shared_weights(ctx, weights, bases_rows):
        w = _fe(weights); b = np.ascontiguousarray(bases_rows, dtype=np.uint64).reshape(-1, w.shape[0], 8)
        out = np.zeros((b.shape[0], 8), dtype=np.uint64)
        ctx.check(ctx.L.sp2_msm_shared_weights(ctx.h, _p(w), C.c_uint32(w.shape[0]), _p(b), C.c_uint32(b.shape[0]), _p(out)))
        return out


class HyraxPCS:
    """src/provider/pcs/hyrax_pc.rs"""

    @staticmethod
    def commit(ctx, ck, v, blinds, is_small=False):
        v = _fe(v); blinds = _fe(blinds); rows = blinds.shape[0]
        out = np.zeros((rows, 8), dtype=np.uint64)
        vv = v if v.shape[0] else np.zeros((1, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_hyrax_commit(ctx.h, ck.h, _p(vv), C.c_uint64(v.shape[0]), _p(blinds), C.c_uint64(rows), C.c_int32(int(is_small)), _p(out)))
        return out

    @staticmethod
    def commit_without_blind(ctx, ck, v, is_small=False):
        v = _fe(v); rows = (v.shape[0] + ck.n - 1) // ck.n
        out = np.zeros((rows, 8), dtype=np.uint64)
        ctx.check(ctx.L.sp2_hyrax_commit_without_blind(ctx.h, ck.h, _p(v if v.size else np.zeros((1, 4), dtype=np.uint64)), C.c_uint64(v.shape[0]),
                                                       C.c_int32(int(is_small)), _p(out)))
        return out

    @staticmethod
    def commit_incremental(ctx, ck, raw, delta, blinds):
        raw = np.ascontiguousarray(raw, dtype=np.uint64).reshape(-1, 8); delta = _fe(delta); blinds = _fe(blinds)
        rows = (delta.shape[0] + ck.n - 1) // ck.n
        out = np.zeros((rows, 8), dtype=np.uint64)
        ctx.check(ctx.L.sp2_hyrax_commit_incremental(ctx.h, ck.h, _p(raw if raw.size else np.zeros((1, 8), dtype=np.uint64)), C.c_uint64(raw.shape[0]),
                                                     _p(delta), C.c_uint64(delta.shape[0]), _p(blinds), _p(out)))
        return out

    @staticmethod
    def rerandomize_commitment(ctx, ck, comm, r_old, r_new):
        comm = np.ascontiguousarray(comm, dtype=np.uint64).reshape(-1, 8); r_old = _fe(r_old); r_new = _fe(r_new)
        if not (comm.shape[0] == r_old.shape[0] == r_new.shape[0]):
            raise SpartanError(-2, "rerandomize_commitment: commitment and blinds must have the same length")
        out = np.zeros_like(comm)
        ctx.check(ctx.L.sp2_hyrax_rerandomize(ctx.h, ck.h, _p(comm), _p(r_old), _p(r_new), C.c_uint64(comm.shape[0]), _p(out)))
        return out

    @staticmethod
    def fold_blinds(ctx, blinds, n, rows, w):
        blinds = _fe(blinds); w = _fe(w); out = np.zeros((rows, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_fold_blinds(ctx.h, _p(blinds), C.c_uint32(n), C.c_uint32(rows), _p(w), _p(out)))
        return out

    @staticmethod
    def fold_commitments_partial(ctx, ck, comms, n, rows, w, num_data_rows, folded_blind):
        comms = np.ascontiguousarray(comms, dtype=np.uint64).reshape(-1, 8); w = _fe(w); fb = _fe(folded_blind)
        out = np.zeros((rows, 8), dtype=np.uint64)
        ctx.check(ctx.L.sp2_fold_commitments_partial(ctx.h, ck.h, _p(comms), C.c_uint32(n), C.c_uint32(rows), _p(w), C.c_uint32(num_data_rows), _p(fb), _p(out)))
        return out

    @staticmethod
    def bind_with_delayed(ctx, poly, L, r_len):
        poly = _fe(poly); L = _fe(L)
        out = np.zeros((r_len, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_hyrax_bind(ctx.h, _p(poly), _p(L), C.c_uint64(L.shape[0]), C.c_uint64(r_len), _p(out)))
        return out


class _ProofC(C.Structure):
    _fields_ = [("num_rounds_x", C.c_uint64), ("num_rounds_y", C.c_uint64), ("num_comm_rows", C.c_uint64), ("num_cols", C.c_uint64)] + \
               [(n, C.c_void_p) for n in ("comm_W", "outer_polys", "claims_outer", "inner_polys", "eval_W", "blind_eval_W", "delta", "beta",
                                          "z_vec", "z_delta", "z_beta")]


class _RandC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("blinds_W", "blind_eval_W", "d_vec", "r_delta", "r_beta")]


class SpartanProof:
    """Flat SpartanSNARK proof (include/spartan2_b200.h: sp2_spartan_proof)."""
    FIELDS = ["comm_W", "outer_polys", "claims_outer", "inner_polys", "eval_W", "blind_eval_W", "delta", "beta", "z_vec", "z_delta", "z_beta"]

    _layout_cache = {}

    def __init__(self, l, nry, rows, num_cols):
        # one backing buffer; the fields are views made on first access (the wrapper's per-call overhead counts in the end-to-end time
        # of a ~1.3 ms prove).  The C side writes every field of a successful prove.
        self.l, self.nry, self.rows, self.num_cols = l, nry, rows, num_cols
        key = (l, nry, rows, num_cols)
        lay = SpartanProof._layout_cache.get(key)
        if lay is None:
            shapes = ((rows, 8), (3 * l, 4), (3, 4), (2 * nry, 4), (1, 4), (1, 4), (1, 8), (1, 8), (num_cols, 4), (1, 4), (1, 4))
            offs, o = {}, 0
            for f, sh in zip(self.FIELDS, shapes):
                offs[f] = (o, sh); o += sh[0] * sh[1]
            lay = SpartanProof._layout_cache[key] = (offs, o)
        self._offs, total = lay
        self._buf = np.zeros(total, dtype=np.uint64)
        self._phase = None

    def __getattr__(self, name):            # only called for attributes not set yet: the field views, phase_ms
        offs = self.__dict__.get("_offs")
        if offs is not None and name in offs:
            o, sh = offs[name]
            v = self._buf[o:o + sh[0] * sh[1]].reshape(sh)
            self.__dict__[name] = v
            return v
        if name == "phase_ms":
            ph = self.__dict__.get("_phase")
            d = None if ph is None else dict(zip(PHASES, [float(x) for x in ph]))
            self.__dict__["phase_ms"] = d
            return d
        raise AttributeError(name)

    def cview(self):
        base = _addr(self._buf)
        offs = self._offs
        return _ProofC(self.l, self.nry, self.rows, self.num_cols, *[base + 8 * offs[f][0] for f in self.FIELDS])


PHASES = ["commit_transcript", "matrix_vector_multiply", "outer_sumcheck", "prepare_poly_ABC", "inner_sumcheck", "pcs_prove", "ipa_response", "total"]


class SpartanPrepSNARK:
    def __init__(self, ctx, h, comm):
        self.ctx, self.h, self.comm = ctx, h, comm

    def free(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.sp2_prep_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SpartanSNARK:
    """src/spartan.rs — prep_prove / prove / verify on the device (setup = SplitR1CSShape + CommitmentKey uploads)."""

    @staticmethod
    def verify(ctx, shape, ck, vk_digest, public_values, proof):
        return spartan_verify(ctx, shape, ck, vk_digest, public_values, proof)

    @staticmethod
    def prep_prove(ctx, shape, ck, W_cached, blinds_cached, is_small=True):
        W = _fe(W_cached); b = _fe(blinds_cached)
        cached_len = shape.num_shared + shape.num_precommitted
        if W.shape[0] != cached_len:
            raise SpartanError(-3, "prep_prove expects shared + precommitted (%d) witness values" % cached_len)
        rows = cached_len // ck.n
        if b.shape[0] < rows:
            raise SpartanError(-2, "one blind per commitment row")
        comm = np.zeros((max(rows, 1), 8), dtype=np.uint64)
        h = C.c_void_p()
        Wp = W if W.shape[0] else np.zeros((1, 4), dtype=np.uint64)
        bp = b if b.shape[0] else np.zeros((1, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_spartan_prep_prove(ctx.h, shape.h, ck.h, _p(Wp), _p(bp), C.c_int32(int(is_small)), _p(comm), C.byref(h)))
        return SpartanPrepSNARK(ctx, h, comm[:rows])

    @staticmethod
    def prove(ctx, shape, ck, prep, vk_digest, public_values, W_rest, blinds_W, blind_eval_W, d_vec, r_delta, r_beta, comm=None):
        """comm: a connected Comm when `shape` is a multi-GPU shard (every rank calls with the same inputs and gets the same proof)."""
        N, nv = shape.num_cons, shape.num_vars
        l = N.bit_length() - 1; nry = nv.bit_length()
        rows = nv // ck.n
        P = SpartanProof(l, nry, rows, ck.n)
        pv = P.cview()
        arrs = [_fe(x) for x in (blinds_W, blind_eval_W, d_vec, r_delta, r_beta)]
        rv = _RandC(*[_addr(a) for a in arrs])
        dig = bytes(vk_digest)                                                # (read-only: passed as a pointer to the bytes object's buffer)
        pub = _fe(public_values) if len(public_values) else np.zeros((1, 4), dtype=np.uint64)
        Wr = _fe(W_rest) if W_rest is not None and len(W_rest) else None      # None: all-zero rest section
        ph = (C.c_float * 8)()
        tail = (shape.h, ck.h, prep.h, dig, pub.ctypes.data, Wr.ctypes.data if Wr is not None else None, C.byref(rv), C.byref(pv), ph)
        if comm is None:
            ctx.check(ctx.L.sp2_spartan_prove(ctx.h, *tail))
        else:
            ctx.check(ctx.L.sp2_spartan_prove_sharded(ctx.h, comm.h, *tail))
        P._phase = ph                                                         # (phase_ms: a dict made on first access)
        return P


def spartan_verify(ctx, shape, ck, vk_digest, public_values, proof):
    """SpartanSNARK::verify on the device (sp2_spartan_verify): returns None on accept, raises SpartanError(kind ProofVerifyError) on reject."""
    pv = proof.cview()
    dig = np.frombuffer(bytes(vk_digest), dtype=np.uint8).copy()
    pub = _fe(public_values) if len(public_values) else np.zeros((1, 4), dtype=np.uint64)
    ctx.check(ctx.L.sp2_spartan_verify(ctx.h, shape.h, ck.h, _p(dig), _p(pub), C.byref(pv)))


def shard_cyclic(table, nranks, rank):
    """This rank's shard of a table split cyclically on the low index bits (entries i = rank mod nranks)."""
    return np.ascontiguousarray(np.asarray(table)[rank::nranks])


class Comm:
    """Peer mailboxes for the sharded sum-checks (sp2_comm).  `allgather_bytes(b: bytes) -> list[bytes]` is the host's
    collective (e.g. built on torch.distributed.all_gather_object); it is only used once, to exchange IPC handles."""

    def __init__(self, ctx, rank, nranks, allgather_bytes=None):
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        h = C.c_void_p()
        ctx.check(ctx.L.sp2_comm_create(ctx.h, C.c_int32(rank), C.c_int32(nranks), C.byref(h)))
        self.h = h
        if nranks > 1 and allgather_bytes is not None:
            buf = (C.c_uint8 * 64)()
            ctx.check(ctx.L.sp2_comm_handle(self.h, buf))
            handles = allgather_bytes(bytes(buf))
            assert len(handles) == nranks and all(len(x) == 64 for x in handles)
            allb = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
            ctx.check(ctx.L.sp2_comm_connect(self.h, _p(allb)))

    @staticmethod
    def in_process(ctxs):
        """One comm per context, all in THIS process (several contexts on one GPU, or on peer-enabled GPUs): the mailboxes are
        exchanged as raw device pointers (sp2_comm_connect_ptrs), no CUDA IPC.  Rank q = ctxs[q]."""
        n = len(ctxs)
        comms = [Comm(c, q, n) for q, c in enumerate(ctxs)]
        ptrs = (C.c_void_p * n)()
        for q, cm in enumerate(comms):
            p = C.c_void_p()
            cm.ctx.check(cm.ctx.L.sp2_comm_mailbox(cm.h, C.byref(p)))
            ptrs[q] = p.value
        for cm in comms:
            cm.ctx.check(cm.ctx.L.sp2_comm_connect_ptrs(cm.h, ptrs))
        return comms

    def reset(self):
        """Collective (every rank, host barrier before and after): flags / epochs back to zero after a failed sharded call."""
        self.ctx.check(self.ctx.L.sp2_comm_reset(self.h))

    def free(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.sp2_comm_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def prove_cubic_with_three_inputs(self, claim, taus, A, B, Cz, ts):
        """A, B, Cz: DeviceBuffers holding this rank's cyclic shards."""
        ctx = self.ctx
        taus = _fe(taus); l = taus.shape[0]; claim = _fe(claim)
        polys = np.zeros((l, 4, 4), dtype=np.uint64); r = np.zeros((l, 4), dtype=np.uint64); claims = np.zeros((3, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_sumcheck_cubic_prove_sharded_dev(ctx.h, self.h, _p(claim), _p(taus), C.c_uint32(l), A.ptr, B.ptr, Cz.ptr,
                                                             C.byref(ts), _p(polys), _p(r), _p(claims)))
        return polys, r, claims

    def prove_quad(self, claim, num_rounds, A, B, ts):
        ctx = self.ctx
        claim = _fe(claim); l = int(num_rounds)
        polys = np.zeros((l, 3, 4), dtype=np.uint64); r = np.zeros((l, 4), dtype=np.uint64); claims = np.zeros((2, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_sumcheck_quad_prove_sharded_dev(ctx.h, self.h, _p(claim), C.c_uint32(l), A.ptr, B.ptr, C.byref(ts), _p(polys),
                                                            _p(r), _p(claims)))
        return polys, r, claims


class Keccak256Transcript:
    """src/provider/keccak.rs:18-105 — the host transcript of the library (pure host code: usable without a GPU).
    Scalars are (n, 4) u64 Montgomery limbs; squeeze returns the challenge as a (1, 4) Montgomery scalar."""

    def __init__(self, label):
        from ._lib import lib
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.sp2_transcript_new(bytes(label), C.byref(h))
        if rc != 0:
            raise SpartanError(rc, "transcript")
        self.h = h

    def absorb_bytes(self, label, data):
        d = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
        self.L.sp2_transcript_absorb_bytes(self.h, bytes(label), _p(d), C.c_uint64(len(data)))

    def absorb_scalars(self, label, scalars):
        a = _fe(scalars)
        self.L.sp2_transcript_absorb_scalars(self.h, bytes(label), _p(a if a.shape[0] else np.zeros((1, 4), dtype=np.uint64)), C.c_uint64(a.shape[0]))

    def absorb_commitment(self, label, rows_xy):
        a = np.ascontiguousarray(rows_xy, dtype=np.uint64).reshape(-1, 8)
        self.L.sp2_transcript_absorb_commitment(self.h, bytes(label), _p(a), C.c_uint64(a.shape[0]))

    def dom_sep(self, label):
        self.L.sp2_transcript_dom_sep(self.h, bytes(label))

    def squeeze(self, label):
        out = np.zeros((1, 4), dtype=np.uint64)
        self.L.sp2_transcript_squeeze(self.h, bytes(label), _p(out))
        return out

    def state(self):
        from ._lib import TranscriptState
        t = TranscriptState()
        self.L.sp2_transcript_get_state(self.h, C.byref(t))
        return t

    def __del__(self):
        try:
            if self.h:
                self.L.sp2_transcript_free(self.h); self.h = None
        except Exception:
            pass


class PowPolynomial:
    """src/polys/power.rs"""

    @staticmethod
    def split_evals(ctx, t, left, right):
        t = _fe(t); out = np.zeros((left + right, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_pow_split_evals(ctx.h, _p(t), C.c_uint32(left), C.c_uint32(right), _p(out)))
        return out


class NeutronNovaNIFS:
    """src/neutronnova_zk.rs:511-1273 — the per-round kernels of the multi-folding scheme on device-resident layers.
    A, B, C: DeviceBuffers of n*N scalars, layer-major.  The challenges r_b come from the caller (the reference draws
    them through its in-circuit verifier)."""

    def __init__(self, ctx, E, left, right, A, B, Cz, n):
        self.ctx, self.left, self.right, self.N, self.n = ctx, left, right, left * right, n
        self.dE = ctx.upload(_fe(E)); self.A, self.B, self.C = A, B, Cz
        self.m = n; self.stride = 1; self.t = 0

    def round_eval(self, rhos):
        rhos = _fe(rhos); out = np.zeros((2, 4), dtype=np.uint64); ctx = self.ctx
        ctx.check(ctx.L.sp2_nifs_round_dev(ctx.h, C.c_uint32(self.t), _p(rhos), C.c_uint32(rhos.shape[0]), C.c_uint32(self.left), C.c_uint32(self.right),
                                           self.dE.ptr, self.A.ptr, self.B.ptr, self.C.ptr, C.c_uint64(self.N), C.c_uint64(self.m), C.c_uint64(self.stride), _p(out)))
        return out

    def fold(self, r_b):
        r_b = _fe(r_b); ctx = self.ctx
        ctx.check(ctx.L.sp2_nifs_fold_dev(ctx.h, self.A.ptr, self.B.ptr, self.C.ptr, C.c_uint64(self.N), C.c_uint64(self.m), C.c_uint64(self.stride), _p(r_b)))
        self.m //= 2; self.stride *= 2; self.t += 1

    def layer0(self):
        """the folded (Az, Bz, Cz) layers once m == 1"""
        return tuple(b.download((self.N, 4)) for b in (self.A, self.B, self.C))


class SmallValue:
    """src/big_num/small_value.rs on device-resident layers (i64 tables, SmallAccumulator sums) and the NIFS kernels that
    consume them (src/neutronnova_zk.rs:255-325, 649-693, 1551-1584)."""

    @staticmethod
    def to_small_layers(ctx, tables, n_layers, N):
        """tables: DeviceBuffers of n_layers*N scalars -> (i64 DeviceBuffers, positions DeviceBuffer (u64, ascending), n_large)"""
        outs = [ctx.alloc(n_layers * N * 8) for _ in tables]
        pos = ctx.alloc(max(N, 1) * 8)
        tin = (C.c_void_p * len(tables))(*[t.ptr.value for t in tables]); tout = (C.c_void_p * len(tables))(*[t.ptr.value for t in outs])
        nl = C.c_uint64()
        ctx.check(ctx.L.sp2_to_small_layers_dev(ctx.h, tin, tout, C.c_uint32(len(tables)), C.c_uint64(n_layers), C.c_uint64(N), pos.ptr, C.byref(nl)))
        return outs, pos, int(nl.value)

    @staticmethod
    def nifs_round0(ctx, rhos, left, right, dE, dA64, dB64, dA, dB, dpos, n_large, N, m):
        rhos = _fe(rhos); out = np.zeros((2, 4), dtype=np.uint64)
        rr = rhos if rhos.shape[0] else np.zeros((1, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_nifs_round0_small_dev(ctx.h, _p(rr), C.c_uint32(rhos.shape[0]), C.c_uint32(left), C.c_uint32(right), dE.ptr, dA64.ptr, dB64.ptr,
                                                  dA.ptr, dB.ptr, dpos.ptr, C.c_uint64(n_large), C.c_uint64(N), C.c_uint64(m), _p(out)))
        return out

    @staticmethod
    def cvals(ctx, left, right, dE, dC, dC64, dpos, n_large, N, n):
        out = np.zeros((n, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_nifs_cvals_small_dev(ctx.h, C.c_uint32(left), C.c_uint32(right), dE.ptr, dC.ptr, dC64.ptr, dpos.ptr, C.c_uint64(n_large),
                                                 C.c_uint64(N), C.c_uint64(n), _p(out)))
        return out


def weights_from_r(ctx, r_bs, n):
    r_bs = _fe(r_bs); out = np.zeros((n, 4), dtype=np.uint64)
    ctx.check(ctx.L.sp2_weights_from_r(ctx.h, _p(r_bs), C.c_uint32(r_bs.shape[0]), C.c_uint32(n), _p(out)))
    return out


class R1CSWitness:
    """src/r1cs/mod.rs:540-660"""

    @staticmethod
    def fold_multiple(ctx, r_bs, Ws):
        """Ws: (n, dim, 4) witness vectors; returns the folded W (dim, 4)."""
        Ws = np.ascontiguousarray(Ws, dtype=np.uint64); n, dim = Ws.shape[0], Ws.shape[1]
        w = weights_from_r(ctx, r_bs, n)
        dW = ctx.upload(Ws); out = ctx.alloc(dim * 32)
        ctx.check(ctx.L.sp2_fold_vectors_dev(ctx.h, dW.ptr, C.c_uint64(n), C.c_uint64(dim), _p(w), out.ptr))
        return out.download((dim, 4))


class SumcheckRounds:
    """Per-round evaluation points + binds of the ZK sum-check drivers (src/sumcheck.rs:702-917)."""

    @staticmethod
    def eval_points_cubic_with_outer_pow(ctx, d_pow_left, left, d_pow_right, dA, dB, dC, table_len):
        out = np.zeros((3, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_sc_pow_cubic_eval_dev(ctx.h, d_pow_left.ptr, C.c_uint32(left), d_pow_right.ptr, dA.ptr, dB.ptr, dC.ptr, C.c_uint64(table_len), _p(out)))
        return out

    @staticmethod
    def eval_points_quad(ctx, dA, dB, table_len):
        out = np.zeros((2, 4), dtype=np.uint64)
        ctx.check(ctx.L.sp2_sc_quad_eval_dev(ctx.h, dA.ptr, dB.ptr, C.c_uint64(table_len), _p(out)))
        return out

    @staticmethod
    def bind_poly_var_top(ctx, tables, table_len, r):
        r = _fe(r); arr = (C.c_void_p * len(tables))(*[t.ptr.value for t in tables])
        ctx.check(ctx.L.sp2_bind_tables_dev(ctx.h, arr, C.c_uint32(len(tables)), C.c_uint64(table_len), _p(r)))


def fold_commitments(ctx, comms, n, rows, w):
    """HyraxPCS::fold_commitments: comms (n*rows, 8) affine, w (n, 4) -> (rows, 8)."""
    comms = np.ascontiguousarray(comms, dtype=np.uint64).reshape(-1, 8); w = _fe(w)
    out = np.zeros((rows, 8), dtype=np.uint64)
    ctx.check(ctx.L.sp2_fold_commitments(ctx.h, _p(comms), C.c_uint32(n), C.c_uint32(rows), _p(w), _p(out)))
    return out
