"""The oracle's side of spartan2_b200.neutronnova.run: the same bulk operations on numpy tables through oracle/ (CPU),
so the parity test runs ONE driver over two backends and compares every recorded intermediate value."""
import numpy as np

from oracle import pyoracle as orc


class Tab:
    def __init__(self, a):
        self.a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)


class _Nifs:
    def __init__(self, E, left, right, A, B, Cz, n):
        self.E, self.left, self.right, self.N = E, left, right, left * right
        self.A, self.B, self.C, self.m, self.t = A, B, Cz, n, 0

    def round_eval(self, rhos):
        return orc.nifs_round(self.t, rhos, self.left, self.right, self.E, self.A.a, self.B.a, self.C.a, self.N, self.m)

    def fold(self, r_b):
        for T in (self.A, self.B, self.C):
            T.a = orc.nifs_fold(T.a, self.N, self.m, r_b)
        self.m //= 2; self.t += 1


class OracleOps:
    def __init__(self, shape, dims):
        self.O = shape
        self.N, self.M = dims[0], dims[2] + dims[3] + dims[4]
        self.ncols = self.M + 1 + dims[5] + dims[6]

    def spmv_layers(self, zs):
        outs = [self.O.multiply_vec(z) for z in zs]
        return tuple(Tab(np.concatenate([o[k] for o in outs], axis=0)) for k in range(3))

    def pow_split(self, tau, left, right):
        return orc.pow_split_evals(tau, left, right)

    def nifs(self, E, left, right, A, B, Cz, n):
        return _Nifs(E, left, right, A, B, Cz, n)

    def nifs_result(self, nifs):
        return nifs.A, nifs.B, nifs.C

    def _z(self, W):
        z = np.zeros((2 * self.M, 4), dtype=np.uint64)
        z[: W.shape[0]] = W; z[self.M] = orc.to_mont([1])[0]
        return Tab(z)

    def fold_witness(self, r_bs, Ws):
        Ws = np.ascontiguousarray(Ws, dtype=np.uint64); n, dim = Ws.shape[0], Ws.shape[1]
        return self._z(orc.fold_vectors(Ws.reshape(-1, 4), n, dim, orc.weights_from_r(r_bs, n)))

    def z_table(self, W):
        return self._z(np.ascontiguousarray(W, dtype=np.uint64).reshape(-1, 4))

    def tables(self, arr):
        return Tab(arr)

    def pow_cubic_eval(self, pl, left, pr, A, B, Cz, tl):
        return orc.pow_cubic_eval(pl.a, pr.a, A.a[:tl], B.a[:tl], Cz.a[:tl])

    def quad_eval(self, A, B, tl):
        return orc.quad_eval(A.a[:tl], B.a[:tl])

    def bind(self, tables, tl, r):
        for T in tables:
            T.a = orc.bind_top(T.a[:tl], r)

    def head(self, table, k=1):
        return table.a[:k].copy()

    def eq_table(self, r_x):
        return Tab(orc.eq_evals(r_x))

    def abc_full(self, rx, r):
        out = np.zeros((2 * self.M, 4), dtype=np.uint64)
        v = self.O.abc(rx.a, r)
        out[: v.shape[0]] = v
        return Tab(out)

    def sync(self):
        pass


def sha_chain_instances(n):
    """benches/sha256_neutronnova.rs:161-178, 219-223: step i hashes the block [i as u8; 64]; the core circuit one zero block."""
    from spartan2_b200.frontend import Sha256Circuit
    one = orc.to_mont([1])
    steps = [Sha256Circuit(bytes([i % 256]) * 64, kind="compression") for i in range(n)]
    core = Sha256Circuit(bytes(64), kind="compression")

    def z_of(c):
        W, X = c.witness()
        return np.concatenate([W, one, X], axis=0), W
    zs, Ws = zip(*[z_of(c) for c in steps])
    zc, Wc = z_of(core)
    return steps[0], list(zs), np.stack(Ws), zc, Wc
