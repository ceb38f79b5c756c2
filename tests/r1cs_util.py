"""Seeded synthetic R1CS instances with SHA-256-like statistics: mostly +-1
coefficients, some small integers, some powers of two, a few long rows, a heavily used constant-one
column; boolean witness.  Used by CPU and GPU tests alike (numpy only)."""
import numpy as np

Q = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
R = 1 << 256


def mont(v):
    v = (v % Q) * R % Q
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def random_r1cs(seed, log_cons, log_vars, num_public=2, num_shared=0, rest_frac=0.0, width=None, satisfiable=True, dense_rows=3):
    """Returns dict(dims..., A, B, C as (data, indices, indptr), W (num_vars,4) Montgomery, X (num_public,4)).
    Constraints are (sum a_j z_j) * (sum b_j z_j) = c-row; when `satisfiable` each C row is a single fresh
    witness variable (or the product is forced to match), so Az∘Bz = Cz holds and the outer claim is 0."""
    rng = np.random.default_rng(seed)
    N = 1 << log_cons; M = 1 << log_vars
    width = width or min(M, 2048)
    num_vars = M
    num_rest = int(rest_frac * M) // width * width
    num_pre = M - num_shared - num_rest
    ncols = num_vars + 1 + num_public
    one_col = num_vars
    coef_pool = [1, 1, 1, 1, -1, -1, 2, -2, 3, 4, 5, 7, 8, 16, 1 << 20, 1 << 31, (1 << 40) + 5, -(1 << 17)]
    coef_mont = {c: mont(c) for c in set(coef_pool)}
    # witness: bits, the defined "product" variables are filled in below
    n_unpadded = N - N // 8 if N >= 16 else N      # leave padded (empty) rows at the end
    free_vars = M - n_unpadded if M > n_unpadded else max(1, M // 2)
    w_int = [int(b) for b in rng.integers(0, 2, size=M)]
    x_int = [int(v) for v in rng.integers(0, 1 << 30, size=num_public)]

    def z_val(col):
        if col < num_vars:
            return w_int[col]
        if col == one_col:
            return 1
        return x_int[col - one_col - 1]

    mats = {k: ([], [], [0]) for k in "ABC"}
    for row in range(N):
        if row < n_unpadded:
            sums = {}
            for k in "AB":
                nterm = int(rng.integers(1, 4))
                if dense_rows and row % max(1, n_unpadded // dense_rows) == 1:
                    nterm = int(rng.integers(60, 200))
                cols = rng.choice(min(free_vars, M), size=min(nterm, min(free_vars, M)), replace=False).tolist()
                if rng.random() < 0.4:
                    cols.append(one_col)
                if num_public and rng.random() < 0.05:
                    cols.append(one_col + 1 + int(rng.integers(0, num_public)))
                cols = sorted(set(cols))
                s = 0
                for c in cols:
                    cf = coef_pool[int(rng.integers(0, len(coef_pool)))]
                    mats[k][0].append(coef_mont[cf]); mats[k][1].append(c)
                    s += cf * z_val(c)
                mats[k][2].append(len(mats[k][1]))
                sums[k] = s % Q
            prod = sums["A"] * sums["B"] % Q
            if satisfiable and M > n_unpadded:
                tgt = free_vars + row if free_vars + row < M else None
            else:
                tgt = None
            if tgt is not None:
                w_int[tgt] = prod
                mats["C"][0].append(coef_mont[1]); mats["C"][1].append(tgt)
            else:
                # C row = prod * ONE (general coefficient on the constant column)
                mats["C"][0].append(mont(prod)); mats["C"][1].append(one_col)
            mats["C"][2].append(len(mats["C"][1]))
        else:
            for k in "ABC":
                mats[k][2].append(len(mats[k][1]))
    out = {}
    for k in "ABC":
        d, i, p = mats[k]
        out[k] = (np.array(d, dtype=np.uint64).reshape(-1, 4), np.array(i, dtype=np.uint32), np.array(p, dtype=np.uint32))
    W = np.array([mont(v) for v in w_int], dtype=np.uint64)
    X = np.array([mont(v) for v in x_int], dtype=np.uint64).reshape(-1, 4)
    out.update(num_cons=N, num_cons_unpadded=n_unpadded, num_shared=num_shared, num_precommitted=num_pre, num_rest=num_rest,
               num_public=num_public, num_challenges=0, W=W, X=X, num_vars=num_vars, ncols=ncols)
    return out


def z_of(inst):
    one = np.array([mont(1)], dtype=np.uint64)
    return np.concatenate([inst["W"], one, inst["X"]]) if inst["num_public"] else np.concatenate([inst["W"], one])


def dims(inst):
    return (inst["num_cons"], inst["num_cons_unpadded"], inst["num_shared"], inst["num_precommitted"], inst["num_rest"],
            inst["num_public"], inst["num_challenges"])


def chain_instances(n, log_cons, log_vars, num_public=2, seed=11):
    """n satisfying instances of ONE random R1CS shape (the NeutronNova multi-folding setting: step circuits share their
    shape, witnesses differ) plus a core instance, shaped like the SHA-256 chain: precommitted section = first half of
    the variables, rest section = second half, all zero and unused (the reference's step circuits allocate nothing in
    `synthesize`, so their rest rows are commit_zeros, bellpepper/r1cs.rs:443-470).  Returns (dims, (A, B, C), zs, zc)."""
    assert log_vars - 1 > log_cons - 1, "need more variables than constraints so every C row is a fresh product variable"
    base = random_r1cs(seed, log_cons, log_vars - 1, num_public=num_public, rest_frac=0.0, dense_rows=2)
    Mh = 1 << (log_vars - 1); M = 2 * Mh
    N, n_unp = base["num_cons"], base["num_cons_unpadded"]
    free_vars = Mh - n_unp

    def shift(mat):
        d, i, p = mat
        i = i.copy(); i[i >= Mh] += Mh                      # the constant-one and public columns move behind the rest section
        return d, i, p
    mats = tuple(shift(base[k]) for k in "ABC")
    Ai, Bi = [from_mont_rows(m) for m in mats[:2]]

    def witness(s):
        rng = np.random.default_rng(1000 * seed + s)
        w = [int(b) for b in rng.integers(0, 2, size=Mh)]
        x = [int(v) for v in rng.integers(0, 1 << 30, size=num_public)]
        z = lambda c: w[c] if c < Mh else (1 if c == M else x[c - M - 1])      # noqa: E731
        for row in range(n_unp):
            sa = sum(cf * z(c) for cf, c in Ai[row]) % Q; sb = sum(cf * z(c) for cf, c in Bi[row]) % Q
            w[free_vars + row] = sa * sb % Q
        W = np.array([mont(v) for v in w] + [mont(0)] * Mh, dtype=np.uint64)
        X = np.array([mont(v) for v in x], dtype=np.uint64).reshape(-1, 4)
        one = np.array([mont(1)], dtype=np.uint64)
        return np.concatenate([W, one, X]) if num_public else np.concatenate([W, one])
    dims_ = (N, n_unp, 0, Mh, Mh, num_public, 0)
    return dims_, mats, [witness(s) for s in range(n)], witness(n)


def from_mont_rows(mat):
    """CSR (data Montgomery, indices, indptr) -> per-row lists of (canonical coefficient, column)."""
    d, i, p = mat
    rinv = pow(R, -1, Q)
    vals = [sum(int(l) << (64 * k) for k, l in enumerate(row)) * rinv % Q for row in d]
    return [[(vals[k], int(i[k])) for k in range(int(p[r]), int(p[r + 1]))] for r in range(len(p) - 1)]
