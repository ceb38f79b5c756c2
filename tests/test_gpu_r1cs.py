"""R1CS sparse kernels (Az/Bz/Cz, incremental SpMV, M^T*eq builder) vs the oracle, through the C ABI.
Includes the reference's own SpMV known answer [25, 9, 4] (src/r1cs/sparse.rs:637-653)."""
import numpy as np
import pytest

from tests.gpu_util import ctx, rand_fe  # noqa: F401
from tests.r1cs_util import dims, mont, random_r1cs, z_of

pytestmark = pytest.mark.gpu


def _shapes(ctx, orc, inst):
    import spartan2_b200 as sp
    S = sp.SplitR1CSShape(ctx, *dims(inst), inst["A"], inst["B"], inst["C"])
    O = orc.Shape(*dims(inst), inst["A"], inst["B"], inst["C"])
    return S, O


def test_reference_spmv_known_answer(ctx):
    # the reference's own vector (src/r1cs/sparse.rs:637-653, test_matrix_vector_multiplication), read from the committed
    # fixture tests/golden/reference_kats.json: [[0,2,7],[0,0,3],[4,0,0]] * [1,2,3] = [25,9,4]
    import json
    import os
    import spartan2_b200 as sp
    k = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))["spmv"]
    assert k["matrix"] == [[0, 2, 7], [0, 0, 3], [4, 0, 0]] and k["z"] == [1, 2, 3] and k["out"] == [25, 9, 4]
    data, idx, ptr = [], [], [0]
    for row in k["matrix"]:
        for j, v in enumerate(row):
            if v:
                data.append(mont(v)); idx.append(j)
        ptr.append(len(idx))
    ptr.append(len(idx))                                     # padded to 4 rows
    data = np.array(data, dtype=np.uint64); idx = np.array(idx, dtype=np.uint32); ptr = np.array(ptr, dtype=np.uint32)
    empty = (np.zeros((0, 4), dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros(5, dtype=np.uint32))
    # columns: 2 witness vars + the constant one => z = (1, 2 | 3)
    S = sp.SplitR1CSShape(ctx, 4, 3, 0, 2, 0, 0, 0, (data, idx, ptr), empty, empty)
    z = np.array([mont(v) for v in k["z"]], dtype=np.uint64)
    az, bz, cz = S.multiply_vec(z)
    assert az.tolist() == [mont(v) for v in k["out"]] + [mont(0)]
    assert not bz.any() and not cz.any()


@pytest.mark.parametrize("lc,lv,seed", [(4, 4, 1), (8, 7, 2), (10, 11, 3), (13, 13, 4), (16, 15, 5)])
def test_multiply_vec_and_abc(ctx, orc, lc, lv, seed):
    inst = random_r1cs(seed, lc, lv, num_public=3, rest_frac=0.25)
    S, O = _shapes(ctx, orc, inst)
    z = z_of(inst)
    az, bz, cz = S.multiply_vec(z)
    oa, ob, oc = O.multiply_vec(z)
    assert np.array_equal(az, oa) and np.array_equal(bz, ob) and np.array_equal(cz, oc)
    # the instance is satisfiable: Az∘Bz = Cz
    assert np.array_equal(orc.f_mul(az, bz), cz)
    rng = np.random.default_rng(seed)
    rx = orc.eq_evals(rand_fe(rng, lc)); r = rand_fe(rng, 1)
    got = S.bind_and_prepare_poly_ABC(rx, r)
    assert np.array_equal(got, O.abc(rx, r))
    full = S.bind_and_prepare_poly_ABC(rx, r, full=True)
    assert np.array_equal(full[: got.shape[0]], got) and not full[got.shape[0]:].any()
    sz = S.sizes()
    assert sz["num_cons"] == 1 << lc and sz["nnz"] == sum(len(inst[k][1]) for k in "ABC")


def test_incremental_matches_full(ctx, orc):
    # multiply_vec_incremental_into (mod.rs:1170-1211): cached product over shared+precommitted columns, then the rest
    inst = random_r1cs(11, 13, 13, num_public=4, rest_frac=0.5)
    S, O = _shapes(ctx, orc, inst)
    z = z_of(inst)
    cached_len = inst["num_shared"] + inst["num_precommitted"]
    z_cached = z.copy(); z_cached[cached_len:] = 0
    ca, cb, cc = S.multiply_vec(z_cached)                       # multiply_vec_precommitted (mod.rs:1112-1130)
    az, bz, cz = S.multiply_vec_incremental(z, ca, cb, cc)
    oa, ob, oc = O.multiply_vec(z)
    assert np.array_equal(az, oa) and np.array_equal(bz, ob) and np.array_equal(cz, oc)


def test_wrong_witness_length(ctx, orc):
    import spartan2_b200 as sp
    inst = random_r1cs(1, 4, 4)
    S, _ = _shapes(ctx, orc, inst)
    with pytest.raises(sp.SpartanError) as ei:
        S.multiply_vec(z_of(inst)[:-1])
    assert ei.value.kind == "InvalidWitnessLength"


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("lc,lv,seed", [(8, 7, 2), (13, 13, 4)])
def test_sharded_shape_rows_and_columns(ctx, orc, nranks, lc, lv, seed):
    # multi-GPU shards of the shape, every shard run on this one GPU: rank g owns rows / transposed columns
    # i = g (mod nranks); interleaving the shard outputs must reproduce the whole-shape result of the oracle
    import spartan2_b200 as sp
    inst = random_r1cs(seed, lc, lv, num_public=3, rest_frac=0.25)
    O = orc.Shape(*dims(inst), inst["A"], inst["B"], inst["C"])
    z = z_of(inst)
    oa, ob, oc = O.multiply_vec(z)
    rng = np.random.default_rng(seed)
    rx = orc.eq_evals(rand_fe(rng, lc)); r = rand_fe(rng, 1)
    oabc = O.abc(rx, r)
    cached_len = inst["num_shared"] + inst["num_precommitted"]
    z_cached = z.copy(); z_cached[cached_len:] = 0
    for g in range(nranks):
        S = sp.SplitR1CSShape(ctx, *dims(inst), inst["A"], inst["B"], inst["C"], rank=g, nranks=nranks)
        az, bz, cz = S.multiply_vec(z)
        assert np.array_equal(az, oa[g::nranks]) and np.array_equal(bz, ob[g::nranks]) and np.array_equal(cz, oc[g::nranks])
        ca, cb, cc = S.multiply_vec(z_cached)
        ia, ib, ic = S.multiply_vec_incremental(z, ca, cb, cc)
        assert np.array_equal(ia, az) and np.array_equal(ib, bz) and np.array_equal(ic, cz)
        got = S.bind_and_prepare_poly_ABC(rx, r)
        assert np.array_equal(got, oabc[g::nranks])
        S.free()
