"""GPU parity tests of the e2m1 variant of the dense path (tcgen05 kind::mxf4 on 4-bit operands, f32
accumulators holding exact integers; cv_dense_fp4.cu): integer sums bit-equal to the u8 kernel and to
the oracle, deviations / z bit-equal to the u8 kernel (same epilogue arithmetic on the same sums),
overflow columns for counts and multiplicities outside {0,1,2,3,4,6}, and every fallback to u8."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402
from test_gpu_parity import check_against_oracle, sets_to_csr  # noqa: E402

pytestmark = pytest.mark.gpu

KEYS = ("observed", "sampled", "dev", "z", "variability", "matches", "overlap")


@pytest.fixture
def fp4_engine(engine):
    engine.set_option("path", 2)
    yield engine
    for key, v in (("path", 0), ("dense_fp4", -1), ("dense_epilogue", 0), ("cell_chunk", 0)):
        engine.set_option(key, v)


def _problem(values, seed=4, n_sets=14):
    """20000 x 2000 synthetic counts whose nonzeros are redrawn from `values`."""
    from chromvar_b200 import synthetic
    d = synthetic.make_config("small")
    C = d["counts"].copy()
    rng = np.random.default_rng(seed)
    C.data = rng.choice(np.asarray(values, np.float64), C.nnz)
    P, N = C.shape
    bg = synthetic.random_background(P, 50, rng)
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    sets = [d["annoidx"][d["annoptr"][k]:d["annoptr"][k + 1]] for k in range(n_sets)]
    sets.insert(5, np.zeros(0, np.int32))
    ptr, idx = sets_to_csr(sets)
    return C, P, N, bg, e, ptr, idx, sets


def _both(engine, C, P, N, ptr, idx, bg, e, epilogue):
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    engine.set_option("dense_epilogue", epilogue)
    engine.set_option("dense_fp4", 0)
    u8 = engine.deviations(ptr, idx, bg, e, sums=True)
    st8 = engine.stats()
    engine.set_option("dense_fp4", 1)
    f4 = engine.deviations(ptr, idx, bg, e, sums=True)
    return u8, st8, f4, engine.stats()


@pytest.mark.parametrize("epilogue", [1, 2])
def test_fp4_equals_u8_bitwise_counts_in_alphabet(fp4_engine, epilogue):
    C, P, N, bg, e, ptr, idx, _ = _problem([1, 1, 1, 1, 2, 2, 3, 4, 6])
    u8, st8, f4, st4 = _both(fp4_engine, C, P, N, ptr, idx, bg, e, epilogue)
    assert st8["path"] == 1 and st8["operand_bits"] == 8
    assert st4["path"] == 1 and st4["operand_bits"] == 4 and st4["epilogue"] == (32 if epilogue == 1 else 64)
    for key in KEYS:
        assert np.array_equal(f4[key], u8[key], equal_nan=True), key


def test_fp4_overflow_columns_for_counts(fp4_engine):
    """counts 5, 7, 9, 11, 13, 25 (and a few 40 / 127) need 2..22 terms: overflow columns, several per peak"""
    C, P, N, bg, e, ptr, idx, sets = _problem([2, 2, 3, 4, 5, 6, 7, 9, 11, 13, 25] + [1] * 6000)
    C.data[np.random.default_rng(1).choice(C.nnz, 6, replace=False)] = [40, 127, 127, 40, 17, 29]
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    u8, st8, f4, st4 = _both(fp4_engine, C, P, N, ptr, idx, bg, e, 2)
    assert st4["operand_bits"] == 4 and st4["overflow_columns"] > 100
    for key in KEYS:
        assert np.array_equal(f4[key], u8[key], equal_nan=True), key
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    check_against_oracle(f4, ref)
    # chunked (the overflow columns are per cell chunk) == one chunk
    fp4_engine.set_option("cell_chunk", 512)
    g = fp4_engine.deviations(ptr, idx, bg, e, sums=True)
    assert fp4_engine.stats()["operand_bits"] == 4 and fp4_engine.stats()["n_cell_tiles"] == 4
    for key in KEYS:
        assert np.array_equal(g[key], u8[key], equal_nan=True), key


def test_fp4_overflow_columns_for_multiplicities(fp4_engine):
    """background columns that send 5..9 set peaks to one destination: operand entries outside the
    alphabet (T_q > 1), on peaks that also hold counts outside it (S_q > 1): (s, t) column grids"""
    C, P, N, bg, e, ptr, idx, sets = _problem([5, 8, 13] + [1, 1, 2, 3] * 1500, seed=9)
    bg = bg.copy()
    bg[:4500, 0] = np.arange(4500) // 9 + 1            # 9 sources per destination (500 destinations)
    bg[:2500, 7] = np.arange(2500) // 5 + 1 + 1000     # 5 sources per destination
    big = [np.arange(1, P + 1, dtype=np.int32)[::2], np.arange(1, 3001, dtype=np.int32)]   # dense sets: entries up to 5 / 9
    sets = list(sets) + big
    ptr, idx = sets_to_csr(sets)
    u8, st8, f4, st4 = _both(fp4_engine, C, P, N, ptr, idx, bg, e, 2)
    assert st4["operand_bits"] == 4 and st4["overflow_columns"] > 0
    for key in KEYS:
        assert np.array_equal(f4[key], u8[key], equal_nan=True), key
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    check_against_oracle(f4, ref)


def test_fp4_falls_back_when_the_overflow_reserve_is_exceeded(fp4_engine):
    """every nonzero is an 11 = 6 + 4 + 1: every peak needs two overflow columns, twice the reserve
    (one column per peak): detected on the device, the call redoes itself with u8 operands"""
    C, P, N, bg, e, ptr, idx, sets = _problem([11])
    fp4_engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    fp4_engine.set_option("dense_fp4", 1)
    got = fp4_engine.deviations(ptr, idx, bg, e, sums=True)
    assert fp4_engine.stats()["operand_bits"] == 8
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    check_against_oracle(got, ref)
    # ... and the streamed call does the same
    got2 = fp4_engine.deviations_csc(P, N, C.indptr, C.indices, C.data, ptr, idx, bg, e)
    assert fp4_engine.stats()["operand_bits"] == 8
    for key in ("dev", "z", "variability"):
        assert np.array_equal(got2[key], got[key], equal_nan=True), key
    # the veto does not outlive the call
    C2, *_ = _problem([1, 2])
    fp4_engine.set_counts_csc(P, N, C2.indptr, C2.indices, C2.data)
    fp4_engine.deviations(ptr, idx, bg, np.asarray(C2.sum(axis=1)).ravel() / C2.sum())
    assert fp4_engine.stats()["operand_bits"] == 4


def test_fp4_falls_back_when_sums_exceed_2_24(fp4_engine):
    """f32 accumulators are exact below 2^24 only: a cell whose total reaches that bound runs on u8 / s32"""
    C, P, N, bg, e, ptr, idx, sets = _problem([1, 2, 3])
    x = C.data.copy()
    lo, hi = C.indptr[3], C.indptr[4]
    x[lo:hi] = 250.0                                          # one heavy cell: total ~ 250 x 600
    C = sp.csc_matrix((x, C.indices, C.indptr), shape=C.shape)
    # the bound is (largest cell total) x (largest possible operand entry): a background column that
    # maps 200 peaks to one destination makes the second factor 200
    bg = bg.copy()
    bg[:200, 3] = 17
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    fp4_engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    fp4_engine.set_option("dense_fp4", 1)
    got = fp4_engine.deviations(ptr, idx, bg, e, sums=True)
    st = fp4_engine.stats()
    tot = np.asarray(C.sum(axis=0)).ravel().max()
    assert tot * 200 >= 2 ** 24 and st["operand_bits"] == 8
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    check_against_oracle(got, ref)


def test_fp4_streamed_call_equals_resident_call(fp4_engine):
    C, P, N, bg, e, ptr, idx, _ = _problem([2, 2, 3, 4, 5, 6, 7, 12] + [1] * 5000)
    fp4_engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    fp4_engine.set_option("dense_fp4", 0)
    base = fp4_engine.deviations(ptr, idx, bg, e)
    fp4_engine.set_option("dense_fp4", 1)
    for chunk, xdtype in ((0, np.float64), (512, np.int32), (1024, np.uint8)):
        fp4_engine.set_option("cell_chunk", chunk)
        got = fp4_engine.deviations_csc(P, N, C.indptr.astype(np.int32), C.indices, C.data.astype(xdtype), ptr, idx, bg, e)
        st = fp4_engine.stats()
        assert st["path"] == 1 and st["operand_bits"] == 4 and st["n_cell_tiles"] >= 2
        for key in ("dev", "z", "variability", "matches", "overlap"):
            assert np.array_equal(got[key], base[key], equal_nan=True), (chunk, key)


def test_fp4_duplicate_row_index_in_a_column_falls_back(fp4_engine):
    """Matrix sums duplicate (i, j) entries; nibbles cannot: detected, redone on u8"""
    C, P, N, bg, e, ptr, idx, sets = _problem([1, 2])
    indptr, indices, data = C.indptr.copy(), C.indices.copy(), C.data.copy()
    lo = indptr[10]
    assert indptr[11] - lo >= 2
    indices[lo + 1] = indices[lo]                             # cell 10 lists its first peak twice
    fp4_engine.set_counts_csc(P, N, indptr, indices, data)
    fp4_engine.set_option("dense_fp4", 1)
    Cd = sp.csc_matrix(sp.coo_matrix((data, (indices, np.repeat(np.arange(N), np.diff(indptr)))), shape=(P, N)))
    e2 = np.asarray(Cd.sum(axis=1)).ravel() / Cd.sum()
    e2[e2 == 0] = 1e-9
    try:
        got = fp4_engine.deviations(ptr, idx, bg, e2, sums=True)
    except Exception as ex:                                    # the zero-peak check may reject the edited matrix
        pytest.skip("edited matrix rejected: %s" % ex)
    assert fp4_engine.stats()["operand_bits"] == 8
    ref = orc.compute_deviations_core(Cd, [np.asarray(s) for s in sets], bg, e2, intermediate=True)
    assert np.array_equal(got["observed"], ref["observed"])


def test_fp4_falls_back_when_an_operand_entry_exceeds_a_nibble(fp4_engine):
    """40 peaks of one set drawn to the same background peak: the operand entry 40 does not fit the
    4-bit tile of the builder: flagged on the device, redone on u8"""
    C, P, N, bg, e, ptr, idx, sets = _problem([1, 2, 3])
    bg = bg.copy()
    bg[:40, 3] = 17
    sets = list(sets) + [np.arange(1, 61, dtype=np.int32)]
    ptr, idx = sets_to_csr(sets)
    fp4_engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    fp4_engine.set_option("dense_fp4", 1)
    got = fp4_engine.deviations(ptr, idx, bg, e, sums=True)
    assert fp4_engine.stats()["operand_bits"] == 8
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    check_against_oracle(got, ref)
    # 12 sources on one destination still fit (terms 6 + 6)
    bg[:40, 3] = np.arange(40) % 4 + 17
    got = fp4_engine.deviations(ptr, idx, bg, e, sums=True)
    assert fp4_engine.stats()["operand_bits"] == 4 and fp4_engine.stats()["overflow_columns"] > 0
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    check_against_oracle(got, ref)


def test_fp4_equals_u8_at_config2_size(engine):
    """BASELINE config 2 at full size (100 000 peaks x 10 000 cells x 870 annotations, 50 iterations):
    the e2m1 kernel (overflow columns for 0.4 % of the nonzeros, three cell chunks) and the u8 kernel
    give bit-identical deviations, z-scores and variabilities; integer sums of a slab of annotations are
    compared too (the full K x B x N int64 array would be 3.5 GB)."""
    from chromvar_b200 import synthetic
    d = synthetic.make_config("cfg2_scatac_motifs")
    C = d["counts"]
    P, N = C.shape
    bg = synthetic.random_background(P, 50, np.random.default_rng(1))
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    try:
        engine.set_option("path", 2)
        engine.set_option("dense_fp4", 0)
        u8 = engine.deviations(d["annoptr"], d["annoidx"], bg, e)
        assert engine.stats()["operand_bits"] == 8
        engine.set_option("dense_fp4", 1)
        f4 = engine.deviations(d["annoptr"], d["annoidx"], bg, e)
        st = engine.stats()
        assert st["operand_bits"] == 4 and st["overflow_columns"] > 0 and st["n_cell_tiles"] == 3
        for key in ("dev", "z", "variability", "matches", "overlap"):
            assert np.array_equal(f4[key], u8[key], equal_nan=True), key
        # integer sums of the first 24 annotations, both kernels
        ptr = d["annoptr"][:25]
        idx = d["annoidx"][:ptr[-1]]
        f4s = engine.deviations(ptr, idx, bg, e, sums=True)
        assert engine.stats()["operand_bits"] == 4
        engine.set_option("dense_fp4", 0)
        u8s = engine.deviations(ptr, idx, bg, e, sums=True)
        assert np.array_equal(f4s["observed"], u8s["observed"]) and np.array_equal(f4s["sampled"], u8s["sampled"])
        # observed sums against scipy for those annotations
        A = sp.csr_matrix((np.ones(len(idx)), idx - 1, ptr), shape=(24, P))
        assert np.array_equal(f4s["observed"], np.asarray((A @ C).todense()).astype(np.int64))
    finally:
        for key, v in (("path", 0), ("dense_fp4", -1)):
            engine.set_option(key, v)


@pytest.mark.parametrize("B", [2, 7, 14, 15])
@pytest.mark.parametrize("epilogue", [1, 2])
def test_fp4_few_iterations_many_annotations_per_sub_tile(fp4_engine, B, epilogue):
    """With B < 15 a 256-row sub-tile could hold more than DN_MAX_G = 16 annotations; the epilogue's
    per-annotation scratch is sized for 16 (an overrun corrupted the mbarriers behind it: illegal address
    with K >= 17, found by tests/test_gpu_fuzz.py).  The plan now caps the annotations per sub-tile."""
    rng = np.random.default_rng(B)
    P, N, K = 700, 300, 70
    X = (rng.random((P, N)) < 0.2) * rng.integers(1, 4, (P, N)).astype(np.float64)
    X[X.sum(axis=1) == 0, 0] = 1
    C = sp.csc_matrix(X)
    sets = [rng.choice(np.arange(1, P + 1), int(rng.integers(20, 300)), replace=False) for _ in range(K)]
    ptr, idx = sets_to_csr(sets)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, B)).astype(np.int32))
    e = orc.compute_expectations_core(C)
    u8, st8, f4, st4 = _both(fp4_engine, C, P, N, ptr, idx, bg, e, epilogue)
    assert st4["path"] == 1 and st4["operand_bits"] == 4 and st4["epilogue"] == (32 if epilogue == 1 else 64)
    for key in ("observed", "sampled", "matches", "overlap"):
        assert np.array_equal(f4[key], u8[key], equal_nan=True), key
    check_against_oracle(f4, orc.compute_deviations_core(C, sets, bg, e, intermediate=True))
