"""The few-cells contraction (cv_dense_gather.cu): bulk-like inputs with at most 256 samples.  Per
iteration the counts rows are gathered by the background indices into an MN-major tcgen05 operand
(cp.async into a hand-swizzled shared-memory tile), both u8 digit planes accumulate side by side in
tensor memory, nothing is stacked.  Integer sums bit-equal to the oracle and to the stacked-operand
path.  The expectation table rides on the same GEMM as fixed-point digits in the spare cell columns (the
correctly rounded sum of the exact values), so deviations / z agree with the stacked-operand path -- whose
table is an fp64 sum in list order -- to ~1e-13 instead of bit for bit; with no spare columns
(N > 244) the table comes from the same kernel as there and everything is bit-equal."""
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


def _bulk(P, N, rng, top=3000):
    a = rng.lognormal(0.0, 1.0, (P, 1))
    X = rng.poisson(a * rng.uniform(5, 60, (1, N))).astype(np.float64)
    X[rng.integers(0, P, 5), rng.integers(0, N, 5)] = top
    X[X.sum(axis=1) == 0, 0] = 1
    return X


@pytest.fixture
def geng(engine):
    yield engine
    for key, v in (("path", 0), ("small_n_gather", -1), ("dense_fp4", -1)):
        engine.set_option(key, v)


@pytest.mark.parametrize("shape", [(3000, 200, 40, 50), (5001, 37, 7, 50), (1000, 256, 300, 3), (2176, 2, 5, 20)])
def test_gather_path_equals_oracle_and_stacked_path(geng, shape):
    P, N, K, B = shape
    rng = np.random.default_rng(P + N)
    X = _bulk(P, N, rng)
    C = sp.csc_matrix(X)
    sets = [rng.choice(np.arange(1, P + 1), int(rng.integers(1, max(2, P // 3))), replace=False) for _ in range(K)]
    if K > 3:
        sets[2] = np.zeros(0, np.int64)                      # an empty annotation: all-NA row
    ptr, idx = sets_to_csr(sets)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, B)).astype(np.int32))
    e = orc.compute_expectations_core(C)
    geng.set_option("path", 2)
    geng.set_option("small_n_gather", 0)
    geng.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    base = geng.deviations(ptr, idx, bg, e, sums=True)
    assert geng.stats()["few_cells"] == 0 and geng.stats()["planes"] == 2
    geng.set_option("small_n_gather", 1)
    got = geng.deviations(ptr, idx, bg, e, sums=True)
    st = geng.stats()
    assert st["few_cells"] == 1 and st["path"] == 1
    exact_table = N + 12 > 256                            # no spare columns: same table kernel as the stacked path
    for key in KEYS:
        if exact_table or key in ("observed", "sampled", "matches", "overlap"):
            assert np.array_equal(got[key], base[key], equal_nan=True), key
        else:
            assert np.array_equal(np.isnan(got[key]), np.isnan(base[key])), key
            ok = np.isfinite(base[key])
            assert np.array_equal(got[key][~ok & ~np.isnan(base[key])], base[key][~ok & ~np.isnan(base[key])]), key
            assert np.max(np.abs(got[key][ok] - base[key][ok]) / np.maximum(np.abs(base[key][ok]), 1e-3)) < 1e-9, key
    if K <= 40:
        live = [k for k in range(K) if len(sets[k])]
        ref = orc.compute_deviations_core(C, [sets[k] for k in live], bg, e, intermediate=True)
        sub = {key: (got[key][live] if got[key].shape[0] == K else got[key]) for key in KEYS}
        check_against_oracle(sub, ref)
    # dense (base matrix) input and the queued call take the same path
    geng.set_counts_dense(X)
    again = geng.deviations(ptr, idx, bg, e, sums=False)
    assert geng.stats()["few_cells"] == 1
    for key in ("dev", "z"):
        assert np.array_equal(again[key], got[key], equal_nan=True), key
    geng.deviations_upload(ptr, idx, bg, e)
    geng.deviations_compute(rebuild_counts_layout=True)
    geng.deviations_compute_async(rebuild_counts_layout=True)
    geng.deviations_compute_async(rebuild_counts_layout=True)
    geng.deviations_wait()
    q = geng.deviations_download()
    for key in ("dev", "z", "variability"):
        assert np.array_equal(q[key], got[key], equal_nan=True), key


def test_gather_path_declines_what_it_cannot_represent(geng):
    rng = np.random.default_rng(9)
    P, N, B = 2000, 64, 10
    X = _bulk(P, N, rng)
    C = sp.csc_matrix(X)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, B)).astype(np.int32))
    e = orc.compute_expectations_core(C)
    geng.set_option("path", 2)
    geng.set_option("small_n_gather", 1)
    geng.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    # a list that repeats a peak: the general paths take over, the multiplicity counts
    sets = [np.array([5, 5, 9, 200]), rng.choice(np.arange(1, P + 1), 300, replace=False)]
    ptr, idx = sets_to_csr(sets)
    got = geng.deviations(ptr, idx, bg, e, sums=True)
    assert geng.stats()["few_cells"] == 0
    ref = orc.compute_deviations_core(C, sets, bg, e, intermediate=True)
    check_against_oracle(got, ref)
    # counts of 65536 and more do not fit two u8 planes
    X2 = X.copy()
    X2[3, 3] = 70000
    C2 = sp.csc_matrix(X2)
    geng.set_counts_csc(P, N, C2.indptr, C2.indices, C2.data)
    sets = [rng.choice(np.arange(1, P + 1), 300, replace=False)]
    ptr, idx = sets_to_csr(sets)
    e2 = orc.compute_expectations_core(C2)
    got = geng.deviations(ptr, idx, bg, e2, sums=True)
    assert geng.stats()["few_cells"] == 0
    check_against_oracle(got, orc.compute_deviations_core(C2, sets, bg, e2, intermediate=True))
    # an expectation outside [0, 1) cannot be written in the digit columns: the general paths take over
    sets = [rng.choice(np.arange(1, P + 1), 300, replace=False)]
    ptr, idx = sets_to_csr(sets)
    e_big = e * 1e6
    geng.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    got = geng.deviations(ptr, idx, bg, e_big, sums=True)
    assert geng.stats()["few_cells"] == 0
    check_against_oracle(got, orc.compute_deviations_core(C, sets, bg, e_big, intermediate=True))
    # an out-of-range background index is still an error
    bad = bg.copy()
    bad[7, 2] = P + 1
    geng.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    from chromvar_b200 import _lib
    with pytest.raises(_lib.ChromVARError):
        geng.deviations(ptr, idx, bad, e)
