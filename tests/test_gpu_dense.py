"""GPU parity tests of the dense tensor-core path (tcgen05 u8 x u8 -> s32 GEMM with the stacked
annotation x iteration operand and the TMEM-resident fused epilogue), forced with
cv_set_option("path", 2): same bars as the row-gather path -- integer sums bit-exact against the
oracle (and against the row-gather kernel), deviations / z within 1e-5 relative, identical NA masks."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402
from test_gpu_parity import assert_close, check_against_oracle, sets_to_csr, _mini, _example  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["cta_pair", "single_cta"])
def dense_engine(engine, request):
    """Dense path forced, once per GEMM kernel: tcgen05 cta_group::2 (256-cell tiles on two SMs,
    the default) and the single-CTA kernel (128-cell tiles)."""
    engine.set_option("path", 2)
    engine.set_option("dense_cta_pair", 1 if request.param == "cta_pair" else 0)
    yield engine
    engine.set_option("path", 0)
    engine.set_option("dense_cta_pair", -1)
    engine.set_option("cell_chunk", 0)
    engine.set_option("dense_epilogue", 0)


def _run(engine, C, sets, bg, e, sums=True):
    Cc = sp.csc_matrix(C)
    engine.set_counts_csc(Cc.shape[0], Cc.shape[1], Cc.indptr, Cc.indices, Cc.data)
    ptr, idx = sets_to_csr(sets)
    return engine.deviations(ptr, idx, bg, e, index_base=1, sums=sums)


def test_dense_known_answer(dense_engine, golden_deviations_test):
    g = golden_deviations_test
    got = _run(dense_engine, g["counts"], [g["set0"], g["set1"], g["set2"]], g["bg"],
               g["counts"].sum(axis=1) / g["counts"].sum())
    assert dense_engine.stats()["path"] == 1 and dense_engine.stats()["epilogue"] == 64   # tiny: fp64 epilogue
    assert np.max(np.abs(got["dev"] - g["deviations"])) < 1e-12
    assert np.max(np.abs(got["z"] - g["z"])) < 1e-12
    dense_engine.set_option("dense_epilogue", 1)          # float-pair epilogue on the reference's known answer
    got = _run(dense_engine, g["counts"], [g["set0"], g["set1"], g["set2"]], g["bg"],
               g["counts"].sum(axis=1) / g["counts"].sum())
    assert dense_engine.stats()["epilogue"] == 32
    assert np.max(np.abs(got["dev"] - g["deviations"]) / np.maximum(np.abs(g["deviations"]), 1e-3)) < 1e-5
    assert np.max(np.abs(got["z"] - g["z"]) / np.maximum(np.abs(g["z"]), 1e-3)) < 1e-5


def test_dense_mini_against_oracle(dense_engine, golden_mini):
    C, A = _mini(golden_mini)
    sets = orc.convert_to_ix_list(A)
    e = orc.compute_expectations_core(C)
    bg = orc.get_background_peaks_core(C, golden_mini["bias"], rng=np.random.default_rng(3))
    got = _run(dense_engine, C, sets, bg, e)
    assert dense_engine.stats()["path"] == 1
    ref = orc.compute_deviations_core(C, sets, bg, e, intermediate=True)
    check_against_oracle(got, ref)
    assert_close(got["variability"], orc.row_sds(ref["z"], True), "variability")


def test_dense_example_counts_edge_cases(dense_engine, golden_example_counts):
    """29k peaks x 50 cells: single-peak sets, duplicates, the whole peak set, NA mask, and more
    annotations than one 5-annotation accumulator group (tile raster + partial last group)."""
    from chromvar_b200._lib import is_na
    C = _example(golden_example_counts)
    C = C[np.nonzero(np.asarray(C.sum(axis=1)).ravel() > 0)[0], :].tocsc()
    P = C.shape[0]
    rng = np.random.default_rng(42)
    sizes = [1, 1, 2, 3, 7, 40, 300, 2000, 9000, P, 55, 1200, 17]
    sets = [np.sort(rng.choice(P, s, replace=False)) + 1 for s in sizes]
    sets.append(np.array([5, 5, 5, 9, 12]))
    sets.append(rng.integers(1, P + 1, 500))
    bg = orc.get_background_peaks_core(C, rng.uniform(0.3, 0.8, P), niterations=50, rng=np.random.default_rng(8))
    e = orc.compute_expectations_core(C)
    got = _run(dense_engine, C, sets, bg, e)
    assert dense_engine.stats()["path"] == 1
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    check_against_oracle(got, ref)
    fps = np.asarray(C.sum(axis=0)).ravel()
    fail = np.array([e[np.asarray(s) - 1].sum() * fps < 1 for s in sets])
    assert fail.any() and np.array_equal(is_na(got["z"]), fail)


def test_dense_equals_row_gather_at_size(engine):
    """20000 x 2000 synthetic: both kernels give identical integer sums; dev / z agree to 1e-9;
    an empty annotation in the middle; more cells than one 128-cell block."""
    from chromvar_b200 import synthetic
    d = synthetic.make_config("small")
    C = d["counts"]
    P, N = C.shape
    rng = np.random.default_rng(2)
    bg = synthetic.random_background(P, 50, rng)
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    sets = [d["annoidx"][d["annoptr"][k]:d["annoptr"][k + 1]] for k in range(12)]
    sets.insert(6, np.zeros(0, np.int32))                     # an empty annotation in the middle
    ptr, idx = sets_to_csr(sets)
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    engine.set_option("path", 1)
    a = engine.deviations(ptr, idx, bg, e, sums=True)
    assert engine.stats()["path"] == 0
    engine.set_option("path", 2)
    engine.set_option("dense_epilogue", 2)                    # fp64 epilogue: agrees with the row-gather kernel to rounding
    try:
        b = engine.deviations(ptr, idx, bg, e, sums=True)
        assert engine.stats()["path"] == 1 and engine.stats()["epilogue"] == 64
        engine.set_option("dense_epilogue", 1)                # float-pair epilogue: same sums, z / dev to ~1e-7
        f = engine.deviations(ptr, idx, bg, e, sums=True)
        assert engine.stats()["epilogue"] == 32
    finally:
        engine.set_option("path", 0)
        engine.set_option("dense_epilogue", 0)
    assert np.array_equal(f["observed"], b["observed"]) and np.array_equal(f["sampled"], b["sampled"])
    assert np.array_equal(np.isnan(f["z"]), np.isnan(b["z"]))
    okf = ~np.isnan(b["z"])
    assert np.max(np.abs(f["z"][okf] - b["z"][okf]) / np.maximum(np.abs(b["z"][okf]), 1e-3)) < 5e-6
    assert np.max(np.abs(f["dev"][okf] - b["dev"][okf]) / np.maximum(np.abs(b["dev"][okf]), 1e-3)) < 5e-6
    assert np.allclose(f["variability"], b["variability"], rtol=5e-6, equal_nan=True)
    assert np.array_equal(a["observed"], b["observed"])
    assert np.array_equal(a["sampled"], b["sampled"])
    assert np.array_equal(np.isnan(a["z"]), np.isnan(b["z"]))
    ok = ~np.isnan(a["z"])
    assert np.max(np.abs(a["z"][ok] - b["z"][ok]) / np.maximum(np.abs(a["z"][ok]), 1e-3)) < 1e-9
    assert np.max(np.abs(a["dev"][ok] - b["dev"][ok]) / np.maximum(np.abs(a["dev"][ok]), 1e-3)) < 1e-9
    assert np.allclose(a["variability"], b["variability"], rtol=1e-9, equal_nan=True)
    assert np.array_equal(a["matches"], b["matches"]) and np.allclose(a["overlap"], b["overlap"], equal_nan=True)


@pytest.mark.parametrize("B", [2, 15, 31, 50, 64, 127, 255])
def test_dense_niterations(dense_engine, golden_mini, B):
    C, A = _mini(golden_mini)
    sets = orc.convert_to_ix_list(A)
    e = orc.compute_expectations_core(C)
    bg = np.random.default_rng(B).integers(1, 1001, (1000, B)).astype(np.int32)
    got = _run(dense_engine, C, sets, bg, e)
    ref = orc.compute_deviations_core(C, sets, bg, e, intermediate=True)
    check_against_oracle(got, ref)


def test_dense_digit_planes_for_large_counts(dense_engine):
    """Counts above 255 run as u8 digit planes (one GEMM per plane, int64 sums combined, same fp64
    finish): bulk-like dense counts up to ~70000 (3 planes), 300 samples (3 cell blocks, the last
    one partial), bit-exact sums against the oracle."""
    rng = np.random.default_rng(0)
    P, N = 900, 300
    D = rng.negative_binomial(2, 0.001, (P, N)).astype(np.float64)
    D[rng.random(D.shape) < 0.05] = 0
    D[0, 0] = 70000.0
    D[D.sum(axis=1) == 0, 0] = 1
    bg = rng.integers(1, P + 1, (P, 50)).astype(np.int32)
    e = D.sum(axis=1) / D.sum()
    sets = [np.sort(rng.choice(P, s, replace=False)) + 1 for s in (1, 30, 200, 900, 45, 77, 3)]
    ptr, idx = sets_to_csr(sets)
    dense_engine.set_counts_dense(D)
    got = dense_engine.deviations(ptr, idx, bg, e, sums=True)
    st = dense_engine.stats()
    assert st["path"] == 1 and st["planes"] == 3
    ref = orc.compute_deviations_core(sp.csc_matrix(D), sets, bg, e, intermediate=True)
    check_against_oracle(got, ref)
    assert_close(got["variability"], orc.row_sds(ref["z"], True), "variability")
    # two planes, and the row-gather kernel agrees bit for bit on the sums
    D2 = np.minimum(D, 60000.0)
    dense_engine.set_counts_dense(D2)
    e2 = D2.sum(axis=1) / D2.sum()
    a = dense_engine.deviations(ptr, idx, bg, e2, sums=True)
    assert dense_engine.stats()["planes"] == 2
    dense_engine.set_option("path", 1)
    b = dense_engine.deviations(ptr, idx, bg, e2, sums=True)
    dense_engine.set_option("path", 2)
    assert np.array_equal(a["observed"], b["observed"]) and np.array_equal(a["sampled"], b["sampled"])
    assert_close(a["z"], b["z"], "z planes vs row-gather")


def test_dense_refuses_what_it_cannot_represent(dense_engine):
    """A peak listed 300 times in one annotation does not fit the u8 operand: forced dense refuses,
    auto mode falls back to the exact row-gather kernel."""
    from chromvar_b200 import _lib
    rng = np.random.default_rng(0)
    D = rng.integers(0, 200, (300, 20)).astype(np.float64)
    D[D.sum(axis=1) == 0, 0] = 1
    bg = rng.integers(1, 301, (300, 10)).astype(np.int32)
    dense_engine.set_counts_dense(D)
    sets = [np.concatenate([np.full(300, 7), np.arange(1, 100)]), np.arange(5, 50)]
    ptr, idx = sets_to_csr(sets)
    with pytest.raises(_lib.ChromVARError) as ei:
        dense_engine.deviations(ptr, idx, bg, D.sum(axis=1) / D.sum())
    assert ei.value.code == _lib.CV_ERR_UNSUPPORTED
    dense_engine.set_option("path", 0)
    got = dense_engine.deviations(ptr, idx, bg, D.sum(axis=1) / D.sum(), sums=True)
    ref = orc.compute_deviations_core(sp.csc_matrix(D), sets, bg, D.sum(axis=1) / D.sum(), intermediate=True)
    check_against_oracle(got, ref)
    assert dense_engine.stats()["path"] == 0


def _small_problem():
    from chromvar_b200 import synthetic
    d = synthetic.make_config("small")
    C = d["counts"]
    P, N = C.shape
    bg = synthetic.random_background(P, 50, np.random.default_rng(2))
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    sets = [d["annoidx"][d["annoptr"][k]:d["annoptr"][k + 1]] for k in range(12)]
    sets.insert(6, np.zeros(0, np.int32))
    ptr, idx = sets_to_csr(sets)
    return C, P, N, bg, e, ptr, idx


def test_dense_cell_chunks_and_kernels_agree_bitwise(engine):
    """20000 x 2000: one chunk vs chunks of 256 / 768 cells (the last one partial), CTA-pair vs
    single-CTA kernel: every output is bit-identical (same integer sums, same fp64 epilogue)."""
    C, P, N, bg, e, ptr, idx = _small_problem()
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    engine.set_option("path", 2)
    try:
        base = engine.deviations(ptr, idx, bg, e, sums=True)
        assert engine.stats()["cell_tile"] == 256          # host copies beside the GEMM: CTA pair by default
        for pair, chunk in ((1, 256), (1, 768), (0, 0), (0, 640)):
            engine.set_option("dense_cta_pair", pair)
            engine.set_option("cell_chunk", chunk)
            got = engine.deviations(ptr, idx, bg, e, sums=True)
            st = engine.stats()
            assert st["cell_tile"] == (256 if pair else 128)
            if chunk:
                assert st["n_cell_tiles"] == -(-N // (-(-chunk // 256) * 256))
            for key in ("observed", "sampled", "dev", "z", "variability", "matches", "overlap"):
                assert np.array_equal(got[key], base[key], equal_nan=True), (pair, chunk, key)
        # operand rows built by the per-annotation scatter kernel instead of the gather kernel:
        # identical operand (identical integer sums); the expected-fraction table is summed in a
        # different order, so dev / z agree to rounding
        engine.set_option("dense_rows_scatter", 1)
        got = engine.deviations(ptr, idx, bg, e, sums=True)
        for key in ("observed", "sampled", "matches", "overlap"):
            assert np.array_equal(got[key], base[key], equal_nan=True), key
        ok = ~np.isnan(base["z"])
        assert np.array_equal(np.isnan(got["z"]), ~ok)
        assert np.max(np.abs(got["z"][ok] - base["z"][ok]) / np.maximum(np.abs(base["z"][ok]), 1e-3)) < 1e-9
    finally:
        engine.set_option("path", 0)
        engine.set_option("dense_cta_pair", -1)
        engine.set_option("cell_chunk", 0)
        engine.set_option("dense_rows_scatter", 0)


def test_streamed_call_equals_resident_call(engine):
    """cv_deviations_csc (counts arriving in cell chunks while earlier chunks are in the GEMM,
    results leaving chunk by chunk) against cv_set_counts_csc + cv_deviations: bit-identical, and
    the counts are resident afterwards."""
    C, P, N, bg, e, ptr, idx = _small_problem()
    engine.set_option("path", 2)
    try:
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
        base = engine.deviations(ptr, idx, bg, e)
        for chunk, xdtype in ((0, np.float64), (512, np.float64), (0, np.int32), (1024, np.uint8)):
            engine.set_option("cell_chunk", chunk)
            got = engine.deviations_csc(P, N, C.indptr.astype(np.int32), C.indices, C.data.astype(xdtype),
                                        ptr, idx, bg, e)
            st = engine.stats()
            assert st["path"] == 1 and st["n_cell_tiles"] >= 2          # really chunked
            for key in ("dev", "z", "variability", "matches", "overlap"):
                assert np.array_equal(got[key], base[key], equal_nan=True), (chunk, key)
            fpp, fps, tot = engine.fragments()
            assert np.array_equal(fpp, np.asarray(C.sum(axis=1)).ravel())
            assert np.array_equal(fps, np.asarray(C.sum(axis=0)).ravel()) and tot == C.sum()
            assert np.array_equal(engine.expectations(), e)
        # default expectation (NULL): the plain sequence inside the same entry point
        got = engine.deviations_csc(P, N, C.indptr, C.indices, C.data, ptr, idx, bg, None)
        for key in ("dev", "z"):
            assert np.array_equal(got[key], base[key], equal_nan=True), key
    finally:
        engine.set_option("path", 0)
        engine.set_option("cell_chunk", 0)


def test_streamed_call_validates_and_falls_back(engine):
    from chromvar_b200 import _lib
    C, P, N, bg, e, ptr, idx = _small_problem()
    engine.set_option("path", 2)
    try:
        x = C.data.astype(np.float64).copy()
        x[len(x) // 2] = 2.5                                   # a fractional count in the middle chunk
        with pytest.raises(_lib.ChromVARError) as ei:
            engine.deviations_csc(P, N, C.indptr, C.indices, x, ptr, idx, bg, e)
        # a restriction of this engine (exact integer sums), reported as such -- NOT as counts_check,
        # which only requires counts >= 0 (R/utils.R:3-12)
        assert ei.value.code == _lib.CV_ERR_UNSUPPORTED and "fractional" in str(ei.value)
        assert "counts_check" not in str(ei.value)
        x[len(x) // 2] = -1.0                                  # counts_check proper
        with pytest.raises(_lib.ChromVARError) as ei:
            engine.deviations_csc(P, N, C.indptr, C.indices, x, ptr, idx, bg, e)
        assert ei.value.code == _lib.CV_ERR_INVALID and "counts_check" in str(ei.value)
        bad_bg = bg.copy()
        bad_bg[17, 3] = P + 1
        with pytest.raises(_lib.ChromVARError) as ei:
            engine.deviations_csc(P, N, C.indptr, C.indices, C.data, ptr, idx, bad_bg, e)
        assert ei.value.code == _lib.CV_ERR_INVALID
        # a count above 255 shows up only once K1 has seen it: the call redoes itself with digit planes
        x = C.data.astype(np.float64).copy()
        x[-5] = 1000.0
        C2 = sp.csc_matrix((x, C.indices, C.indptr), shape=C.shape)
        e2 = np.asarray(C2.sum(axis=1)).ravel() / C2.sum()
        got = engine.deviations_csc(P, N, C.indptr, C.indices, x, ptr, idx, bg, e2)
        assert engine.stats()["planes"] == 2
        engine.set_counts_csc(P, N, C2.indptr, C2.indices, C2.data)
        ref = engine.deviations(ptr, idx, bg, e2)
        for key in ("dev", "z", "variability"):
            assert np.array_equal(got[key], ref[key], equal_nan=True), key
    finally:
        engine.set_option("path", 0)


@pytest.mark.parametrize("maxcount", [9, 3000])
def test_dense_many_peaks(dense_engine, maxcount):
    """260k peaks: operand rows longer than one shared-memory chunk (three chunks in the row
    builder, two in the cell builder), one and two u8 planes, against the oracle."""
    rng = np.random.default_rng(11)
    P, N, B = 260_000, 70, 50
    nnz_col = 9000
    rows = np.concatenate([np.sort(rng.choice(P, nnz_col, replace=False)) for _ in range(N)])
    vals = rng.integers(1, maxcount + 1, rows.size).astype(np.float64)
    indptr = np.arange(N + 1, dtype=np.int64) * nnz_col
    C = sp.csc_matrix((vals, rows, indptr), shape=(P, N))
    empty = np.nonzero(np.asarray(C.sum(axis=1)).ravel() == 0)[0]
    C = (C + sp.csc_matrix((np.ones(empty.size), (empty, rng.integers(0, N, empty.size))), shape=(P, N))).tocsc()
    bg = rng.integers(1, P + 1, (P, B)).astype(np.int32)
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    sets = [np.sort(rng.choice(P, s, replace=False)) + 1 for s in (1, 700, 40_000, 130_000, 2_500, 260_000, 9)]
    got = _run(dense_engine, C, sets, bg, e)
    st = dense_engine.stats()
    assert st["path"] == 1 and st["planes"] == (1 if maxcount < 256 else 2)
    ref = orc.compute_deviations_core(C, sets, bg, e, intermediate=True)
    check_against_oracle(got, ref)


def test_streamed_call_host_narrowing(engine):
    """Wide value types are narrowed to bytes on host threads before the upload ("host_narrow"):
    same results bit for bit, fewer bytes on the wire; a chunk holding a count above 255 travels in
    the caller's type while the other chunks still travel as bytes."""
    C, P, N, bg, e, ptr, idx = _small_problem()
    engine.set_option("path", 2)
    try:
        engine.set_option("cell_chunk", 512)
        x64 = C.data.astype(np.float64)
        engine.set_option("host_narrow", 0)
        base = engine.deviations_csc(P, N, C.indptr, C.indices, x64, ptr, idx, bg, e)
        wide = engine.stats()["h2d_bytes"]
        for threads in (1, 3):
            engine.set_option("host_narrow", threads)
            got = engine.deviations_csc(P, N, C.indptr, C.indices, x64, ptr, idx, bg, e)
            st = engine.stats()
            assert st["n_cell_tiles"] >= 2
            assert wide - st["h2d_bytes"] == 7 * C.nnz                   # 1 byte instead of 8 per nonzero
            for key in ("dev", "z", "variability", "matches", "overlap"):
                assert np.array_equal(got[key], base[key], equal_nan=True), key
            fpp, fps, tot = engine.fragments()
            assert np.array_equal(fpp, np.asarray(C.sum(axis=1)).ravel()) and tot == C.sum()
        # one large count in the last cell: only the chunks whose slices hold it keep 8 bytes per value
        x = x64.copy()
        x[-3] = 700.0
        C2 = sp.csc_matrix((x, C.indices, C.indptr), shape=C.shape)
        e2 = np.asarray(C2.sum(axis=1)).ravel() / C2.sum()
        engine.set_option("host_narrow", 0)
        ref = engine.deviations_csc(P, N, C.indptr, C.indices, x, ptr, idx, bg, e2)
        engine.set_option("host_narrow", 1)
        got = engine.deviations_csc(P, N, C.indptr, C.indices, x, ptr, idx, bg, e2)
        assert engine.stats()["planes"] == 2
        for key in ("dev", "z", "variability"):
            assert np.array_equal(got[key], ref[key], equal_nan=True), key
    finally:
        engine.set_option("path", 0)
        engine.set_option("cell_chunk", 0)
        engine.set_option("host_narrow", -1)


def test_streamed_call_stages_pageable_memory(engine):
    """Default ("host_narrow" = -1): pageable caller arrays (numpy's, R's) are staged by host threads --
    row indices copied, values narrowed, results landed in page-locked memory -- pinned ones travel
    directly.  Same results bit for bit either way."""
    import torch
    C, P, N, bg, e, ptr, idx = _small_problem()
    assert C.nnz >= 1 << 20
    engine.set_option("path", 2)
    try:
        x64 = np.ascontiguousarray(C.data.astype(np.float64))
        ri = np.ascontiguousarray(C.indices.astype(np.int32))
        engine.set_option("host_narrow", 0)
        base = engine.deviations_csc(P, N, C.indptr, ri, x64, ptr, idx, bg, e)
        wide = engine.stats()["h2d_bytes"]
        engine.set_option("host_narrow", -1)
        got = engine.deviations_csc(P, N, C.indptr, ri, x64, ptr, idx, bg, e)      # pageable: staged
        assert wide - engine.stats()["h2d_bytes"] == 7 * C.nnz
        for key in ("dev", "z", "variability", "matches", "overlap"):
            assert np.array_equal(got[key], base[key], equal_nan=True), key
        # pinned inputs: the default leaves them alone (all 8 bytes per value cross PCIe)
        xp = torch.from_numpy(x64).pin_memory().numpy()
        rp = torch.from_numpy(ri).pin_memory().numpy()
        got = engine.deviations_csc(P, N, C.indptr, rp, xp, ptr, idx, bg, e)
        assert engine.stats()["h2d_bytes"] == wide
        for key in ("dev", "z"):
            assert np.array_equal(got[key], base[key], equal_nan=True), key
    finally:
        engine.set_option("path", 0)
        engine.set_option("host_narrow", -1)


@pytest.mark.parametrize("shuffled", [False, True])
def test_dense_long_columns_split_builder(engine, shuffled):
    """Bulk-like input: every peak of every sample holds a count, so the cells operand is built by
    (cell, peak range) items (k_dense_cells_split) -- by binary search into sorted columns, by a
    filtered scan when a column lists its rows out of order (K1 notices)."""
    rng = np.random.default_rng(21)
    P, N, B = 70_000, 40, 50
    X = rng.poisson(rng.lognormal(1.0, 1.0, (P, 1)) * rng.uniform(2, 6, (1, N))).astype(np.float64)
    X[X.sum(axis=1) == 0, 0] = 1
    C = sp.csc_matrix(X)
    assert C.nnz / N > 16384 and X.max() > 255
    indptr, indices, data = C.indptr.copy(), C.indices.copy(), C.data.copy()
    if shuffled:
        for c in range(0, N, 3):
            a, b = indptr[c], indptr[c + 1]
            perm = rng.permutation(b - a)
            indices[a:b], data[a:b] = indices[a:b][perm], data[a:b][perm]
    sets = [rng.choice(np.arange(1, P + 1), n, replace=False) for n in (5000, 700, 90, 20000)]
    ptr, idx = sets_to_csr(sets)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, B)).astype(np.int32))
    e = orc.compute_expectations_core(C)
    engine.set_option("path", 2)
    try:
        engine.set_counts_csc(P, N, indptr, indices, data)
        got = engine.deviations(ptr, idx, bg, e, sums=True)
        st = engine.stats()
        assert st["path"] == 1 and st["planes"] == 2
        ref = orc.compute_deviations_core(C, sets, bg, e, intermediate=True)
        check_against_oracle(got, ref)
    finally:
        engine.set_option("path", 0)
