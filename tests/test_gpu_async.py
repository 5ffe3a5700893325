"""cv_deviations_compute_async / cv_deviations_wait: the resident call with the host out of the loop
(a replay of the last fully checked call) gives bit-identical results, several calls may be queued,
and anything that changes inputs or options, or that took a fallback, goes back to the checked call."""
import numpy as np
import pytest
import scipy.sparse as sp

from chromvar_b200 import synthetic

pytestmark = pytest.mark.gpu


def _problem():
    d = synthetic.make_config("small")
    C = d["counts"]
    P, N = C.shape
    bg = synthetic.random_background(P, 50, np.random.default_rng(2))
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    K = 12
    ptr = d["annoptr"][:K + 1].astype(np.int64)
    idx = d["annoidx"][:ptr[-1]].astype(np.int32)
    return C, P, N, bg, e, ptr, idx


def _same(a, b):
    for key in ("dev", "z", "variability", "matches", "overlap"):
        assert np.array_equal(a[key], b[key], equal_nan=True), key


def test_async_replay_is_bit_identical(engine):
    C, P, N, bg, e, ptr, idx = _problem()
    engine.set_option("path", 2)
    try:
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
        engine.deviations_upload(ptr, idx, bg, e, index_base=1)
        # no verdict yet: the async entry point is the checked call
        engine.deviations_compute_async(True)
        base = engine.deviations_download()
        assert engine.stats()["path"] == 1
        launches0 = engine.stats()["kernel_launches"]
        # five replays queued behind each other, settled by the wait
        for _ in range(5):
            engine.deviations_compute_async(True)
        engine.deviations_wait()
        st = engine.stats()
        assert st["kernel_launches"] - launches0 >= 5 * 20            # every kernel of the path ran again, five times
        assert st["ms_contract"] > 0 and st["path"] == 1
        _same(engine.deviations_download(), base)
        # settled implicitly by the download
        engine.deviations_compute_async(True)
        _same(engine.deviations_download(), base)
        # the checked call after replays, and replays after it
        engine.deviations_compute(True)
        _same(engine.deviations_download(), base)
        engine.deviations_compute_async(False)
        engine.deviations_compute_async(True)
        _same(engine.deviations_download(), base)
        # new annotations invalidate the verdict: results follow the new inputs
        ptr2 = ptr[:7].copy()
        idx2 = idx[:ptr2[-1]]
        engine.deviations_upload(ptr2, idx2, bg, e, index_base=1)
        engine.deviations_compute_async(True)
        got = engine.deviations_download()
        for key in ("dev", "z"):
            assert np.array_equal(got[key], base[key][:6], equal_nan=True), key
        engine.deviations_compute_async(True)
        engine.deviations_compute_async(True)
        got2 = engine.deviations_download()
        for key in ("dev", "z", "variability"):
            assert np.array_equal(got2[key], got[key], equal_nan=True), key
        # an option change too
        engine.set_option("dense_fp4", 0)
        engine.deviations_compute_async(True)
        got3 = engine.deviations_download()
        assert engine.stats()["operand_bits"] == 8
        engine.deviations_compute_async(True)
        got4 = engine.deviations_download()
        for key in ("dev", "z"):
            assert np.array_equal(got4[key], got3[key], equal_nan=True), key
    finally:
        engine.set_option("path", 0)
        engine.set_option("dense_fp4", -1)


def test_async_with_a_fallback_stays_on_the_checked_call(engine):
    """A list that repeats a peak makes the checked call rebuild the operand rows by scatter: such a
    verdict is not replayed, the async entry point keeps running the checked call."""
    C, P, N, bg, e, ptr, idx = _problem()
    idx = idx.copy()
    idx[1] = idx[0]                                                # annotation 0 lists its first peak twice
    engine.set_option("path", 2)
    try:
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
        engine.deviations_upload(ptr, idx, bg, e, index_base=1)
        engine.deviations_compute(True)
        base = engine.deviations_download()
        for _ in range(3):
            engine.deviations_compute_async(True)
        _same(engine.deviations_download(), base)
        # new counts: checked call again, then replays on the new matrix
        C2 = sp.csc_matrix((np.minimum(C.data * 2, 9), C.indices, C.indptr), shape=C.shape)
        engine.set_counts_csc(P, N, C2.indptr, C2.indices, C2.data)
        engine.deviations_upload(ptr, idx, bg, e, index_base=1)
        engine.deviations_compute_async(True)
        a = engine.deviations_download()
        engine.deviations_compute_async(True)
        _same(engine.deviations_download(), a)
        assert not np.array_equal(a["z"], base["z"], equal_nan=True)
    finally:
        engine.set_option("path", 0)


def test_async_after_an_e2m1_veto_stays_on_the_checked_call(engine):
    """Every nonzero is 11 = 6 + 4 + 1: the overflow reserve of the e2m1 operands is exceeded, the
    checked call redoes itself on u8 operands under a veto that ends with the call.  Such a call leaves
    no verdict to replay, so the async entry point keeps going through the checked call."""
    C, P, N, bg, e, ptr, idx = _problem()
    C = sp.csc_matrix((np.full(C.nnz, 11.0), C.indices, C.indptr), shape=C.shape)
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    engine.set_option("path", 2)
    engine.set_option("dense_fp4", 1)
    try:
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
        engine.deviations_upload(ptr, idx, bg, e, index_base=1)
        engine.deviations_compute(True)
        base = engine.deviations_download()
        assert engine.stats()["operand_bits"] == 8
        for _ in range(3):
            engine.deviations_compute_async(True)
        engine.deviations_wait()
        assert engine.stats()["operand_bits"] == 8
        _same(engine.deviations_download(), base)
    finally:
        engine.set_option("path", 0)
        engine.set_option("dense_fp4", -1)
