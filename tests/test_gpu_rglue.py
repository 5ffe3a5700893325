"""The .Call glue (rpkg/src/cv_rglue.c) on the GPU, driven through its REGISTERED routine table by the
stand-in R runtime (rpkg/rstub/) with the argument types and order of the R shim
(rpkg/R/chromvar_b200.R): the binding that mirrors src/RcppExports.cpp:59-137 / R/RcppExports.R:20-38.

Covers the reference's own known-answer test (tests/testthat/test_compute_deviations.R:34-88,
inst/extdata/deviations_test), the bundled mini data, the background sampler, row_sds / row_sds_perm,
counts passed as several column blocks, several shards, and the stop() paths.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402
from test_gpu_parity import assert_close  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture
def R():
    """One R session per test: the glue keeps ONE engine per process (like the reference's package
    state), so cvb_init() here replaces whatever an earlier test left."""
    from rcall import RSession
    s = RSession()
    s.call("cvb_init", np.array([0], np.int32), 12345.0)
    yield s
    s.call("cvb_shutdown")


def _anno(sets):
    """what the shim passes: annoptr as double, the concatenated 1-based indices as integer"""
    return (np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.float64),
            np.concatenate([np.asarray(s) for s in sets]).astype(np.int32))


def test_reference_known_answer_through_dot_call(R, golden_deviations_test):
    from rcall import dgc_slots
    g = golden_deviations_test
    C = sp.csc_matrix(g["counts"])
    p, i, x = dgc_slots(C)
    annoptr, annoidx = _anno([g["set0"], g["set1"], g["set2"]])
    bg_double = np.asarray(g["bg"], np.float64)          # getBackgroundPeaks returns storage mode double
    e = orc.compute_expectations_core(C)
    # one call on host data, default expectation (NULL) and explicit expectation
    for ex in (None, e):
        res = R.call("cvb_deviations_csc", p, i, x, 3, annoptr, annoidx, bg_double, ex)
        dev, z, matches, overlap, variab = res
        assert dev.shape == (3, 4) and z.shape == (3, 4)
        assert np.max(np.abs(dev - g["deviations"])) < 1e-12
        assert np.max(np.abs(z - g["z"])) < 1e-12
        assert np.allclose(matches, [1.0, 2 / 3, 2 / 3])
    # resident counts: cvb_set_counts -> cvb_expectations -> cvb_deviations (integer background)
    R.call("cvb_set_counts", p, i, x, 3)
    e2 = R.call("cvb_expectations", 3, False, None, 0)
    assert np.max(np.abs(e2 - e)) < 1e-15
    fr = R.call("cvb_fragments", 3, 4)
    assert np.array_equal(fr[0], np.asarray(C.sum(axis=1)).ravel()) and fr[2][0] == C.sum()
    res = R.call("cvb_deviations", annoptr, annoidx, np.asarray(g["bg"], np.int32), e2, 4)
    assert np.max(np.abs(res[0] - g["deviations"])) < 1e-12 and np.max(np.abs(res[1] - g["z"])) < 1e-12
    # base-matrix input
    R.call("cvb_set_counts_dense", np.asarray(g["counts"], np.float64))
    res = R.call("cvb_deviations", annoptr, annoidx, bg_double, e2, 4)
    assert np.max(np.abs(res[1] - g["z"])) < 1e-12


def test_mini_pipeline_through_dot_call(R, golden_mini):
    from rcall import dgc_slots
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    A = sp.csc_matrix((g["ix_x"].astype(bool), g["ix_i"], g["ix_p"]), shape=tuple(g["ix_dim"]))
    P, N = C.shape
    sets = orc.convert_to_ix_list(A)
    annoptr, annoidx = _anno(sets)
    p, i, x = dgc_slots(C)
    R.call("cvb_set_counts", p, i, x, P)
    R.call("cvb_set_seed", 4711.0)
    bg = R.call("cvb_background_peaks", g["bias"], 50, 0.1, 50)
    assert bg.dtype == np.float64 and bg.shape == (P, 50) and bg.min() >= 1 and bg.max() <= P
    assert np.array_equal(bg, np.floor(bg))
    R.call("cvb_set_seed", 4711.0)
    assert np.array_equal(R.call("cvb_background_peaks", g["bias"], 50, 0.1, 50), bg)
    e = R.call("cvb_expectations", P, False, None, 0)
    dev, z, matches, overlap, variab = R.call("cvb_deviations", annoptr, annoidx, bg, e, N)
    ref = orc.compute_deviations_core(C, sets, bg.astype(np.int64), e)
    assert_close(dev, ref["dev"], "deviations")
    assert_close(z, ref["z"], "z")
    assert np.array_equal(matches, g["fractionMatches"])
    assert_close(overlap, ref["overlap"], "overlap")
    assert np.corrcoef(z.ravel(), g["z"].ravel())[0, 1] > 0.95          # bundled mini_dev, statistically
    # computeExpectations(norm = TRUE) and group variants (factor codes are 1-based in R)
    grp = (np.arange(N) % 3) + 1
    en = R.call("cvb_expectations", P, True, None, 0)
    eg = R.call("cvb_expectations", P, False, grp.astype(np.int32), 3)
    assert np.max(np.abs(en - orc.compute_expectations_core(C, norm=True))) < 1e-12
    assert np.max(np.abs(eg - orc.compute_expectations_core(C, group=grp))) < 1e-12
    # row_sds / row_sds_perm on the z matrix
    sd = R.call("cvb_row_sds", z, True)
    assert_close(sd, orc.row_sds(z, True), "row_sds")
    assert_close(sd, variab, "fused variability")
    boot = R.call("cvb_row_sds_perm", z, True, 200)
    assert boot.shape == (200, 3)
    assert np.all(np.abs(boot.mean(axis=0) - sd) < 0.2 * sd)
    # the kNN sampler of makePermutedSets
    bgk = R.call("cvb_background_peaks_knn", g["bias"], 20, 10)
    assert bgk.shape == (P, 20) and bgk.min() >= 1 and bgk.max() <= P


def test_column_blocks_and_shards_through_dot_call(golden_mini):
    """An atlas as a list of dgCMatrix blocks, three shards (three engines on this device)."""
    from rcall import RSession, dgc_slots
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    A = sp.csc_matrix((g["ix_x"].astype(bool), g["ix_i"], g["ix_p"]), shape=tuple(g["ix_dim"]))
    P, N = C.shape
    sets = orc.convert_to_ix_list(A)
    annoptr, annoidx = _anno(sets)
    S = RSession()
    try:
        S.call("cvb_init", np.array([0, 0, 0], np.int32), 99.0)
        assert S.call("cvb_n_devices")[0] == 3
        blocks = [C[:, :11].tocsc(), C[:, 11:12].tocsc(), C[:, 12:].tocsc()]
        p, i, x = dgc_slots(blocks)
        S.call("cvb_set_counts", p, i, x, P)
        fr = S.call("cvb_fragments", P, N)
        assert np.array_equal(fr[0], np.asarray(C.sum(axis=1)).ravel())
        assert np.array_equal(fr[1], np.asarray(C.sum(axis=0)).ravel())
        e = S.call("cvb_expectations", P, False, None, 0)
        bg = S.call("cvb_background_peaks", g["bias"], 50, 0.1, 50)
        ref = orc.compute_deviations_core(C, sets, bg.astype(np.int64), e)
        for res in (S.call("cvb_deviations", annoptr, annoidx, bg, e, N),
                    S.call("cvb_deviations_csc", p, i, x, P, annoptr, annoidx, bg, e),
                    S.call("cvb_deviations_csc", p, i, x, P, annoptr, annoidx, bg, None)):
            assert res[0].shape == (3, N)
            assert_close(res[0], ref["dev"], "deviations")
            assert_close(res[1], ref["z"], "z")
            assert_close(res[4], orc.row_sds(ref["z"], True), "variability over all shards")
        # the norm / group variants over the three shards, and a base matrix split into column slabs
        grp = (np.arange(N) % 3).astype(np.int32) + 1
        assert np.allclose(S.call("cvb_expectations", P, True, None, 0), orc.compute_expectations_core(C, norm=True),
                           rtol=1e-12, atol=0)
        for nrm in (False, True):
            assert np.allclose(S.call("cvb_expectations", P, nrm, grp, 3),
                               orc.compute_expectations_core(C, norm=nrm, group=grp - 1), rtol=1e-12, atol=0)
        S.call("cvb_set_counts_dense", np.asfortranarray(C.toarray()))
        fr = S.call("cvb_fragments", P, N)
        assert np.array_equal(fr[0], np.asarray(C.sum(axis=1)).ravel())
        assert np.array_equal(fr[1], np.asarray(C.sum(axis=0)).ravel())
        res = S.call("cvb_deviations", annoptr, annoidx, bg, e, N)
        assert_close(res[1], ref["z"], "z (dense input, three shards)")
    finally:
        S.call("cvb_shutdown")


def test_stop_paths_through_dot_call(R, golden_mini):
    from rcall import RError, dgc_slots
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    P, N = C.shape
    p, i, x = dgc_slots(C)
    bg = np.random.default_rng(0).integers(1, P + 1, size=(P, 50)).astype(np.float64)
    e = orc.compute_expectations_core(C)
    annoptr, annoidx = _anno([[1, 2, 3], [5, 9]])
    # a peak index outside 1..P
    bad_ptr, bad_idx = _anno([[1, 2, P + 1]])
    with pytest.raises(RError, match="Annotations are not valid"):
        R.call("cvb_deviations_csc", p, i, x, P, bad_ptr, bad_idx, bg, e)
    # a background matrix with a fractional entry, and one with the wrong number of rows
    bgf = bg.copy()
    bgf[3, 3] += 0.5
    with pytest.raises(RError, match="whole peak indices"):
        R.call("cvb_deviations_csc", p, i, x, P, annoptr, annoidx, bgf, e)
    with pytest.raises(RError, match="nrow\\(counts_mat\\) == nrow\\(background_peaks\\)"):
        R.call("cvb_deviations_csc", p, i, x, P, annoptr, annoidx, bg[:-1], e)
    with pytest.raises(RError, match="length\\(expectation\\)"):
        R.call("cvb_deviations_csc", p, i, x, P, annoptr, annoidx, bg, e[:-1])
    # a peak without fragments: the reference's message
    C0 = C.tolil()
    C0[7, :] = 0
    C0 = C0.tocsc()
    p0, i0, x0 = dgc_slots(C0)
    with pytest.raises(RError, match="All peaks must have at least one fragment in one sample"):
        R.call("cvb_deviations_csc", p0, i0, x0, P, annoptr, annoidx, bg, e)
    # negative counts: counts_check
    xn = [x[0].copy()]
    xn[0][5] = -1.0
    with pytest.raises(RError, match="counts"):
        R.call("cvb_set_counts", p, i, xn, P)
    # slots that do not fit together
    with pytest.raises(RError, match="@i and @x differ in length"):
        R.call("cvb_set_counts", p, i, [x[0][:-1]], P)
    # and the session is still usable afterwards
    res = R.call("cvb_deviations_csc", p, i, x, P, annoptr, annoidx, bg, e)
    assert res[0].shape == (2, N)
