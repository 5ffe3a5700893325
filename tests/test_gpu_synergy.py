"""GPU parity tests of the batched callers of the single-set kernel: synergy,
correlation, covariability, permuted data -- against the oracle's literal restatements
(parity unpinned: the reference holds no test for these)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402
from test_gpu_parity import assert_close, _mini  # noqa: E402

pytestmark = pytest.mark.gpu


def _inputs(golden_mini, extra_sets=True):
    C, A = _mini(golden_mini)
    sets = [np.asarray(s) for s in orc.convert_to_ix_list(A)]
    if extra_sets:
        rng = np.random.default_rng(5)
        sets.append(np.sort(rng.choice(1000, 150, replace=False)) + 1)
        sets.append(np.unique(np.concatenate([sets[0][:100], sets[1]])))
    e = orc.compute_expectations_core(C)
    bg = orc.get_background_peaks_core(C, golden_mini["bias"], rng=np.random.default_rng(3))
    return C, sets, bg, e


def test_variability_batch(engine, golden_mini):
    from chromvar_b200 import synergy
    C, sets, bg, e = _inputs(golden_mini)
    got = synergy.compute_variability_batch(C, sets, bg, e)
    ref = np.array([orc.compute_variability_single(s, C, bg, e) for s in sets])
    assert_close(got, ref, "variability")


def test_synergy_matches_oracle(engine, golden_mini):
    from chromvar_b200 import synergy
    C, sets, bg, e = _inputs(golden_mini)
    got, names = synergy.getAnnotationSynergy(C, sets, bg, e, nbg=6, rng=np.random.default_rng(77))
    ref, order = orc.get_annotation_synergy_core(C, sets, bg, e, nbg=6, rng=np.random.default_rng(77))
    assert names == [str(k + 1) for k in order]
    assert got.shape == (5, 5) and np.array_equal(np.diag(got), np.zeros(5))
    # a set that contains the whole top set makes every background draw a permutation of that set:
    # sd(bgvar) is 0 up to summation order and the boost is 0/0 (noise in the reference as well) --
    # such entries are excluded, every other entry must agree
    degenerate = np.zeros((5, 5), bool)
    for i in range(4):
        top = sets[order[i]]
        for jj, k in enumerate(order[i + 1:]):
            if np.isin(top, sets[k]).all():
                degenerate[i, i + 1 + jj] = degenerate[i + 1 + jj, i] = True
    assert degenerate.sum() == 2
    assert_close(np.where(degenerate, 0.0, got), np.where(degenerate, 0.0, ref), "synergy")
    # variabilities given (the data.frame path) reproduce the same matrix
    var = {"variability": synergy.compute_variability_batch(C, sets, bg, e)}
    again, _ = synergy.getAnnotationSynergy(C, sets, bg, e, variabilities=var, nbg=6, rng=np.random.default_rng(77))
    assert np.array_equal(again, got, equal_nan=True)


def test_correlation_matches_oracle(engine, golden_mini):
    from chromvar_b200 import synergy
    C, sets, bg, e = _inputs(golden_mini)
    sets.append(sets[0].copy())                       # identical sets: both differences empty -> NA
    got, names = synergy.getAnnotationCorrelation(C, dict(zip("abcdef", sets)), bg, e)
    ref, order = orc.get_annotation_correlation_core(C, sets, bg, e)
    assert names == ["abcdef"[k] for k in order]
    assert np.array_equal(np.diag(got), np.ones(6))
    # NA wherever a difference set is empty (identical sets, subsets) or has expected < 1 in some cell
    ia, i_f = names.index("a"), names.index("f")
    assert np.isnan(got[ia, i_f]) and np.array_equal(np.isnan(got), np.isnan(ref))
    assert 0 < np.isnan(got).sum() < 30
    assert_close(got, ref, "correlation")


def test_covariability_and_permuted_data(engine, golden_mini):
    from chromvar_b200 import api, synergy
    C, A = _mini(golden_mini)
    counts = api.SummarizedExperiment({"counts": C}, rowData={"bias": golden_mini["bias"]})
    ix = api.SummarizedExperiment({"matches": A})
    api.set_seed(11)
    bg = api.getBackgroundPeaks(counts)
    dev = api.computeDeviations(counts, ix, bg)
    got = synergy.deviationsCovariability(dev)
    assert_close(got, orc.deviations_covariability(api.deviationScores(dev)), "covariability")
    # permuted data: every permuted column is a row-gather of the original column by a valid
    # background matrix; the two iterations use different draws; fragments per sample mostly survive
    api.set_seed(12)
    perm = synergy.getPermutedData(counts, niterations=2)
    assert set(perm.assays) == {"counts", "perm_1", "perm_2"}
    p1, p2 = perm.assays["perm_1"], perm.assays["perm_2"]
    assert p1.shape == C.shape and (p1 != p2).nnz > 0
    # reproduce iteration 1 draw for draw: same seed stream -> same background matrix
    api.set_seed(12)
    api.get_engine().set_seed(api._next_sampler_seed())
    bgN = synergy._bg_for_perm(counts, C.shape[1], 0.1, 50)
    ref = orc.reorder_columns(C, bgN)
    assert (ref != p1).nnz == 0
    invalid = (bgN < 1) | (bgN > C.shape[0])
    assert not invalid.any()
    assert p1.has_sorted_indices and np.all(np.diff(p1.indptr) >= 0) and p1.data.min() >= 1
    # covariability of the z still resident on the device == of the host copy; a larger K x K case
    eng = api.get_engine()
    rng = np.random.default_rng(3)
    Z = rng.normal(size=(150, 333))
    Z[7, 5] = np.nan                                   # cov(use = "everything"): row / column 7 become NA
    got = eng.deviations_covariability(Z)
    ref = orc.deviations_covariability(Z)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert_close(got, ref, "covariability 150 x 150")
