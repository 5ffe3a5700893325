"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and through the R-API
mirror, against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): integer observed / background sums bit-exact; deviations and
z-scores within 1e-5 relative (abs(d) <= 1e-5 * max(abs(ref), 1e-3)); identical NA masks.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def assert_close(got, ref, what=""):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), "%s: NA/NaN masks differ (%d vs %d)" % (what, nan_g.sum(), nan_r.sum())
    inf_g, inf_r = np.isinf(got), np.isinf(ref)
    assert np.array_equal(inf_g, inf_r) and np.array_equal(got[inf_g], ref[inf_r]), what + ": inf mismatch"
    ok = ~(nan_r | inf_r)
    err = np.abs(got[ok] - ref[ok])
    tol = RTOL * np.maximum(np.abs(ref[ok]), 1e-3)
    assert np.all(err <= tol), "%s: max rel err %.3g" % (what, float((err / np.maximum(np.abs(ref[ok]), 1e-3)).max()))


def sets_to_csr(sets):
    ptr = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    idx = np.concatenate([np.asarray(s, np.int32) for s in sets]) if ptr[-1] else np.zeros(0, np.int32)
    return ptr, idx.astype(np.int32)


def run_both(engine, C, sets, bg, e, sums=True):
    Cc = sp.csc_matrix(C)
    engine.set_counts_csc(Cc.shape[0], Cc.shape[1], Cc.indptr, Cc.indices, Cc.data)
    ptr, idx = sets_to_csr(sets)
    got = engine.deviations(ptr, idx, bg, e, index_base=1, sums=sums)
    ref = orc.compute_deviations_core(Cc, [np.asarray(s) for s in sets], bg, e, intermediate=sums) \
        if all(len(s) for s in sets) else None
    return got, ref


def check_against_oracle(got, ref, sums=True):
    if sums:
        assert np.array_equal(got["observed"], ref["observed"].astype(np.int64)), "observed sums differ"
        assert np.array_equal(got["sampled"], ref["sampled"].astype(np.int64)), "background sums differ"
    assert_close(got["dev"], ref["dev"], "deviations")
    assert_close(got["z"], ref["z"], "z")
    assert np.allclose(got["matches"], ref["matches"], rtol=1e-15, atol=0)
    assert_close(got["overlap"], ref["overlap"], "overlap")


# --------------------------------------------------------------------------------------------
# golden vector 1: the reference's own known-answer test (7 input-type combinations)
# --------------------------------------------------------------------------------------------
def _golden_inputs(g):
    from chromvar_b200 import api
    mat = g["counts"]
    anno_list = [g["set0"], g["set1"], g["set2"]]
    anno_mat = sp.csc_matrix((np.ones(7, bool), ([0, 1, 2, 0, 1, 1, 2], [0, 0, 0, 1, 1, 2, 2])), shape=(3, 3))
    counts_se = api.SummarizedExperiment({"counts": sp.csc_matrix(mat)})
    anno_se = api.SummarizedExperiment({"matches": anno_mat})
    return mat, anno_list, anno_mat, counts_se, anno_se, g["bg"]


@pytest.mark.parametrize("combo", ["se_se", "se_list", "se_matrix", "matrix_se", "matrix_list",
                                   "matrix_matrix", "Matrix_Matrix"])
def test_reference_known_answer(golden_deviations_test, combo):
    """tests/testthat/test_compute_deviations.R:34-88 against inst/extdata/deviations_test,
    with the reference's tolerance (expect_equal, 1.5e-8) tightened to 1e-12."""
    from chromvar_b200 import api
    g = golden_deviations_test
    mat, anno_list, anno_mat, counts_se, anno_se, bg = _golden_inputs(g)
    obj = {"se": counts_se, "matrix": mat, "Matrix": sp.csc_matrix(mat)}[combo.split("_")[0]]
    anno = {"se": anno_se, "list": anno_list, "matrix": anno_mat, "Matrix": anno_mat}[combo.split("_")[1]]
    api.invalidate_counts_cache()
    dev = api.computeDeviations(obj, anno, bg)
    assert isinstance(dev, api.chromVARDeviations)
    assert np.max(np.abs(api.deviations(dev) - g["deviations"])) < 1e-12
    assert np.max(np.abs(api.deviationScores(dev) - g["z"])) < 1e-12
    assert dev.rownames == ["1", "2", "3"]


# --------------------------------------------------------------------------------------------
# config 1: bundled mini_counts x mini_ix
# --------------------------------------------------------------------------------------------
def _mini(g):
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    A = sp.csc_matrix((g["ix_x"].astype(bool), g["ix_i"], g["ix_p"]), shape=tuple(g["ix_dim"]))
    return C, A


def test_mini_exact_against_oracle_with_shared_background(engine, golden_mini):
    g = golden_mini
    C, A = _mini(g)
    sets = orc.convert_to_ix_list(A)
    e = orc.compute_expectations_core(C)
    bg = orc.get_background_peaks_core(C, g["bias"], rng=np.random.default_rng(3))
    got, ref = run_both(engine, C, sets, bg, e)
    check_against_oracle(got, ref)
    assert np.array_equal(got["matches"], g["fractionMatches"])
    assert_close(got["variability"], orc.row_sds(ref["z"], True), "variability")


def test_mini_end_to_end_reproduces_bundled_mini_dev(golden_mini):
    """computeDeviations(mini_counts, mini_ix) with the GPU sampler vs the bundled mini_dev:
    statistical agreement (the reference's RNG state was not recorded)."""
    from chromvar_b200 import api
    g = golden_mini
    C, A = _mini(g)
    counts = api.SummarizedExperiment({"counts": C}, rowData={"bias": g["bias"]},
                                      colData={"Cell_Type": g["cell_type"], "depth": g["depth"]})
    ix = api.SummarizedExperiment({"matches": A}, colData={"name": g["ix_names"]})
    api.invalidate_counts_cache()
    api.set_seed(2017)
    dev = api.computeDeviations(counts, ix)
    z, d = api.deviationScores(dev), api.deviations(dev)
    assert z.shape == (3, 35) and not np.isnan(z).any()
    assert np.array_equal(dev.rowData["fractionMatches"], g["fractionMatches"])
    assert np.max(np.abs(dev.rowData["fractionBackgroundOverlap"] - g["fractionBackgroundOverlap"])) < 0.03
    assert np.corrcoef(z.ravel(), g["z"].ravel())[0, 1] > 0.95
    assert np.corrcoef(d.ravel(), g["dev"].ravel())[0, 1] > 0.95
    var = api.computeVariability(dev, bootstrap_error=True, bootstrap_samples=200)
    assert list(var["name"]) == list(g["ix_names"])
    assert np.all(var["bootstrap_lower_bound"] <= var["variability"])
    assert np.all(var["bootstrap_upper_bound"] >= var["variability"])
    assert np.max(np.abs(var["variability"] - [0.64613, 1.03573, 0.89700])) < 0.2


# --------------------------------------------------------------------------------------------
# example_counts (29269 x 50): many random sets, edge cases
# --------------------------------------------------------------------------------------------
def _example(g):
    return sp.csc_matrix((g["counts_x"].astype(np.float64), g["counts_i"], g["counts_p"]),
                         shape=tuple(g["counts_dim"]))


def test_example_counts_random_sets(engine, golden_example_counts):
    C = _example(golden_example_counts)
    keep = np.asarray(C.sum(axis=1)).ravel() > 0
    C = C[np.nonzero(keep)[0], :].tocsc()
    P = C.shape[0]
    rng = np.random.default_rng(42)
    sizes = [1, 1, 2, 3, 7, 40, 300, 2000, 9000, P]
    sets = [np.sort(rng.choice(P, s, replace=False)) + 1 for s in sizes]
    sets.append(np.array([5, 5, 5, 9, 12]))                 # duplicates keep their multiplicity
    sets.append(rng.integers(1, P + 1, 500))                # unsorted with repeats
    bias = rng.uniform(0.3, 0.8, P)
    bg = orc.get_background_peaks_core(C, bias, niterations=50, rng=np.random.default_rng(8))
    e = orc.compute_expectations_core(C)
    got, ref = run_both(engine, C, sets, bg, e)
    check_against_oracle(got, ref)
    # small sets give expected < 1 for shallow cells: the NA mask must be R's NA_real_ payload
    from chromvar_b200._lib import is_na
    fail = _fail_mask(C, sets, e)
    assert fail.any() and np.array_equal(is_na(got["z"]), fail) and np.array_equal(is_na(got["dev"]), fail)


def _fail_mask(C, sets, e):
    fps = np.asarray(C.sum(axis=0)).ravel()
    out = np.zeros((len(sets), C.shape[1]), bool)
    for k, s in enumerate(sets):
        out[k] = e[np.asarray(s) - 1].sum() * fps < 1
    return out


def test_empty_set_and_invalid_annotations(engine, golden_example_counts):
    from chromvar_b200 import _lib
    C = _example(golden_example_counts)[:2000, :].tocsc()
    C = C[np.nonzero(np.asarray(C.sum(axis=1)).ravel() > 0)[0], :].tocsc()
    P, N = C.shape
    rng = np.random.default_rng(0)
    bg = rng.integers(1, P + 1, (P, 10)).astype(np.int32)
    e = orc.compute_expectations_core(C)
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    ptr, idx = sets_to_csr([[1, 2, 3], [], [7]])
    got = engine.deviations(ptr, idx, bg, e)
    assert _lib.is_na(got["z"][1]).all() and _lib.is_na(got["dev"][1]).all()    # :458-461
    assert got["matches"][1] == 0 and _lib.is_na(got["overlap"][1:2]).all()
    ref = orc.compute_deviations_core(C, [[1, 2, 3], [7]], bg, e)
    assert_close(got["z"][[0, 2]], ref["z"], "z around an empty set")
    for bad in ([0], [P + 1], [-3]):
        ptr, idx = sets_to_csr([[1, 2], bad])
        with pytest.raises(_lib.ChromVARError) as ei:
            engine.deviations(ptr, idx, bg, e)
        assert ei.value.code == _lib.CV_ERR_INVALID and "not valid" in str(ei.value)
    bad_bg = bg.copy()
    bad_bg[3, 2] = P + 5
    ptr, idx = sets_to_csr([[1, 2]])
    with pytest.raises(_lib.ChromVARError):
        engine.deviations(ptr, idx, bad_bg, e)


def test_zero_peak_is_rejected(engine):
    from chromvar_b200 import api
    C = np.array([[1, 0, 2], [0, 0, 0], [3, 1, 0]], float)
    api.invalidate_counts_cache()
    with pytest.raises(ValueError, match="All peaks must have at least one fragment"):
        api.computeDeviations(C, [[1, 2]], np.ones((3, 4), np.int32), np.full(3, 1 / 3))
    with pytest.raises(ValueError, match="All peaks must have at least one fragment"):
        api.getBackgroundPeaks(C, bias=np.array([0.4, 0.5, 0.6]))


def test_counts_validation(engine):
    from chromvar_b200 import _lib
    ptr = np.array([0, 2, 3], np.int64)
    idx = np.array([0, 1, 1], np.int32)
    # counts_check (R/utils.R:3-12): min(counts) >= 0 -- negative and non-finite values
    for bad in ([1.0, -1.0, 2.0], [1.0, np.nan, 2.0], [1.0, np.inf, 2.0]):
        with pytest.raises(_lib.ChromVARError) as ei:
            engine.set_counts_csc(2, 2, ptr, idx, np.array(bad))
        assert ei.value.code == _lib.CV_ERR_INVALID and "counts_check" in str(ei.value)
    # fractional counts pass the reference's check; this engine (exact integer sums) does not take
    # them and says so in its own words
    with pytest.raises(_lib.ChromVARError) as ei:
        engine.set_counts_csc(2, 2, ptr, idx, np.array([1.0, 0.5, 2.0]))
    assert ei.value.code == _lib.CV_ERR_UNSUPPORTED and "fractional" in str(ei.value) and "counts_check" not in str(ei.value)
    with pytest.raises(_lib.ChromVARError):
        engine.set_counts_csc(2, 2, ptr, np.array([0, 5, 1], np.int32), np.ones(3))


# --------------------------------------------------------------------------------------------
# K1: fragment sums and expectations (default / norm / group), sparse and dense input
# --------------------------------------------------------------------------------------------
def test_fragments_and_expectations(engine, golden_example_counts):
    C = _example(golden_example_counts)
    P, N = C.shape
    for xdtype in (np.float64, np.int32, np.uint8, np.float32):
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data.astype(xdtype))
        fpp, fps, tot = engine.fragments()
        assert np.array_equal(fpp, orc.get_fragments_per_peak(C))
        assert np.array_equal(fps, orc.get_fragments_per_sample(C))
        assert tot == orc.get_total_fragments(C)
    engine.set_counts_csc(P, N, C.indptr.astype(np.int64), C.indices, C.data)
    assert np.array_equal(engine.expectations(), orc.compute_expectations_core(C))
    assert np.allclose(engine.expectations(norm=True), orc.compute_expectations_core(C, norm=True), rtol=1e-12)
    grp = np.random.default_rng(1).integers(0, 4, N)
    assert np.allclose(engine.expectations(group=grp, n_groups=4),
                       orc.compute_expectations_core(C, group=grp), rtol=1e-12)
    assert np.allclose(engine.expectations(norm=True, group=grp, n_groups=4),
                       orc.compute_expectations_core(C, norm=True, group=grp), rtol=1e-12)


def test_dense_input_equals_sparse_input(engine):
    rng = np.random.default_rng(9)
    D = rng.negative_binomial(3, 0.02, (700, 13)).astype(np.float64)    # bulk-like counts, hundreds
    D[rng.random(D.shape) < 0.2] = 0
    D[D.sum(axis=1) == 0, 0] = 1
    P, N = D.shape
    sets = [np.sort(rng.choice(P, s, replace=False)) + 1 for s in (1, 5, 60, 300)]
    bg = rng.integers(1, P + 1, (P, 20)).astype(np.int32)
    e = orc.compute_expectations_core(D)
    ptr, idx = sets_to_csr(sets)
    engine.set_counts_dense(D)
    fpp, fps, tot = engine.fragments()
    assert np.array_equal(fpp, D.sum(axis=1)) and np.array_equal(fps, D.sum(axis=0))
    got_d = engine.deviations(ptr, idx, bg, e, sums=True)
    got_s, ref = run_both(engine, sp.csc_matrix(D), sets, bg, e)
    check_against_oracle(got_d, ref)
    for key in ("observed", "sampled", "dev", "z"):
        assert np.array_equal(got_d[key], got_s[key], equal_nan=True)


def test_large_counts_use_wide_accumulators(engine):
    """Counts large enough that (largest set) x (largest count) >= 2^31: the library switches to
    64-bit accumulators; sums stay exact."""
    rng = np.random.default_rng(4)
    P, N = 400, 40
    D = rng.integers(0, 6_000_000, (P, N)).astype(np.float64)    # values >= 2^20 are split in the layout
    D[D.sum(axis=1) == 0, 0] = 1
    bg = rng.integers(1, P + 1, (P, 8)).astype(np.int32)
    e = orc.compute_expectations_core(D)
    full = np.concatenate([np.arange(1, P + 1)] * 3)             # every peak with multiplicity 3
    sets2 = [full, np.sort(rng.choice(P, 100, replace=False)) + 1]
    ptr, idx = sets_to_csr(sets2)
    engine.set_counts_dense(D)
    got = engine.deviations(ptr, idx, bg, e, sums=True)
    assert engine.stats()["acc_bytes"] == 8
    ref = orc.compute_deviations_core(sp.csc_matrix(D), sets2, bg, e, intermediate=True)
    check_against_oracle(got, ref)
    assert got["observed"].max() > 2 ** 31


def test_properties_at_size(engine):
    """Size-independent properties on a 20000 x 2000 synthetic matrix: linearity over disjoint
    sets, independence of work order / tiling, layout rebuild idempotence."""
    from chromvar_b200 import synthetic
    d = synthetic.make_config("small")
    C = d["counts"]
    P, N = C.shape
    rng = np.random.default_rng(2)
    bg = synthetic.random_background(P, 50, rng)
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    perm = rng.permutation(P) + 1
    a, b = np.sort(perm[:3000]), np.sort(perm[3000:7000])
    sets = [a, b, np.concatenate([a, b])]
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    ptr, idx = sets_to_csr(sets)
    got = engine.deviations(ptr, idx, bg, e, sums=True)
    assert np.array_equal(got["observed"][0] + got["observed"][1], got["observed"][2])
    assert np.array_equal(got["sampled"][0] + got["sampled"][1], got["sampled"][2])
    # checksum of checksums: sum over cells of observed = sum of per-peak totals over the set
    fpp = np.asarray(C.sum(axis=1)).ravel()
    assert got["observed"][0].sum() == fpp[a - 1].sum()
    assert got["sampled"][0].sum() == sum(fpp[bg[a - 1, i] - 1].sum() for i in range(50))
    # same answer when recomputed with the counts layout rebuilt, and for a subset of sets
    engine.deviations_upload(ptr, idx, bg, e)
    engine.deviations_compute(rebuild_counts_layout=True, keep_sums=True)
    again = engine.deviations_download(sums=True)
    for key in ("observed", "sampled", "dev", "z", "variability"):
        assert np.array_equal(again[key], got[key], equal_nan=True), key
    ptr1, idx1 = sets_to_csr([b])
    one = engine.deviations(ptr1, idx1, bg, e)
    assert np.array_equal(one["z"][0], got["z"][1], equal_nan=True)
    # oracle on a slice of cells (the oracle is slow at full size)
    cols = np.sort(rng.choice(N, 64, replace=False))
    Cs = sp.csr_matrix(C[:, cols])      # a cell slice leaves some peaks empty: skip the core's check
    res = [orc.compute_deviations_single(s, Cs, bg, e, intermediate_results=True) for s in sets]
    ref = dict(observed=np.vstack([r["observed"] for r in res]), sampled=np.stack([r["sampled"] for r in res]),
               dev=np.vstack([r["dev"] for r in res]), z=np.vstack([r["z"] for r in res]))
    assert np.array_equal(got["observed"][:, cols], ref["observed"].astype(np.int64))
    assert np.array_equal(got["sampled"][:, :, cols], ref["sampled"].astype(np.int64))
    # the expectation is an input, so dev / z of a cell depend only on that cell's column
    assert_close(got["dev"][:, cols], ref["dev"], "dev (cell slice)")
    assert_close(got["z"][:, cols], ref["z"], "z (cell slice)")


def test_niterations_other_than_50(engine, golden_mini):
    g = golden_mini
    C, A = _mini(g)
    sets = orc.convert_to_ix_list(A)
    e = orc.compute_expectations_core(C)
    for B in (2, 31, 32, 33, 64, 200):
        bg = np.random.default_rng(B).integers(1, 1001, (1000, B)).astype(np.int32)
        got, ref = run_both(engine, C, sets, bg, e)
        check_against_oracle(got, ref)


@pytest.mark.parametrize("B", [207, 208, 255, 256, 300, 1000])
def test_many_iterations(engine, golden_mini, B):
    """niterations around and beyond what one accumulator tile holds (e2m1: B + 1 <= 208 rows, u8: 256):
    the automatic choice must always produce the reference's numbers, whichever kernel it takes."""
    g = golden_mini
    C, A = _mini(g)
    sets = orc.convert_to_ix_list(A)
    e = orc.compute_expectations_core(C)
    bg = np.random.default_rng(B).integers(1, 1001, (1000, B)).astype(np.int32)
    got, ref = run_both(engine, C, sets, bg, e)
    check_against_oracle(got, ref)
    if B + 1 <= 256:                                     # the dense kernels, forced
        engine.set_option("path", 2)
        try:
            got, _ = run_both(engine, C, sets, bg, e)
            assert engine.stats()["path"] == 1
            check_against_oracle(got, ref)
        finally:
            engine.set_option("path", 0)


def test_no_annotations_runs_on_individual_peaks(golden_mini, capsys):
    """computeDeviations(object) without annotations (R/compute_deviations.R:284-296, :340-353): every
    peak is its own set -- the reference's tf_count == 1 branch (:467-478), K = P rows."""
    from chromvar_b200 import api
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    P, N = C.shape
    counts = api.SummarizedExperiment({"counts": C}, rowData={"bias": g["bias"]})
    api.invalidate_counts_cache()
    api.set_seed(5)
    bg = api.getBackgroundPeaks(counts)
    e = api.computeExpectations(counts)
    dev = api.computeDeviations(counts, background_peaks=bg, expectation=e)
    assert "individual peaks" in capsys.readouterr().out
    z, d = api.deviationScores(dev), api.deviations(dev)
    assert z.shape == (P, N) and dev.rownames == [str(i + 1) for i in range(P)]
    pick = np.random.default_rng(0).choice(P, 60, replace=False)
    ref = orc.compute_deviations_core(C, [np.array([p + 1]) for p in pick], bg, e)
    assert_close(d[pick], ref["dev"], "deviations")
    assert_close(z[pick], ref["z"], "z")
    assert np.allclose(dev.rowData["fractionMatches"][pick], 1.0 / P)
    assert_close(dev.rowData["fractionBackgroundOverlap"][pick], ref["overlap"], "overlap")
    # the same through a plain matrix object (:340-353)
    api.invalidate_counts_cache()
    dev2 = api.computeDeviations(C, background_peaks=bg, expectation=e)
    assert np.array_equal(api.deviationScores(dev2), z, equal_nan=True)
