"""Pins the CPU oracle against every golden vector the reference holds for the hot path
(SURVEY.md section 8c): inst/extdata/deviations_test (exact) and the bundled mini_* data
(statistical, because the reference generated mini_dev with an unrecorded RNG state)."""
import numpy as np
import scipy.sparse as sp

from oracle import chromvar_oracle as orc


def _mini(g):
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    A = sp.csc_matrix((g["ix_x"].astype(bool), g["ix_i"], g["ix_p"]), shape=tuple(g["ix_dim"]))
    return C, A


def test_deviations_test_known_answer(golden_deviations_test):
    g = golden_deviations_test
    C = g["counts"]
    sets = [g["set0"], g["set1"], g["set2"]]
    e = orc.compute_expectations_core(C)
    assert np.allclose(e, C.sum(axis=1) / C.sum(), rtol=0, atol=1e-16)
    out = orc.compute_deviations_core(C, sets, g["bg"], e)
    # the reference's own tolerance is 1.5e-8 (testthat::expect_equal); we hold 1e-13
    assert np.max(np.abs(out["dev"] - g["deviations"])) < 1e-13
    assert np.max(np.abs(out["z"] - g["z"])) < 1e-13


def test_deviations_test_sparse_and_matrix_annotations(golden_deviations_test):
    g = golden_deviations_test
    C = sp.csc_matrix(g["counts"])
    anno = sp.csc_matrix((np.ones(7, bool), ([0, 1, 2, 0, 1, 1, 2], [0, 0, 0, 1, 1, 2, 2])), shape=(3, 3))
    sets = orc.convert_to_ix_list(anno)
    assert [list(s) for s in sets] == [[1, 2, 3], [1, 2], [2, 3]]
    out = orc.compute_deviations_core(C, sets, g["bg"], orc.compute_expectations_core(C))
    assert np.max(np.abs(out["z"] - g["z"])) < 1e-13


def test_mini_fixture_facts(golden_mini):
    """The numbers SURVEY.md 8c recorded from the bundled mini_counts / mini_ix."""
    g = golden_mini
    C, A = _mini(g)
    assert C.shape == (1000, 35) and C.nnz == 3590 and C.sum() == 4873
    assert list(g["counts_p"][:6]) == [0, 61, 153, 350, 448, 520]
    assert list(g["ix_p"]) == [0, 322, 410, 523]
    assert np.allclose(g["bias"][:6], [0.718, 0.458, 0.684, 0.614, 0.472, 0.54])
    assert np.allclose(g["fractionMatches"], [0.322, 0.088, 0.113])


def test_mini_whitening_check_values(golden_mini):
    g = golden_mini
    C, _ = _mini(g)
    fh = orc.background_front_half(orc.get_fragments_per_peak(C), g["bias"], 50)
    assert np.allclose(fh["cov"], [[0.1269259255, 0.0177642314], [0.0177642314, 0.0134652951]], atol=1e-9)
    assert np.allclose(fh["chol"], [[0.3562666495, 0.049862179], [0, 0.1047810013]], atol=1e-9)
    T = fh["trans_norm_mat"]
    assert abs(T[:, 0].min()) < 1e-12 and abs(T[:, 0].max() - 4.2987) < 1e-3
    assert abs(T[:, 1].min() - 1.9677) < 1e-3 and abs(T[:, 1].max() - 7.5192) < 1e-3
    assert int((fh["bin_density"] > 0).sum()) == 361 and fh["bin_density"].max() == 13
    assert fh["bin_density"].sum() == 1000


def test_mini_dev_statistical(golden_mini):
    """mini_dev is one draw of the reference's sampler: correlation and overlap fractions."""
    g = golden_mini
    C, A = _mini(g)
    sets = orc.convert_to_ix_list(A)
    e = orc.compute_expectations_core(C)
    bg = orc.get_background_peaks_core(C, g["bias"], rng=np.random.default_rng(11))
    out = orc.compute_deviations_core(C, sets, bg, e)
    assert np.array_equal(out["matches"], g["fractionMatches"])
    assert np.max(np.abs(out["overlap"] - g["fractionBackgroundOverlap"])) < 0.02
    assert np.corrcoef(out["z"].ravel(), g["z"].ravel())[0, 1] > 0.95
    assert np.corrcoef(out["dev"].ravel(), g["dev"].ravel())[0, 1] > 0.95
    sd = orc.row_sds(out["z"], True)
    assert np.max(np.abs(sd - [0.64613, 1.03573, 0.89700])) < 0.15


def test_literal_and_generative_samplers_agree(golden_mini):
    """bg_sample_helper restated literally (alias tables) vs the generative statement: same
    destination-bin distribution (chi-square on pooled bins)."""
    from scipy.stats import chi2
    g = golden_mini
    C, _ = _mini(g)
    fh = orc.background_front_half(orc.get_fragments_per_peak(C), g["bias"], 50)
    bin_p = orc.bin_weight_matrix(fh["bin_data"], 0.1)
    bm0 = fh["bin_membership"] - 1
    rng = np.random.default_rng(5)
    lit = orc.bg_sample_helper(bm0, bin_p, fh["bin_density"], 200, lambda n: rng.random(n))
    assert lit.min() >= 1 and lit.max() <= 1000
    Pbin = orc.bin_probability_matrix(bin_p, fh["bin_density"])
    src = int(np.argmax(fh["bin_density"]))
    rows = np.nonzero(bm0 == src)[0]
    dest = bm0[lit[rows].ravel() - 1]
    obs = np.bincount(dest, minlength=2500).astype(float)
    exp = Pbin[src] * dest.size
    keep = exp >= 5
    o = np.append(obs[keep], obs[~keep].sum())
    x = np.append(exp[keep], exp[~keep].sum())
    stat = ((o - x) ** 2 / np.maximum(x, 1e-300))[x > 0].sum()
    assert stat < chi2.ppf(0.9999, int((x > 0).sum()) - 1)


def test_dnorm_matches_closed_form():
    x = np.array([0.0, 0.05, 0.3, 0.49, 0.51, 1.0, 3.0, 3.9])
    ref = np.exp(-0.5 * (x / 0.1) ** 2) / (0.1 * np.sqrt(2 * np.pi))
    got = orc.dnorm(x, 0.0, 0.1)
    assert np.allclose(got, ref, rtol=1e-12, atol=0)
    assert orc.dnorm(np.array([4.0]), 0.0, 0.1)[0] == 0.0   # 40 sigma: past the underflow bound


def test_expectation_variants_small():
    rng = np.random.default_rng(0)
    C = rng.poisson(1.0, (30, 8)).astype(float)
    C[C.sum(axis=1) == 0, 0] = 1
    fps = C.sum(axis=0)
    assert np.allclose(orc.compute_expectations_core(C, norm=True), (C / fps).sum(axis=1))
    grp = np.array(list("aabbbcca"))
    exp = np.mean([C[:, grp == l].sum(axis=1) / C[:, grp == l].sum() for l in "abc"], axis=0)
    assert np.allclose(orc.compute_expectations_core(C, group=grp), exp)
    assert np.allclose(orc.compute_expectations_core(sp.csc_matrix(C), group=grp), exp)


def test_single_peak_and_empty_sets(golden_deviations_test):
    g = golden_deviations_test
    C = g["counts"]
    e = orc.compute_expectations_core(C)
    one = orc.compute_deviations_single([2], C, g["bg"], e)
    many = orc.compute_deviations_single([2, 2], C, g["bg"], e)   # duplicate = weight 2
    assert np.all(np.isfinite(one["z"]) | np.isnan(one["z"]))
    assert one["matches"] == 1 / 3 and many["matches"] == 2 / 3
    emp = orc.compute_deviations_single([], C, g["bg"], e)
    assert np.all(np.isnan(emp["z"])) and emp["matches"] == 0 and np.isnan(emp["overlap"])


def test_alternative_background_restatement(golden_mini):
    """get_background_peaks_alternative (R/test_sets.R:291-330), parity unpinned (no golden values in
    the reference): the whitening leaves an identity covariance (what chol + forwardsolve does), the
    brute-force neighbour lists agree with an independent exact kd-tree (scipy cKDTree) on every
    k-th distance, the query is never its own neighbour, draws stay inside the neighbour lists."""
    from scipy.spatial import cKDTree
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    tv = orc.alternative_whitening(orc.get_fragments_per_peak(C), g["bias"])
    assert np.allclose(np.cov(tv, rowvar=False), np.eye(2), atol=1e-12)
    for window in (1, 10, 200):
        nn = orc.knn_other(tv, window)
        assert not np.any(nn == np.arange(tv.shape[0])[:, None])
        d_ref, _ = cKDTree(tv).query(tv, k=window + 1)
        d = np.sqrt(((tv[nn] - tv[:, None, :]) ** 2).sum(-1))
        assert np.all(np.diff(d, axis=1) >= -1e-15)
        # with duplicates of the query point the kd-tree may list a twin instead of the query itself;
        # the multiset of the window + 1 smallest distances is the same either way
        mine = np.sort(np.concatenate([np.zeros((tv.shape[0], 1)), d], axis=1), axis=1)
        assert np.max(np.abs(mine - d_ref)) < 1e-12
    bg = orc.get_background_peaks_alternative(C, g["bias"], 7, 10, rng=np.random.default_rng(1))
    nn = orc.knn_other(tv, 10)
    assert bg.shape == (1000, 7)
    assert all(set(bg[p] - 1) <= set(nn[p]) for p in range(1000))
    sets = orc.make_permuted_sets(C, [[1, 2, 3], [5]], g["bias"], window=10, rng=np.random.default_rng(2))
    assert [len(s) for s in sets] == [3, 1]
