"""GPU parity tests of getBackgroundPeaks (K2-K4) and of the reference's registered native
routines (row_sds, row_sds_perm, ProbSampleReplace, bg_sample_helper, euc_dist).

The reference holds no golden values for any of these (parity unpinned), so the
checks are: deterministic front half equal to the oracle's restatement; bin-probability matrix
<= 1e-12; chi-square of sampled destination-bin occupancy against the analytic distribution and
against the oracle's literal alias sampler; within-bin uniformity; and, because the library
documents its Philox counter mapping, exact draw-for-draw reproduction by a numpy Philox.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.stats import chi2

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import philox_ref as ph  # noqa: E402
from oracle import chromvar_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def _mini(g):
    return sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))


def _chi2_ok(obs, exp, alpha=1e-4):
    keep = exp >= 5
    o = np.append(obs[keep], obs[~keep].sum())
    x = np.append(exp[keep], exp[~keep].sum())
    nz = x > 0
    assert o[~nz].sum() == 0, "draws landed where the expected count is zero"
    stat = ((o[nz] - x[nz]) ** 2 / x[nz]).sum()
    return stat < chi2.ppf(1 - alpha, max(1, int(nz.sum()) - 1)), stat


@pytest.mark.parametrize("dataset", ["mini", "example"])
def test_front_half_and_bin_probabilities(engine, golden_mini, golden_example_counts, dataset):
    if dataset == "mini":
        C, bias = _mini(golden_mini), golden_mini["bias"]
    else:
        g = golden_example_counts
        C = sp.csc_matrix((g["counts_x"].astype(np.float64), g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
        C = C[np.nonzero(np.asarray(C.sum(axis=1)).ravel() > 0)[0], :].tocsc()
        bias = np.round(np.clip(np.random.default_rng(5).normal(0.5, 0.1, C.shape[0]), 0.2, 0.9), 3)
    P, N = C.shape
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    bg, dbg = engine.background_peaks(bias, 50, 0.1, 50, debug=True)
    fh = orc.background_front_half(orc.get_fragments_per_peak(C), bias, 50)
    assert np.max(np.abs(dbg["trans_norm_mat"] - fh["trans_norm_mat"])) < 1e-11
    mism = dbg["bin_membership"] != fh["bin_membership"]
    assert mism.sum() == 0, "%d peaks binned differently" % mism.sum()
    assert np.array_equal(dbg["bin_density"], fh["bin_density"])
    Pbin = orc.bin_probability_matrix(orc.bin_weight_matrix(fh["bin_data"], 0.1), fh["bin_density"])
    assert np.max(np.abs(dbg["bin_prob"] - Pbin)) < 1e-12
    assert bg.shape == (P, 50) and bg.min() >= 1 and bg.max() <= P
    if dataset == "mini":
        assert int((dbg["bin_density"] > 0).sum()) == 361 and dbg["bin_density"].max() == 13


def test_sampler_distribution(engine, golden_mini):
    """Destination-bin occupancy per source bin: chi-square against n_i * B * Pbin[i, .] and a
    two-sample comparison against the oracle's literal alias sampler (src/utils.cpp:38-99)."""
    C, bias = _mini(golden_mini), golden_mini["bias"]
    P, N = C.shape
    B = 400
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    engine.set_seed(99)
    bg, dbg = engine.background_peaks(bias, B, 0.1, 50, debug=True)
    bm0 = dbg["bin_membership"] - 1
    Pbin = dbg["bin_prob"]
    dens = dbg["bin_density"]
    src_bins = np.argsort(-dens)[:8]
    fails = 0
    for src in src_bins:
        rows = np.nonzero(bm0 == src)[0]
        dest = bm0[bg[rows].ravel() - 1]
        ok, _ = _chi2_ok(np.bincount(dest, minlength=2500).astype(float), Pbin[src] * dest.size)
        fails += not ok
    assert fails == 0
    # pooled over all peaks: expected occupancy of destination bin j = B * sum_i n_i Pbin[i, j]
    dest_all = np.bincount(bm0[bg.ravel() - 1], minlength=2500).astype(float)
    ok, stat = _chi2_ok(dest_all, B * (dens[:, None] * Pbin).sum(axis=0))
    assert ok, stat
    # within-bin uniformity: inside the fullest bin every peak is drawn equally often
    j = int(np.argmax(dens))
    members = np.nonzero(bm0 == j)[0]
    picks = bg.ravel() - 1
    cnt = np.bincount(picks[bm0[picks] == j], minlength=P)[members].astype(float)
    ok, stat = _chi2_ok(cnt, np.full(members.size, cnt.sum() / members.size))
    assert ok, stat
    # two-sample chi-square against the literal alias sampler of the oracle
    rng = np.random.default_rng(1)
    bin_p = orc.bin_weight_matrix(orc.background_front_half(orc.get_fragments_per_peak(C), bias, 50)["bin_data"], 0.1)
    lit = orc.bg_sample_helper(bm0, bin_p, dens, 100, lambda n: rng.random(n))
    a = np.bincount(bm0[bg[:, :100].ravel() - 1], minlength=2500).astype(float)
    b = np.bincount(bm0[lit.ravel() - 1], minlength=2500).astype(float)
    keep = (a + b) >= 10
    stat = (((a - b) ** 2) / (a + b))[keep].sum()
    assert stat < chi2.ppf(1 - 1e-4, int(keep.sum()) - 1)


def test_sampler_is_a_pure_function_of_seed(engine, golden_mini):
    C, bias = _mini(golden_mini), golden_mini["bias"]
    engine.set_counts_csc(C.shape[0], C.shape[1], C.indptr, C.indices, C.data)
    engine.set_seed(7)
    a = engine.background_peaks(bias, 50, 0.1, 50)
    b = engine.background_peaks(bias, 50, 0.1, 50)
    c = engine.background_peaks(bias, 20, 0.1, 50)
    engine.set_seed(8)
    d = engine.background_peaks(bias, 50, 0.1, 50)
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, :20], c)          # counter = (peak, iteration): independent of niterations
    assert not np.array_equal(a, d)


def test_sampler_reproduced_by_numpy_philox(engine, golden_mini):
    """The documented mapping: draw (p, it) = Philox4x32-10(counter = (p, it, stream 0, 0),
    key = seed); u1 picks the destination bin by CDF search, u2 the peak inside it."""
    C, bias = _mini(golden_mini), golden_mini["bias"]
    P = C.shape[0]
    seed = 0x1234_5678_9ABC
    engine.set_counts_csc(P, C.shape[1], C.indptr, C.indices, C.data)
    engine.set_seed(seed)
    B = 16
    bg, dbg = engine.background_peaks(bias, B, 0.1, 50, debug=True)
    bm0 = dbg["bin_membership"] - 1
    dens = dbg["bin_density"].astype(np.int64)
    start = np.concatenate([[0], np.cumsum(dens)])
    order = np.argsort(bm0, kind="stable")
    k0, k1 = ph.seed_words(seed)
    pp, it = np.meshgrid(np.arange(P), np.arange(B), indexing="ij")
    r = ph.philox4x32_10(pp, it, ph.STREAM_BACKGROUND, 0, k0, k1)
    u1, u2 = ph.u01_53(r[0], r[1]), ph.u01_53(r[2], r[3])
    cdf = np.cumsum(dbg["bin_prob"], axis=1)
    mism = 0
    for p in range(P):
        row = cdf[bm0[p]] / cdf[bm0[p], -1]
        j = np.minimum(np.searchsorted(row, u1[p], side="right"), 2499)
        # a draw within rounding distance of a CDF step may legitimately fall either side
        edge = np.minimum(np.abs(row[j] - u1[p]), np.abs(row[np.maximum(j - 1, 0)] - u1[p])) < 1e-12
        within = np.minimum((u2[p] * dens[j]).astype(np.int64), dens[j] - 1)
        exp = order[start[j] + within] + 1
        mism += int(((exp != bg[p]) & ~edge).sum())
    assert mism == 0


def test_bg_sample_helper_matches_given_matrices(engine):
    """cv_bg_sample_helper with an arbitrary bin_p: chi-square against the induced distribution."""
    rng = np.random.default_rng(3)
    nb, P, B = 40, 3000, 200
    bm0 = rng.integers(0, nb - 5, P).astype(np.int32)          # the last five bins stay empty
    bin_p = rng.random((nb, nb)) ** 3
    dens = np.bincount(bm0, minlength=nb).astype(float)
    out = engine.bg_sample_helper(bm0, bin_p, dens, B)
    assert out.shape == (P, B) and out.min() >= 1 and out.max() <= P
    Pbin = orc.bin_probability_matrix(bin_p, dens)
    for src in (0, 7, 20):
        rows = np.nonzero(bm0 == src)[0]
        dest = bm0[out[rows].ravel() - 1]
        ok, stat = _chi2_ok(np.bincount(dest, minlength=nb).astype(float), Pbin[src] * dest.size)
        assert ok, (src, stat)


def test_prob_sample_replace(engine):
    rng = np.random.default_rng(0)
    prob = rng.random(500) ** 2
    prob[::7] = 0
    seed = 424242
    engine.set_seed(seed)
    size = 200000
    out = engine.prob_sample_replace(size, prob)
    assert out.min() >= 1 and out.max() <= 500
    ok, stat = _chi2_ok(np.bincount(out - 1, minlength=500).astype(float), prob / prob.sum() * size)
    assert ok, stat
    # draw-for-draw with the numpy Philox (stream 1, counter = draw index)
    k0, k1 = ph.seed_words(seed)
    k = np.arange(size)
    r = ph.philox4x32_10(k, 0, ph.STREAM_PROBSAMPLE, 0, k0, k1)
    cdf = np.cumsum(prob)
    u = ph.u01_53(r[0], r[1]) * cdf[-1]
    exp = np.searchsorted(cdf, u, side="right") + 1
    edge = np.abs(cdf[np.minimum(exp - 1, 499)] - u) < 1e-9
    assert ((exp != out) & ~edge).sum() == 0
    # agreement with the oracle's Walker alias sampler (two-sample chi-square)
    ref = orc.prob_sample_replace(size, prob / prob.sum(), lambda n: rng.random(n))
    a = np.bincount(out - 1, minlength=500).astype(float)
    b = np.bincount(ref - 1, minlength=500).astype(float)
    keep = (a + b) >= 10
    assert (((a - b) ** 2) / (a + b))[keep].sum() < chi2.ppf(1 - 1e-4, int(keep.sum()) - 1)


def test_row_sds_and_euc_dist(engine):
    rng = np.random.default_rng(2)
    x = rng.normal(size=(37, 501))
    assert np.allclose(engine.row_sds(x, False), orc.row_sds(x, False), rtol=1e-12)
    x[3, 5] = np.nan
    x[9, :] = np.nan
    x[11, 1:] = np.inf
    got, ref = engine.row_sds(x, True), orc.row_sds(x, True)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.allclose(got[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-12)
    got = engine.row_sds(x, False)
    assert np.isnan(got[[3, 9, 11]]).all() and np.allclose(np.delete(got, [3, 9, 11]),
                                                            np.delete(orc.row_sds(x, False), [3, 9, 11]), rtol=1e-12)
    pts = rng.normal(size=(300, 2))
    assert np.max(np.abs(engine.euc_dist(pts) - orc.euc_dist(pts))) < 1e-14


def test_row_sds_perm_bootstrap(engine):
    """Bootstrap replicates: exact reproduction with the numpy Philox column draws, and the
    sampling distribution of the SD (mean close to the plain SD)."""
    rng = np.random.default_rng(6)
    K, N, R = 9, 257, 64
    z = rng.normal(size=(K, N)) * np.arange(1, K + 1)[:, None]
    z[2, 10] = np.nan
    seed = 77
    engine.set_seed(seed)
    boot = engine.row_sds_perm(z, True, R)
    assert boot.shape == (R, K)
    k0, k1 = ph.seed_words(seed)
    dd, rr = np.meshgrid(np.arange(N), np.arange(R), indexing="ij")
    r = ph.philox4x32_10(dd, rr, ph.STREAM_BOOTSTRAP, 0, k0, k1)
    cols = np.minimum((ph.u01_53(r[0], r[1]) * N).astype(np.int64), N - 1)      # [N draws, R]
    for rep in (0, 1, R - 1):
        ref = orc.row_sds(z[:, cols[:, rep]], True)
        assert np.allclose(boot[rep], ref, rtol=1e-10)
    plain = orc.row_sds(z, True)
    assert np.all(np.abs(boot.mean(axis=0) / plain - 1) < 0.05)
