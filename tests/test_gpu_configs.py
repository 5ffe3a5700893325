"""GPU parity on the shapes of BASELINE.json's configs 3, 4 and 5, through the C ABI.

Each test runs the FULL annotation set of the configuration on the GPU (so the kernel sees the
operand shape of the real configuration: 2080 k-mer rows / 500k peaks / u8 digit planes) and lets the
CPU oracle restate a RANDOM sample of the annotations on the same inputs:

  * integer observed and background sums bit-equal (R/compute_deviations.R:486, :501),
  * deviations and z within 1e-5 relative (abs(d) <= 1e-5 * max(abs(ref), 1e-3)), NA masks equal,
  * fractionMatches exact, fractionBackgroundOverlap 1e-12.

Cells are a slice of the configuration where the full matrix would take the oracle minutes (configs 3
and 5: every cell is independent of the others once the expectation is fixed, R/compute_deviations.R:
486-504); config 4 runs at its full 500k x 200 size.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.stats import chi2

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402
from test_gpu_parity import assert_close  # noqa: E402

pytestmark = pytest.mark.gpu


def _defaults(engine):
    for key, v in (("path", 0), ("dense_fp4", -1), ("dense_epilogue", 0), ("cell_chunk", 0), ("check_zero_peaks", 1)):
        engine.set_option(key, v)


def _sample_check(engine, C, d, bg, e, n_sample, seed, expect_bits=None, expect_planes=None):
    """GPU: all K annotations with the integer sums kept; oracle: n_sample random annotations."""
    P, N = C.shape
    ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
    K = len(ptr) - 1
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    got = engine.deviations(ptr, idx, bg, e, index_base=1, sums=True)
    st = engine.stats()
    assert st["path"] == 1, "the dense tcgen05 path was expected for this shape"
    if expect_bits is not None:
        assert st["operand_bits"] == expect_bits, st
    if expect_planes is not None:
        assert st["planes"] == expect_planes, st
    ks = np.sort(np.random.default_rng(seed).choice(K, n_sample, replace=False))
    Cr = sp.csr_matrix(C, dtype=np.float64)
    for k in ks:
        ref = orc.compute_deviations_single(idx[ptr[k]:ptr[k + 1]], Cr, bg, e, intermediate_results=True)
        assert np.array_equal(got["observed"][k], ref["observed"].astype(np.int64)), "observed sums, annotation %d" % k
        assert np.array_equal(got["sampled"][k], ref["sampled"].astype(np.int64)), "background sums, annotation %d" % k
        assert_close(got["dev"][k], ref["dev"], "deviations[%d]" % k)
        assert_close(got["z"][k], ref["z"], "z[%d]" % k)
        assert got["matches"][k] == ref["matches"]
        assert abs(got["overlap"][k] - ref["overlap"]) <= 1e-12
    # the row SD of z over this handle's cells (fused computeVariability) on the sampled rows
    assert_close(got["variability"][ks], orc.row_sds(got["z"][ks], True), "variability")
    return got, st


def test_config3_kmers_cell_slice(engine):
    """Config 3: 2080 RC-collapsed 6-mer annotations (density ~0.21) x 200 000 peaks, 2048 of the
    50 000 cells.  Dense k-mer sets push most multiplicities outside the e2m1 alphabet."""
    from chromvar_b200 import synthetic
    _defaults(engine)
    d = synthetic.make_config("cfg3_kmers", N=2048)
    C = d["counts"]
    P = C.shape[0]
    assert P == 200_000 and len(d["annoptr"]) - 1 == 2080
    e = orc.compute_expectations_core(C)
    engine.set_counts_csc(P, C.shape[1], C.indptr, C.indices, C.data)
    engine.set_seed(20174)
    bg = engine.background_peaks(d["bias"], 50, 0.1, 50)
    _sample_check(engine, C, d, bg, e, n_sample=6, seed=31)


def test_config4_bulk_full_size_two_digit_planes(engine):
    """Config 4 at full size: dense bulk counts 500 000 peaks x 200 samples (values to ~2e4: two u8
    digit planes, int64 sums), 870 annotations, background from the GPU sampler at P = 5e5."""
    from chromvar_b200 import synthetic
    _defaults(engine)
    d = synthetic.make_config("cfg4_bulk")
    X = d["counts"]
    P, N = X.shape
    assert (P, N) == (500_000, 200) and X.max() > 255
    engine.set_counts_dense(X)
    fpp, fps, tot = engine.fragments()
    assert np.array_equal(fpp, X.sum(axis=1)) and np.array_equal(fps, X.sum(axis=0)) and tot == X.sum()
    e = engine.expectations()
    assert np.max(np.abs(e - fpp / tot)) <= 1e-18
    engine.set_seed(20175)
    bg = engine.background_peaks(d["bias"], 50, 0.1, 50)
    C = sp.csc_matrix(X)
    ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
    got = engine.deviations(ptr, idx, bg, e, index_base=1, sums=True)
    st = engine.stats()
    assert st["path"] == 1 and st["planes"] == 2, st
    Cr = sp.csr_matrix(C, dtype=np.float64)
    for k in np.sort(np.random.default_rng(41).choice(870, 5, replace=False)):
        ref = orc.compute_deviations_single(idx[ptr[k]:ptr[k + 1]], Cr, bg, e, intermediate_results=True)
        assert np.array_equal(got["observed"][k], ref["observed"].astype(np.int64))
        assert np.array_equal(got["sampled"][k], ref["sampled"].astype(np.int64))
        assert_close(got["dev"][k], ref["dev"], "deviations[%d]" % k)
        assert_close(got["z"][k], ref["z"], "z[%d]" % k)
        assert got["matches"][k] == ref["matches"] and abs(got["overlap"][k] - ref["overlap"]) <= 1e-12


def test_config4_sampler_at_500k_peaks(engine):
    """getBackgroundPeaks at P = 5e5 (BASELINE config 4): deterministic front half against the
    oracle, bin-probability matrix <= 1e-12, chi-square of the destination-bin occupancy."""
    from chromvar_b200 import synthetic
    rng = np.random.default_rng(20175)
    P, N = 500_000, 64
    bias = synthetic.make_bias(P, rng)
    # per-peak totals with the spread of bulk data; a handful of cells is enough for the sampler
    a = rng.lognormal(0.0, 1.0, P)
    X = np.asfortranarray(rng.poisson(np.outer(a, rng.uniform(20, 60, N))).astype(np.float64))
    X[X.sum(axis=1) == 0, 0] = 1
    engine.set_counts_dense(X)
    engine.set_seed(5)
    B = 50
    bg, dbg = engine.background_peaks(bias, B, 0.1, 50, debug=True)
    fh = orc.background_front_half(X.sum(axis=1), bias, 50)
    assert np.max(np.abs(dbg["trans_norm_mat"] - fh["trans_norm_mat"])) < 1e-10
    # a peak whose whitened coordinate sits within rounding of a bin boundary may legitimately differ
    mism = np.nonzero(dbg["bin_membership"] != fh["bin_membership"])[0]
    assert mism.size <= 2, "%d of %d peaks binned differently" % (mism.size, P)
    if mism.size == 0:
        assert np.array_equal(dbg["bin_density"], fh["bin_density"])
        Pbin = orc.bin_probability_matrix(orc.bin_weight_matrix(fh["bin_data"], 0.1), fh["bin_density"])
        assert np.max(np.abs(dbg["bin_prob"] - Pbin)) < 1e-12
    assert bg.shape == (P, B) and bg.min() >= 1 and bg.max() <= P
    bm0 = dbg["bin_membership"] - 1
    dens, Pb = dbg["bin_density"], dbg["bin_prob"]
    dest_all = np.bincount(bm0[bg.ravel() - 1], minlength=2500).astype(float)
    exp = B * (dens[:, None] * Pb).sum(axis=0)
    keep = exp >= 5
    o = np.append(dest_all[keep], dest_all[~keep].sum())
    x = np.append(exp[keep], exp[~keep].sum())
    nz = x > 0
    assert o[~nz].sum() == 0
    stat = ((o[nz] - x[nz]) ** 2 / x[nz]).sum()
    assert stat < chi2.ppf(1 - 1e-4, int(nz.sum()) - 1), stat
    # per source bin (the three fullest)
    for src in np.argsort(-dens)[:3]:
        rows = np.nonzero(bm0 == src)[0]
        dest = np.bincount(bm0[bg[rows].ravel() - 1], minlength=2500).astype(float)
        ex = Pb[src] * dest.sum()
        kp = ex >= 5
        o2 = np.append(dest[kp], dest[~kp].sum())
        x2 = np.append(ex[kp], ex[~kp].sum())
        ok = x2 > 0
        assert o2[~ok].sum() == 0
        assert ((o2[ok] - x2[ok]) ** 2 / x2[ok]).sum() < chi2.ppf(1 - 1e-4, max(1, int(ok.sum()) - 1))


def test_config5_atlas_shard_cell_slice(engine):
    """Config 5's shard shape: 500 000 peaks x 870 annotations at 1 % density, 4096 of the shard's
    125 000 cells, run like a shard of a larger matrix (check_zero_peaks = 0, expectation from global
    totals that this slice alone does not reproduce)."""
    from chromvar_b200 import synthetic
    _defaults(engine)
    d = synthetic.make_config("cfg5_atlas_shard", N=4096)
    C = d["counts"]
    P = C.shape[0]
    assert P == 500_000
    rng = np.random.default_rng(77)
    # "global" per-peak totals: this slice plus the (synthetic) rest of the atlas
    glob = np.asarray(C.sum(axis=1)).ravel() + rng.poisson(40.0, P)
    e = glob / glob.sum()
    engine.set_option("check_zero_peaks", 0)
    try:
        engine.set_counts_csc(P, C.shape[1], C.indptr, C.indices, C.data)
        engine.set_seed(20176)
        bg = engine.background_peaks(d["bias"], 50, 0.1, 50)
        _sample_check(engine, C, d, bg, e, n_sample=6, seed=51, expect_bits=4)
    finally:
        engine.set_option("check_zero_peaks", 1)
