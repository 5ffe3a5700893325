"""GPU routines against the reference's OWN native code (src/utils.cpp compiled unmodified into
oracle/_ref/libchromvar_ref.so, drawing through R's Mersenne-Twister): the sampler's
destination-bin occupancy passes a two-sample chi-square test against the reference sampler
(BASELINE.json north_star), the deterministic routines agree to rounding.
RNG streams differ by construction (Philox here, Mersenne-Twister there), so sampled outputs are
compared in distribution only.
"""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.stats import chi2, ks_2samp

from oracle import chromvar_oracle as orc
from oracle import ref_native as rn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _ref_lib():
    if not rn.available() and rn.build() is None:
        pytest.fail("oracle/_ref/libchromvar_ref.so did not travel with the repo snapshot")
    rn.load()


def _two_sample_chi2(a, b, alpha=1e-4):
    """Occupancy vectors of two samples of equal size."""
    assert a.sum() == b.sum()
    keep = (a + b) >= 10
    stat = (((a - b) ** 2) / (a + b))[keep].sum()
    return stat < chi2.ppf(1 - alpha, max(1, int(keep.sum()) - 1)), stat


def test_sampler_against_reference_sampler(engine, golden_mini):
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    P, N = C.shape
    B = 400
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    engine.set_seed(31337)
    bg, dbg = engine.background_peaks(g["bias"], B, 0.1, 50, debug=True)
    bm0 = dbg["bin_membership"] - 1
    dens = dbg["bin_density"].astype(np.float64)
    # the reference sampler, fed the reference-side weights of the same bins
    fh = orc.background_front_half(orc.get_fragments_per_peak(C), g["bias"], 50)
    assert np.array_equal(fh["bin_membership"] - 1, bm0)
    bin_p = orc.bin_weight_matrix(fh["bin_data"], 0.1)
    rn.set_seed(2017)
    ref = rn.bg_sample_helper(bm0, bin_p, dens, B)
    assert ref.shape == bg.shape
    # pooled destination-bin occupancy
    ok, stat = _two_sample_chi2(np.bincount(bm0[bg.ravel() - 1], minlength=2500).astype(float),
                                np.bincount(bm0[ref.ravel() - 1], minlength=2500).astype(float))
    assert ok, stat
    # per source bin, the eight fullest
    for src in np.argsort(-dens)[:8]:
        rows = np.nonzero(bm0 == src)[0]
        ok, stat = _two_sample_chi2(np.bincount(bm0[bg[rows].ravel() - 1], minlength=2500).astype(float),
                                    np.bincount(bm0[ref[rows].ravel() - 1], minlength=2500).astype(float))
        assert ok, (src, stat)
    # per-peak pick frequencies (which peak inside the destination bin)
    ok, stat = _two_sample_chi2(np.bincount(bg.ravel() - 1, minlength=P).astype(float),
                                np.bincount(ref.ravel() - 1, minlength=P).astype(float))
    assert ok, stat


def test_bg_sample_helper_against_reference(engine):
    rng = np.random.default_rng(3)
    nb, P, B = 40, 3000, 200
    bm0 = rng.integers(0, nb - 5, P).astype(np.int32)
    bin_p = rng.random((nb, nb)) ** 3
    dens = np.bincount(bm0, minlength=nb).astype(float)
    engine.set_seed(5)
    got = engine.bg_sample_helper(bm0, bin_p, dens, B)
    rn.set_seed(5)
    ref = rn.bg_sample_helper(bm0, bin_p, dens, B)
    for src in (0, 7, 20):
        rows = np.nonzero(bm0 == src)[0]
        ok, stat = _two_sample_chi2(np.bincount(bm0[got[rows].ravel() - 1], minlength=nb).astype(float),
                                    np.bincount(bm0[ref[rows].ravel() - 1], minlength=nb).astype(float))
        assert ok, (src, stat)
    ok, stat = _two_sample_chi2(np.bincount(got.ravel() - 1, minlength=P).astype(float),
                                np.bincount(ref.ravel() - 1, minlength=P).astype(float))
    assert ok, stat


def test_prob_sample_replace_against_reference(engine):
    rng = np.random.default_rng(0)
    prob = rng.random(500) ** 2
    prob[::7] = 0
    prob /= prob.sum()
    size = 200000
    engine.set_seed(9)
    got = engine.prob_sample_replace(size, prob)
    rn.set_seed(9)
    ref = rn.prob_sample_replace(size, prob)
    ok, stat = _two_sample_chi2(np.bincount(got - 1, minlength=500).astype(float),
                                np.bincount(ref - 1, minlength=500).astype(float))
    assert ok, stat
    assert np.all(prob[got - 1] > 0)


def test_row_sds_and_euc_dist_against_reference(engine):
    rng = np.random.default_rng(2)
    x = rng.normal(size=(37, 501)) * rng.lognormal(size=(37, 1)) + 3
    assert np.allclose(engine.row_sds(x, False), rn.row_sds(x, False), rtol=1e-12, atol=0)
    x[3, 5] = np.nan
    x[9, :] = np.nan
    x[11, 1:] = np.inf
    for na_rm in (True, False):
        got, ref = engine.row_sds(x, na_rm), rn.row_sds(x, na_rm)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), na_rm
        m = ~np.isnan(ref)
        assert np.allclose(got[m], ref[m], rtol=1e-12, atol=0)
    pts = rng.normal(size=(300, 2))
    assert np.max(np.abs(engine.euc_dist(pts) - rn.euc_dist(pts))) < 1e-14


def test_row_sds_perm_against_reference(engine):
    """Bootstrap replicates of the row SD: same sampling distribution as the reference's
    row_sds_perm (two-sample Kolmogorov-Smirnov per row)."""
    rng = np.random.default_rng(6)
    K, N, R = 5, 200, 1000
    z = rng.standard_t(5, size=(K, N)) * np.arange(1, K + 1)[:, None]
    z[2, 10] = np.nan
    engine.set_seed(77)
    got = engine.row_sds_perm(z, True, R)                  # R x K
    rn.set_seed(77)
    ref = np.stack([rn.row_sds_perm(z, True) for _ in range(R)])
    assert rn.draws() == R * N
    for k in range(K):
        assert ks_2samp(got[:, k], ref[:, k]).pvalue > 1e-4, k
