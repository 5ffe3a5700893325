"""Adversarial inputs for the fixed-point form of the fused epilogue (cv_dense_dev.cuh, DN_EPI_FUSED32):
the per-iteration statistics run without fp64 -- the mean over iterations as an exact 64-bit integer
sum, the variance as an fp32 sum of squared exact integer differences to the exact integer mean
(second sweep over the accumulator in tensor memory).  Inputs built to make a one-pass, pivoted
variance cancel:

  * the first background iteration is an outlier (one large sum against 49 small ones),
  * every background iteration but one is zero (rare annotation, sparse cells),
  * sums just below 2^24 (the largest the f32 accumulators of the e2m1 kernel hold exactly),
  * B in {2, 3, 207}.

Bars: integer sums bit-equal to the oracle, deviations / z within 1e-5 relative, and
-- the margin the two-sweep form is meant to give -- within 2e-6 on these cases.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402
from test_gpu_parity import assert_close, sets_to_csr  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture
def fx_engine(engine):
    engine.set_option("path", 2)
    engine.set_option("dense_epilogue", 1)
    yield engine
    for key, v in (("path", 0), ("dense_fp4", -1), ("dense_epilogue", 0), ("cell_chunk", 0)):
        engine.set_option(key, v)


def _max_rel(got, ref):
    ok = np.isfinite(ref) & np.isfinite(got)
    return float((np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-3)).max()) if ok.any() else 0.0


def _run(engine, C, sets, bg, e, fp4, tight=2e-6):
    C = sp.csc_matrix(C)
    P, N = C.shape
    ptr, idx = sets_to_csr(sets)
    engine.set_option("dense_fp4", 1 if fp4 else 0)
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    got = engine.deviations(ptr, idx, bg, e, index_base=1, sums=True)
    st = engine.stats()
    assert st["path"] == 1 and st["epilogue"] == 32, st
    assert st["operand_bits"] == (4 if fp4 else 8), st
    ref = orc.compute_deviations_core(C, [np.asarray(s) for s in sets], bg, e, intermediate=True)
    assert np.array_equal(got["observed"], ref["observed"].astype(np.int64))
    assert np.array_equal(got["sampled"], ref["sampled"].astype(np.int64))
    assert_close(got["dev"], ref["dev"], "deviations")
    assert_close(got["z"], ref["z"], "z")
    worst = max(_max_rel(got["dev"], ref["dev"]), _max_rel(got["z"], ref["z"]))
    assert worst <= tight, "fixed-point epilogue: max rel err %.3g" % worst
    # and the fp64 epilogue on the same sums
    engine.set_option("dense_epilogue", 2)
    g64 = engine.deviations(ptr, idx, bg, e, index_base=1, sums=False)
    assert engine.stats()["epilogue"] == 64
    engine.set_option("dense_epilogue", 1)
    assert_close(g64["z"], ref["z"], "z (fp64 epilogue)")
    return got, ref, worst


def _sparse_counts(P, N, density, rng, maxv=3):
    C = sp.random(P, N, density=density, format="csc", random_state=np.random.RandomState(int(rng.integers(1 << 30))),
                  data_rvs=lambda n: rng.integers(1, maxv + 1, n).astype(np.float64))
    C = C.tolil()
    empty = np.nonzero(np.asarray(C.sum(axis=1)).ravel() == 0)[0]
    for p in empty:
        C[p, int(rng.integers(N))] = 1.0
    return C.tocsc()


@pytest.mark.parametrize("fp4", [True, False])
def test_outlier_first_background_iteration(fx_engine, fp4):
    """bg[:, 0] sends every peak of the sets to a handful of heavy peaks, all other iterations to
    light ones: the first iteration's deviation is an outlier by orders of magnitude."""
    rng = np.random.default_rng(11)
    P, N, B = 6000, 512, 50
    C = _sparse_counts(P, N, 0.02, rng).tolil()
    heavy = np.arange(600)
    for p in heavy:                                      # heavy peaks: a fragment (or six) in most cells
        cols = rng.choice(N, int(0.9 * N), replace=False)
        C[p, cols] = rng.choice([1.0, 2.0, 3.0, 4.0, 6.0], cols.size)
    C = C.tocsc()
    e = orc.compute_expectations_core(C)
    bg = np.asfortranarray(rng.integers(601, P + 1, size=(P, B)).astype(np.int32))
    bg[:, 0] = rng.choice(heavy + 1, P)                  # the outlier iteration
    sets = [rng.choice(np.arange(601, P + 1), n, replace=False) for n in (30, 75, 200, 600, 1500, 12, 7)]
    _run(fx_engine, C, sets, bg, e, fp4)


@pytest.mark.parametrize("fp4", [True, False])
def test_one_nonzero_background_iteration(fx_engine, fp4):
    """Rare annotations on sparse cells: for most (annotation, cell) pairs a single background
    iteration holds a fragment and 49 hold none (the case the round-1 pivot amplified ~50x)."""
    rng = np.random.default_rng(12)
    P, N, B = 8000, 768, 50
    C = _sparse_counts(P, N, 0.004, rng, maxv=2)
    e = orc.compute_expectations_core(C)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, B)).astype(np.int32))
    sets = [rng.choice(np.arange(1, P + 1), n, replace=False) for n in (3, 5, 8, 13, 21, 34, 55, 4, 6)]
    _run(fx_engine, C, sets, bg, e, fp4)


@pytest.mark.parametrize("fp4", [True, False])
def test_sums_just_below_2_24(fx_engine, fp4):
    """u8 operands: every cell total is a little under 2^24 (the bound up to which the fixed-point
    form is allowed to run) and the first annotation is 97 % of the peaks -- observed sums of 1.6e7.  e2m1
    operands: the largest sums dense counts inside the alphabet reach at this size (3.5e5; the f32
    accumulators' exactness up to 2^24 is what tools/fp4_probe.cu measures).  Background columns
    are permutations (multiplicity bound 1), so the exactness check lets the fixed-point form run."""
    rng = np.random.default_rng(13)
    P, N, B = 66_000, 256, 50
    vals = np.array([252.0, 253.0, 254.0]) if not fp4 else np.array([6.0, 6.0, 4.0])
    X = rng.choice(vals, size=(P, N))
    tot = X.sum(axis=0)
    assert tot.max() < 2 ** 24
    C = sp.csc_matrix(X)
    e = orc.compute_expectations_core(C)
    bg = np.empty((P, B), np.int32, order="F")
    for j in range(B):
        bg[:, j] = rng.permutation(P) + 1
    # (not ALL peaks: with permutation columns every iteration would then give the same sum and the
    # reference's sd is rounding noise around zero)
    sets = [rng.choice(np.arange(1, P + 1), int(0.97 * P), replace=False), rng.choice(np.arange(1, P + 1), P // 2, replace=False),
            rng.choice(np.arange(1, P + 1), 5000, replace=False), np.arange(1, 40)]
    got, ref, _ = _run(fx_engine, C, sets, bg, e, fp4)
    if not fp4:
        assert got["observed"].max() > 0.93 * 2 ** 24


@pytest.mark.parametrize("B", [2, 3, 207])
def test_fixed_point_niterations(fx_engine, golden_mini, B):
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    A = sp.csc_matrix((g["ix_x"].astype(bool), g["ix_i"], g["ix_p"]), shape=tuple(g["ix_dim"]))
    sets = orc.convert_to_ix_list(A)
    rng = np.random.default_rng(B)
    sets = list(sets) + [rng.choice(np.arange(1, 1001), n, replace=False) for n in (2, 9, 40, 160)]
    e = orc.compute_expectations_core(C)
    bg = orc.get_background_peaks_core(C, g["bias"], niterations=B, rng=np.random.default_rng(3))
    for fp4 in (True, False):
        _run(fx_engine, C, sets, bg, e, fp4)
