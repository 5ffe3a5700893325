"""Seeded random cases through every contraction path against the oracle: ragged shapes (cells not a
multiple of any tile, a single annotation, B in {1, 2, 3, 7, 50}), value ranges that select the e2m1 /
u8 one-plane / two-plane / three-plane operands, every accepted value type, columns given unsorted,
lists that repeat a peak, empty lists, background columns that point at one peak many times.

Integer observed and background sums are compared bit for bit, deviations and z with the 1e-5 bar
(tests/test_gpu_parity.py) -- except where the REFERENCE's own value is rounding noise, which tiny
inputs produce easily (2 cells with one dominating, 37 peaks whose background sets collide with the
set): all B background deviations agree to 7 digits (their sd is the rounding of the last bits, z =
dev / noise), the deviation cancels to 1e-7 of its terms (z = noise / sd), or the expected count equals
the NA threshold 1 to the last bits (the order of the sum of expectations decides; the reference's is
its sparse-matrix library's).  Those entries are recognised from the oracle's intermediates and not
compared (z, and for the threshold case the deviation); everything else is.  The cases are
fixed (seeds below); CV_FUZZ_CASES=n adds n more from fresh seeds, printed on failure."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import chromvar_oracle as orc  # noqa: E402
from test_gpu_parity import assert_close, sets_to_csr  # noqa: E402

pytestmark = pytest.mark.gpu

SEEDS = list(range(100, 124))
if os.environ.get("CV_FUZZ_CASES"):
    SEEDS += [int(s) for s in np.random.SeedSequence().generate_state(int(os.environ["CV_FUZZ_CASES"]))]


def _case(seed):
    rng = np.random.default_rng(seed)
    P = int(rng.choice([37, 256, 300, 1025, 2500, 4097]))
    N = int(rng.choice([2, 3, 31, 130, 256, 257, 600]))
    K = int(rng.choice([1, 2, 9, 33, 70]))
    B = int(rng.choice([1, 2, 3, 7, 50]))
    top = int(rng.choice([3, 6, 40, 255, 300, 70000]))              # largest count: operand type
    density = float(rng.choice([0.02, 0.15, 0.6]))
    X = (rng.random((P, N)) < density) * rng.integers(1, min(top, 6) + 1, (P, N))
    if top > 6:
        hot = rng.random((P, N)) < 0.01
        X = np.where(hot, rng.integers(1, top + 1, (P, N)), X)
    X = X.astype(np.float64)
    X[X.sum(axis=1) == 0, int(rng.integers(N))] = 1                  # every peak has a fragment
    sets = []
    repeats = rng.random() < 0.25                                    # (a repeat takes the whole call off the bit-column builders)
    for k in range(K):
        n = int(rng.integers(1, max(2, min(P, 400))))
        s = rng.choice(np.arange(1, P + 1), n, replace=False)
        if repeats and rng.random() < 0.3 and n > 2:
            s[1] = s[0]                                              # a list that repeats a peak
        sets.append(s)
    if K > 2 and rng.random() < 0.5:
        sets[1] = np.zeros(0, np.int64)                              # an empty list: all-NA row
    bg = rng.integers(1, P + 1, size=(P, B)).astype(np.int32)
    if rng.random() < 0.3:                                           # one peak drawn many times
        bg[rng.random((P, B)) < float(rng.choice([0.01, 0.3]))] = int(rng.integers(1, P + 1))
    dtype = rng.choice(["f64", "f32", "i32", "u8"])
    if dtype == "u8" and top > 255:
        dtype = "i32"
    shuffle = rng.random() < 0.4
    return dict(P=P, N=N, K=K, B=B, X=X, sets=sets, bg=np.asfortranarray(bg), dtype=str(dtype), shuffle=shuffle, top=top)


def _csc_arrays(X, dtype, shuffle, rng):
    C = sp.csc_matrix(X)
    C.sort_indices()
    indptr, indices, data = C.indptr.copy(), C.indices.copy(), C.data.copy()
    if shuffle:                                                      # rows of a column in any order
        for c in range(X.shape[1]):
            a, b = indptr[c], indptr[c + 1]
            perm = rng.permutation(b - a)
            indices[a:b], data[a:b] = indices[a:b][perm], data[a:b][perm]
    data = data.astype({"f64": np.float64, "f32": np.float32, "i32": np.int32, "u8": np.uint8}[dtype])
    return C, indptr, indices, data


def _ill_conditioned(C, sets, bg, e, ref):
    """([K, N] mask of entries whose reference z is rounding noise, mask of entries whose expected count
    sits on the `expected < 1` NA threshold to the last bits) -- see the module docstring."""
    f = np.asarray(C.sum(axis=0)).ravel()
    K, B, N = ref["sampled"].shape
    ill = np.zeros((K, N), bool)
    edge = np.zeros((K, N), bool)
    with np.errstate(all="ignore"):
        for k, s in enumerate(sets):
            m = np.asarray(s) - 1
            ex = np.array([e[m].sum()] + [e[bg[m, j] - 1].sum() for j in range(B)])[:, None] * f[None, :]
            sums = np.concatenate([ref["observed"][k][None, :], ref["sampled"][k]], axis=0)
            d = (sums - ex) / ex
            scale = np.nanmax(np.abs(d), axis=0)
            floor = np.maximum(1e-7 * scale, 1e-9 * (1.0 + scale))       # d = sums / ex - 1 carries ~1e-16 (1 + |d|) of rounding
            sd = np.std(d[1:], axis=0, ddof=1) if B > 1 else np.full(N, np.nan)
            dev = d[0] - d[1:].mean(axis=0)
            ill[k] = (sd <= floor) | (np.abs(dev) <= floor)
            edge[k] = np.abs(ex[0] - 1.0) <= 1e-9                         # fail_filter: expected < 1 (R/compute_deviations.R:509)
    return ill | edge, edge


def _compare(got, ref, masks, sums=True):
    ill, edge = masks
    if sums:
        assert np.array_equal(got["observed"], ref["observed"].astype(np.int64)), "observed sums differ"
        assert np.array_equal(got["sampled"], ref["sampled"].astype(np.int64)), "background sums differ"
        assert np.allclose(got["matches"], ref["matches"], rtol=1e-15, atol=0)
        assert_close(got["overlap"], ref["overlap"], "overlap")
    assert_close(np.where(edge, 0.0, got["dev"]), np.where(edge, 0.0, ref["dev"]), "deviations")
    assert_close(np.where(ill, 0.0, got["z"]), np.where(ill, 0.0, ref["z"]), "z")


SEEN = set()        # (operand bits, digit planes, few-cells kernel, epilogue) of every dense run, path of the others


VARIANTS = (("auto", dict()),
            ("row gather", dict(path=1)),
            ("dense u8 fp64 epilogue", dict(path=2, dense_fp4=0, dense_epilogue=2)),
            ("dense fixed-point epilogue", dict(path=2, dense_epilogue=1)),
            ("dense, stacked operand only", dict(path=2, small_n_gather=0)),
            ("dense, few-cells kernel", dict(path=2, small_n_gather=1)))
DEFAULTS = dict(path=0, dense_fp4=-1, dense_epilogue=0, small_n_gather=-1, dense_cta_pair=-1)


@pytest.mark.parametrize("seed", SEEDS)
def test_random_case_on_every_path(engine, seed):
    from chromvar_b200 import _lib
    c = _case(seed)
    rng = np.random.default_rng(seed + 1)
    C, indptr, indices, data = _csc_arrays(c["X"], c["dtype"], c["shuffle"], rng)
    P, N = c["P"], c["N"]
    e = orc.compute_expectations_core(C)
    live = [k for k in range(c["K"]) if len(c["sets"][k])]
    ref = orc.compute_deviations_core(C, [c["sets"][k] for k in live], c["bg"], e, intermediate=True)
    ill = _ill_conditioned(C, [c["sets"][k] for k in live], c["bg"], e, ref)
    ptr, idx = sets_to_csr(c["sets"])
    what = "seed %d: P %d N %d K %d B %d top %d %s%s" % (seed, P, N, c["K"], c["B"], c["top"], c["dtype"],
                                                         " unsorted" if c["shuffle"] else "")
    try:
        engine.set_counts_csc(P, N, indptr, indices, data)
        assert np.array_equal(engine.expectations(), e), what
        for name, opts in VARIANTS:
            for k, v in {**DEFAULTS, **opts}.items():
                engine.set_option(k, v)
            try:
                got = engine.deviations(ptr, idx, c["bg"], e, index_base=1, sums=True)
            except _lib.ChromVARError as err:
                # a FORCED dense operand may decline multiplicities beyond a byte; the automatic choice may not
                mult = max(int(np.bincount(c["bg"][:, j]).max()) for j in range(c["B"]))
                assert name != "auto" and err.code == _lib.CV_ERR_UNSUPPORTED and "multiplicity" in str(err) and mult > 127, \
                    "%s, %s: %s" % (what, name, err)
                continue
            st = engine.stats()
            SEEN.add(("dense", st["operand_bits"], st["planes"], st["few_cells"], st["epilogue"]) if st["path"] == 1
                     else ("row gather",))
            sub = {key: (v[live] if isinstance(v, np.ndarray) and v.shape[0] == c["K"] else v) for key, v in got.items()}
            try:
                _compare(sub, ref, ill)
            except AssertionError as err:
                raise AssertionError("%s, %s (path %d, bits %d, planes %d, few_cells %d): %s" % (
                    what, name, engine.stats()["path"], engine.stats()["operand_bits"], engine.stats()["planes"],
                    engine.stats()["few_cells"], err)) from err
            dead = [k for k in range(c["K"]) if not len(c["sets"][k])]
            for k in dead:
                assert np.all(np.isnan(got["z"][k])) and np.all(np.isnan(got["dev"][k])), what
        # the one-call form on the same host arrays (streams the counts in, results out)
        for k, v in DEFAULTS.items():
            engine.set_option(k, v)
        one = engine.deviations_csc(P, N, indptr, indices, data, ptr, idx, c["bg"], e, index_base=1)
        sub = {key: (v[live] if isinstance(v, np.ndarray) and v.shape[0] == c["K"] else v) for key, v in one.items()}
        _compare(sub, ref, ill, sums=False)
    finally:
        for k, v in DEFAULTS.items():
            engine.set_option(k, v)


def test_the_fixed_seeds_reach_every_kernel_family():
    """Runs after the cases above (file order): the fixed seeds must have exercised the row-gather kernel,
    the e2m1 and u8 GEMMs with one, two and three digit planes, both epilogues and the few-cells kernel."""
    if len(SEEN) == 0:
        pytest.skip("cases did not run in this process")
    dense = [s for s in SEEN if s[0] == "dense"]
    assert ("row gather",) in SEEN
    assert any(s[1] == 4 for s in dense), SEEN                       # e2m1 operands
    assert {1, 2, 3} <= {s[2] for s in dense if s[1] == 8}, SEEN       # u8, digit planes
    assert any(s[3] == 1 for s in dense), SEEN                       # few-cells kernel
    assert len({s[4] for s in dense}) >= 2, SEEN                     # fixed-point and fp64 epilogues


# ---- medium shapes: several tile waves, cell chunks, overflow columns, kernel options ----------------------------
MEDIUM_SEEDS = list(range(500, 508))
if os.environ.get("CV_FUZZ_MEDIUM"):
    MEDIUM_SEEDS += [int(s) for s in np.random.SeedSequence().generate_state(int(os.environ["CV_FUZZ_MEDIUM"]))]


def _medium_case(seed):
    rng = np.random.default_rng(seed)
    P = int(rng.choice([6000, 20011, 33000]))
    N = int(rng.choice([700, 1500, 2600]))
    K = int(rng.choice([40, 150, 333]))
    B = int(rng.choice([10, 50, 50, 60]))
    values = [np.array([1, 1, 1, 2, 3]), np.array([1, 2, 3, 4, 6, 5, 7, 9, 13]), np.array([1, 2, 200, 255]),
              np.array([1, 3, 300, 1000])][int(rng.integers(4))]
    C = sp.random(P, N, density=float(rng.choice([0.01, 0.04])), format="csc", random_state=int(seed),
                  data_rvs=lambda n: rng.choice(values, n).astype(np.float64))
    empty = np.nonzero(np.asarray(C.sum(axis=1)).ravel() == 0)[0]          # every peak has a fragment
    C = (C + sp.csc_matrix((np.ones(empty.size), (empty, rng.integers(0, N, empty.size))), shape=(P, N))).tocsc()
    C.sort_indices()
    sets = [rng.choice(np.arange(1, P + 1), int(rng.integers(1, 1500)), replace=False) for _ in range(K)]
    sets[K // 2] = np.zeros(0, np.int64)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, B)).astype(np.int32))
    opts = dict(cell_chunk=int(rng.choice([0, 256, 768])), dense_cta_pair=int(rng.choice([-1, 0])),
                dense_lockstep=int(rng.choice([0, 1, 2])), dense_band=int(rng.choice([0, 1, 5])),
                dense_epilogue=int(rng.choice([0, 1, 2])), dense_fp4=int(rng.choice([-1, 0])))
    return C, sets, bg, opts


MEDIUM_DEFAULTS = dict(path=0, cell_chunk=0, dense_cta_pair=-1, dense_lockstep=1, dense_band=0, dense_epilogue=0,
                       dense_fp4=-1)


@pytest.mark.parametrize("seed", MEDIUM_SEEDS)
def test_random_medium_case_with_random_kernel_options(engine, seed):
    C, sets, bg, opts = _medium_case(seed)
    P, N = C.shape
    K = len(sets)
    e = orc.compute_expectations_core(C)
    ptr, idx = sets_to_csr(sets)
    rng = np.random.default_rng(seed + 7)
    check = sorted(rng.choice([k for k in range(K) if len(sets[k])], 6, replace=False).tolist())
    ref = orc.compute_deviations_core(C, [sets[k] for k in check], bg, e, intermediate=True)
    masks = _ill_conditioned(C, [sets[k] for k in check], bg, e, ref)
    what = "seed %d: P %d N %d K %d B %d %s" % (seed, P, N, K, bg.shape[1], opts)
    try:
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
        for path in (2, 1):
            engine.set_option("path", path)
            for k, v in opts.items():
                engine.set_option(k, v)
            got = engine.deviations(ptr, idx, bg, e, index_base=1, sums=True)
            sub = {key: (v[check] if isinstance(v, np.ndarray) and v.shape[0] == K else v) for key, v in got.items()}
            try:
                _compare(sub, ref, masks)
            except AssertionError as err:
                raise AssertionError("%s, path %d: %s" % (what, path, err)) from err
            assert np.all(np.isnan(got["z"][K // 2]))
            assert_close(got["variability"], orc.row_sds(got["z"], True), "variability")
        # streamed one-call form with the same options
        engine.set_option("path", 0)
        one = engine.deviations_csc(P, N, C.indptr, C.indices, C.data, ptr, idx, bg, e, index_base=1)
        sub = {key: (v[check] if isinstance(v, np.ndarray) and v.shape[0] == K else v) for key, v in one.items()}
        _compare(sub, ref, masks, sums=False)
    finally:
        for k, v in MEDIUM_DEFAULTS.items():
            engine.set_option(k, v)


# ---- shards and column blocks ---------------------------------------------------------------------------------
SHARD_SEEDS = list(range(700, 708))
if os.environ.get("CV_FUZZ_SHARDS"):
    SHARD_SEEDS += [int(s) for s in np.random.SeedSequence().generate_state(int(os.environ["CV_FUZZ_SHARDS"]))]


@pytest.mark.parametrize("seed", SHARD_SEEDS)
def test_random_shards_and_blocks(engine, seed):
    """cv_multi on random column blocks and 1-4 shards (engines on this device; across devices when the box
    has them) == the single engine, bit for bit with the path and epilogue pinned."""
    import torch
    from chromvar_b200 import _lib
    rng = np.random.default_rng(seed)
    c = _case(int(rng.integers(1 << 30)))
    X, sets, bg = c["X"], [s for s in c["sets"] if len(s)], c["bg"]
    C = sp.csc_matrix(X)
    P, N = C.shape
    e = orc.compute_expectations_core(C)
    ptr, idx = sets_to_csr(sets)
    n_dev = int(rng.integers(1, 5))
    ng = torch.cuda.device_count()
    devices = [int(d) % ng for d in range(n_dev)]
    cuts = sorted(set(rng.integers(1, N, int(rng.integers(0, 4))).tolist())) if N > 2 else []
    blocks = [C[:, a:b].tocsc() for a, b in zip([0] + cuts, cuts + [N])]
    fixed = dict(path=2, dense_epilogue=int(rng.choice([1, 2])), small_n_gather=0, check_zero_peaks=1)
    try:
        for k, v in fixed.items():
            engine.set_option(k, v)
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
        try:
            one = engine.deviations(ptr, idx, bg, e, index_base=1)
        except _lib.ChromVARError as err:
            assert err.code == _lib.CV_ERR_UNSUPPORTED, err          # forced dense operand declined (multiplicity)
            return
    finally:
        for k, v in dict(path=0, dense_epilogue=0, small_n_gather=-1).items():
            engine.set_option(k, v)
    ref = orc.compute_deviations_core(C, sets, bg, e, intermediate=True)
    masks = _ill_conditioned(C, sets, bg, e, ref)
    m = _lib.MultiHandle(devices, seed=3)
    try:
        for k, v in fixed.items():
            m.set_option(k, v)
        m.set_counts(blocks)
        assert np.array_equal(m.expectations(), e)
        for res in (m.deviations(ptr, idx, bg, e), m.deviations_csc(blocks, ptr, idx, bg, None)):
            for key in ("matches", "overlap"):
                assert np.array_equal(res[key], one[key], equal_nan=True), (seed, key, devices, cuts)
            for key in ("dev", "z"):
                if fixed["dense_epilogue"] == 2:                      # fp64 epilogue: the same arithmetic on every operand type
                    assert np.array_equal(res[key], one[key], equal_nan=True), (seed, key, devices, cuts)
                else:                                                # a shard without large counts may take the fixed-point form
                    skip = masks[0] if key == "z" else masks[1]
                    assert_close(np.where(skip, 0.0, res[key]), np.where(skip, 0.0, one[key]), key)
            assert_close(res["variability"], orc.row_sds(res["z"], True), "variability")
    finally:
        m.close()


# ---- getBackgroundPeaks: the deterministic front half and the draws ----------------------------------------------
BG_SEEDS = list(range(900, 916))
if os.environ.get("CV_FUZZ_BG"):
    BG_SEEDS += [int(s) for s in np.random.SeedSequence().generate_state(int(os.environ["CV_FUZZ_BG"]))]


@pytest.mark.parametrize("seed", BG_SEEDS)
def test_random_background_front_half(engine, seed):
    """Random peak counts / bias vectors (ties, rounded values, heavy tails), grid sizes and kernel widths:
    whitening, grid, bin membership, bin density and bin-probability matrix against the oracle; draws
    inside 1..P, only into bins the probability matrix allows, and reproducible from the seed."""
    rng = np.random.default_rng(seed)
    P = int(rng.choice([60, 400, 3000, 20000]))
    N = int(rng.choice([2, 50]))
    bs = int(rng.choice([5, 20, 50, 64]))
    w = float(rng.choice([0.05, 0.1, 1.0]))
    B = int(rng.choice([1, 7, 50]))
    tot = np.maximum(1, rng.lognormal(rng.uniform(1, 5), rng.uniform(0.3, 1.5), P)).astype(np.int64)
    X = np.zeros((P, N))
    X[:, 0] = tot // 2
    X[:, N - 1] += tot - tot // 2
    kind = int(rng.integers(4))
    bias = (rng.normal(0.5, 0.1, P), np.round(rng.normal(0.5, 0.1, P), 2), rng.beta(2, 5, P),
            rng.choice(np.linspace(0.3, 0.7, 9), P))[kind]
    C = sp.csc_matrix(X)
    engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    engine.set_seed(seed)
    bg, dbg = engine.background_peaks(bias, B, w, bs, debug=True)
    fh = orc.background_front_half(orc.get_fragments_per_peak(C), bias, bs)
    what = "seed %d: P %d bs %d w %g B %d bias kind %d" % (seed, P, bs, w, B, kind)
    assert np.max(np.abs(dbg["trans_norm_mat"] - fh["trans_norm_mat"])) < 1e-9, what
    # a point within rounding of a cell border may legitimately fall on either side
    mism = np.nonzero(dbg["bin_membership"] != fh["bin_membership"])[0]
    if mism.size:
        grid = fh["bin_data"]
        for p in mism:
            d = np.sqrt(((grid - fh["trans_norm_mat"][p]) ** 2).sum(axis=1))
            a, b = d[dbg["bin_membership"][p] - 1], d[fh["bin_membership"][p] - 1]
            assert abs(a - b) < 1e-9 * max(1.0, b), "%s: peak %d binned differently (%g vs %g)" % (what, p, a, b)
    else:
        assert np.array_equal(dbg["bin_density"], fh["bin_density"]), what
        Pbin = orc.bin_probability_matrix(orc.bin_weight_matrix(fh["bin_data"], w), fh["bin_density"])
        # a source bin too far from every occupied bin for a narrow kernel has all weights underflow (0 / 0):
        # such rows may only belong to unoccupied bins, which no peak ever draws from
        nan = np.isnan(Pbin)
        assert np.array_equal(np.isnan(dbg["bin_prob"]), nan), what
        assert not np.any(nan[fh["bin_density"] > 0]), what
        assert np.max(np.abs(np.where(nan, 0.0, dbg["bin_prob"] - Pbin))) < 1e-11, what
    assert bg.shape == (P, B) and bg.min() >= 1 and bg.max() <= P, what
    # every draw lands in a bin with positive probability from its source bin
    src = dbg["bin_membership"] - 1
    dst = dbg["bin_membership"][bg - 1] - 1
    assert np.all(dbg["bin_prob"][src[:, None], dst] > 0), what
    engine.set_seed(seed)
    assert np.array_equal(engine.background_peaks(bias, B, w, bs), bg), what
