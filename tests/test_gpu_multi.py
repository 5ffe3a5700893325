"""Several GPUs behind one call (cv_multi, chromvar_b200/csrc/cv_multi.cu): the counts arrive as
column blocks, the library shards the cells, runs one engine per shard on its own host thread and
does the two cross-shard reductions (per-peak totals, variability partials) itself.

The property under test is "sharded == unsharded": every output column depends only on its own
counts column plus replicated inputs (R/compute_deviations.R:486-504), so the sharded result must
equal the single-engine result BITWISE when both run the same epilogue arithmetic -- with shards
whose local per-peak totals contain zeros (check on the reduced totals) and with the expectation
taken from the reduced totals.

One-GPU boxes run the shards as several engines on the same device (devices = [0, 0, 0]: separate
handles, streams and host threads, reductions through the same kernels); with >= 2 devices visible
the same tests also run across devices, over peer memory and over the staged-copy route.
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


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _layouts():
    out = [("one_device_three_shards", [0, 0, 0], 1)]
    n = _n_gpus()
    if n >= 2:
        out.append(("two_devices_p2p", [0, 1], 1))
        out.append(("two_devices_staged", [0, 1], 0))
    if n >= 4:
        out.append(("four_devices", [0, 1, 2, 3], 1))
    return out


def _problem(seed=5):
    from chromvar_b200 import synthetic
    d = synthetic.make_config("small")                 # 20 000 peaks x 2 000 cells, 64 annotations
    C = d["counts"].tocsc()
    rng = np.random.default_rng(seed)
    P, N = C.shape
    # peaks that have fragments in the LAST cells only: zero in every shard but the last
    late = rng.choice(P, 40, replace=False)
    C = C.tolil()
    C[late, :] = 0
    for p in late:
        C[p, N - 1 - int(rng.integers(20))] = 2.0
    C = C.tocsc()
    C.eliminate_zeros()
    C.sort_indices()
    return d, C


def _split(C, cuts):
    cuts = [0] + list(cuts) + [C.shape[1]]
    return [C[:, a:b].tocsc() for a, b in zip(cuts[:-1], cuts[1:])]


@pytest.fixture(scope="module")
def problem():
    return _problem()


@pytest.mark.parametrize("layout", _layouts(), ids=lambda l: l[0])
def test_sharded_equals_unsharded_bitwise(engine, problem, layout):
    from chromvar_b200 import _lib
    name, devices, p2p = layout
    d, C = problem
    P, N = C.shape
    ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
    K = len(ptr) - 1
    # unsharded: one engine
    for key, v in (("path", 2), ("dense_epilogue", 1), ("dense_fp4", -1), ("cell_chunk", 0), ("check_zero_peaks", 1)):
        engine.set_option(key, v)
    try:
        engine.set_counts_csc(P, N, C.indptr, C.indices, C.data)
        engine.set_seed(123)
        e1 = engine.expectations()
        fpp1, fps1, tot1 = engine.fragments()
        bg1 = engine.background_peaks(d["bias"], 50, 0.1, 50)
        one = engine.deviations(ptr, idx, bg1, e1, index_base=1)
    finally:
        for key, v in (("path", 0), ("dense_epilogue", 0)):
            engine.set_option(key, v)
    # sharded, from three ragged column blocks (one of them a single column)
    blocks = _split(C, [700, 701])
    m = _lib.MultiHandle(devices, seed=123)
    try:
        m.set_option("multi_p2p", p2p)
        m.set_option("path", 2)
        m.set_option("dense_epilogue", 1)
        m.set_counts(blocks)
        b = m.shards()
        assert b[0] == 0 and b[-1] == N and np.all(np.diff(b) > 0)
        fpp, fps, tot = m.fragments()
        assert np.array_equal(fpp, fpp1) and np.array_equal(fps, fps1) and tot == tot1
        e = m.expectations()
        assert np.array_equal(e, e1)
        # the norm / group variants are sums over cells of per-cell terms: accumulated per shard, added
        # on the first device (fp64, another order than the single engine's atomics)
        grp = (np.arange(N) * 7 % 4).astype(np.int32)
        for nrm, gr in ((True, None), (False, grp), (True, grp)):
            got = m.expectations(norm=nrm, group=gr, n_groups=4 if gr is not None else 0)
            want = engine.expectations(norm=nrm, group=gr, n_groups=4 if gr is not None else 0)
            assert np.allclose(got, want, rtol=1e-12, atol=0), (nrm, gr is not None)
        assert np.allclose(m.expectations(norm=True, group=grp, n_groups=4),
                           orc.compute_expectations_core(C, norm=True, group=grp), rtol=1e-12, atol=0)
        bg = m.background_peaks(d["bias"], 50, 0.1, 50)
        assert np.array_equal(bg, bg1)                   # same seed, same global totals: same matrix
        res = m.deviations(ptr, idx, bg, e, index_base=1)
        for key in ("dev", "z", "matches", "overlap"):
            assert np.array_equal(res[key], one[key], equal_nan=True), key
        # variability over ALL cells: merged partials vs the single engine's (merge order differs)
        assert_close(res["variability"], one["variability"], "variability")
        assert_close(res["variability"], orc.row_sds(res["z"], True), "variability vs row_sds")
        # the one-call form on host data: streamed per device, reductions afterwards
        res2 = m.deviations_csc(blocks, ptr, idx, bg, e, index_base=1)
        for key in ("dev", "z", "matches", "overlap"):
            assert np.array_equal(res2[key], one[key], equal_nan=True), key
        assert_close(res2["variability"], one["variability"], "variability (one call)")
        # ... and with the default expectation (NULL): global totals first, then the deviations
        res3 = m.deviations_csc(blocks, ptr, idx, bg, None, index_base=1)
        for key in ("dev", "z"):
            assert np.array_equal(res3[key], one[key], equal_nan=True), key
        st = [m.stats(i) for i in range(len(devices))]
        assert all(s["path"] == 1 for s in st)
    finally:
        m.close()
    # against the oracle on a few annotations (the single engine's result is covered elsewhere)
    sets = [idx[ptr[k]:ptr[k + 1]] for k in (0, 7, 31)]
    ref = orc.compute_deviations_core(C, sets, bg1, e1)
    assert_close(res["z"][[0, 7, 31]], ref["z"], "z vs oracle")


def test_zero_peak_is_checked_on_the_reduced_totals(problem):
    from chromvar_b200 import _lib
    d, C = problem
    P, N = C.shape
    ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
    C0 = C.tolil()
    C0[17, :] = 0                                        # no fragment anywhere
    C0 = C0.tocsc()
    bg = np.asfortranarray(np.random.default_rng(1).integers(1, P + 1, size=(P, 50)).astype(np.int32))
    e = np.full(P, 1.0 / P)
    m = _lib.MultiHandle([0, 0], seed=1)
    try:
        with pytest.raises(_lib.ChromVARError, match="All peaks must have at least one fragment"):
            m.deviations_csc(_split(C0, [900]), ptr, idx, bg, e)
        m.set_counts(_split(C0, [900]))
        with pytest.raises(_lib.ChromVARError, match="All peaks must have at least one fragment"):
            m.deviations(ptr, idx, bg, e)
        # bad input in one block is reported with the reference's message
        bad = C.copy()
        bad.data = bad.data.copy()
        bad.data[-3] = -1.0
        with pytest.raises(_lib.ChromVARError, match="counts"):
            m.set_counts(_split(bad, [900]))
    finally:
        m.close()


def test_more_devices_than_cells_and_single_block():
    from chromvar_b200 import _lib
    rng = np.random.default_rng(2)
    P, N = 300, 2
    C = sp.csc_matrix(rng.integers(1, 4, size=(P, N)).astype(np.float64))
    sets = [rng.choice(np.arange(1, P + 1), 30, replace=False) for _ in range(3)]
    ptr = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    idx = np.concatenate(sets).astype(np.int32)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, 50)).astype(np.int32))
    e = orc.compute_expectations_core(C)
    m = _lib.MultiHandle([0, 0, 0], seed=1)
    try:
        res = m.deviations_csc(C, ptr, idx, bg, e)
        assert list(m.shards()) == [0, 1, 2, 2]
        ref = orc.compute_deviations_core(C, sets, bg, e)
        assert_close(res["dev"], ref["dev"], "deviations")
        assert_close(res["z"], ref["z"], "z")
    finally:
        m.close()


def test_two_engines_on_one_gpu_do_not_wait_for_each_other_forever():
    """The e2m1 GEMM's producers rendezvous through a global counter, which assumes all CTA pairs of a
    launch are resident.  Two engines on ONE GPU whose GEMMs each want every SM break that assumption
    (each gets part of the SMs and would wait for the rest forever): the rendezvous gives up after a
    few ms and the launch runs free.  Results must not change."""
    from chromvar_b200 import _lib, synthetic
    d = synthetic.make_config("small", N=5120, K=96)
    C = d["counts"].tocsc()
    P, N = C.shape
    ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
    rng = np.random.default_rng(4)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, 50)).astype(np.int32))
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    one = _lib.Handle(0, seed=1)
    one.set_option("path", 2)
    one.set_option("dense_epilogue", 1)
    one.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    ref = one.deviations(ptr, idx, bg, e)
    assert one.stats()["operand_bits"] == 4
    one.close()
    m = _lib.MultiHandle([0, 0], seed=1)
    try:
        m.set_option("path", 2)
        m.set_option("dense_epilogue", 1)
        m.set_counts([C])
        for _ in range(4):                                # both shards' GEMMs are in flight together
            res = m.deviations(ptr, idx, bg, e)
            for key in ("dev", "z"):
                assert np.array_equal(res[key], ref[key], equal_nan=True), key
    finally:
        m.close()


def test_dense_matrix_is_split_into_column_slabs(engine):
    from chromvar_b200 import _lib
    rng = np.random.default_rng(12)
    P, N, K = 1500, 333, 9
    X = rng.poisson(0.4, (P, N)).astype(np.float64)
    X[X.sum(axis=1) == 0, N - 1] = 1                      # some peaks live in the last slab only
    sets = [rng.choice(np.arange(1, P + 1), int(rng.integers(5, 200)), replace=False) for _ in range(K)]
    ptr = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    idx = np.concatenate(sets).astype(np.int32)
    bg = np.asfortranarray(rng.integers(1, P + 1, size=(P, 50)).astype(np.int32))
    engine.set_counts_dense(X)
    e = engine.expectations()
    one = engine.deviations(ptr, idx, bg, e)
    devs = [0, 1, 0] if _n_gpus() >= 2 else [0, 0, 0]
    m = _lib.MultiHandle(devs, seed=1)
    try:
        m.set_counts_dense(X)
        assert list(m.shards()) == [0, 111, 222, 333]
        fpp, fps, tot = m.fragments()
        assert np.array_equal(fpp, X.sum(axis=1)) and np.array_equal(fps, X.sum(axis=0)) and tot == X.sum()
        assert np.array_equal(m.expectations(), e)
        res = m.deviations(ptr, idx, bg, e)
        for key in ("dev", "z"):
            assert np.array_equal(res[key], one[key], equal_nan=True), key
        m.set_counts_dense(X[:, :2].astype(np.int32))     # more devices than cells
        assert list(m.shards()) == [0, 1, 2, 2]
        assert np.array_equal(m.fragments()[1], X[:, :2].sum(axis=0))
    finally:
        m.close()


def test_host_api_on_several_devices(golden_mini):
    """chromvar_b200.api with set_devices(): the same calls, the cells sharded behind them."""
    from chromvar_b200 import api, synergy
    g = golden_mini
    C = sp.csc_matrix((g["counts_x"], g["counts_i"], g["counts_p"]), shape=tuple(g["counts_dim"]))
    A = sp.csc_matrix((g["ix_x"].astype(bool), g["ix_i"], g["ix_p"]), shape=tuple(g["ix_dim"]))
    P, N = C.shape
    grp = np.array(["a", "b", "c"])[np.arange(N) % 3]

    def run():
        counts = api.SummarizedExperiment({"counts": C}, rowData={"bias": g["bias"]}, colData={"grp": grp})
        ix = api.SummarizedExperiment({"matches": A})
        api.invalidate_counts_cache()
        api.set_seed(77)
        bg = api.getBackgroundPeaks(counts)
        out = dict(bg=bg, e=api.computeExpectations(counts), en=api.computeExpectations(counts, norm=True),
                   eg=api.computeExpectations(counts, group="grp"), fpp=api.getFragmentsPerPeak(counts),
                   fps=api.getFragmentsPerSample(counts))
        dev = api.computeDeviations(counts, ix, background_peaks=bg)
        out["z"], out["dev"] = api.deviationScores(dev), api.deviations(dev)
        api.invalidate_counts_cache()
        dev2 = api.computeDeviations(C, ix, background_peaks=bg)          # counts not resident: the one-call form
        out["z_onecall"] = api.deviationScores(dev2)
        dense = api.computeDeviations(np.asarray(C.todense()), ix, background_peaks=bg)
        out["z_dense"] = api.deviationScores(dense)
        out["var"] = api.computeVariability(dev, bootstrap_error=False)["variability"]
        sets = [np.array([1, 5, 9]), np.array([2, 3])]
        out["vb"] = synergy.compute_variability_batch(C, sets, bg, out["e"])
        return out

    one = run()
    saved = (api._engine, api._device)
    api._engine = api._device = None
    try:
        api.set_devices([0, 1] if _n_gpus() >= 2 else [0, 0])
        assert isinstance(api.get_engine(), api._MultiEngine)
        two = run()
        assert list(api.get_engine().m.shards())[1] not in (0, N)
    finally:
        api.get_engine().close()
        api._engine, api._device = saved
        api.invalidate_counts_cache()
    for key in ("bg", "e", "fpp", "fps"):
        assert np.array_equal(one[key], two[key], equal_nan=True), key
    for key in ("en", "eg"):
        assert np.allclose(one[key], two[key], rtol=1e-11, atol=0, equal_nan=True), key
    # the automatic path choice sees other shapes (and measured rates) per shard: same sums, the epilogue
    # arithmetic may be another one -- bitwise equality is test_sharded_equals_unsharded_bitwise's subject
    for key in ("z", "dev", "z_onecall", "z_dense", "var", "vb"):
        assert_close(two[key], one[key], key)
    assert_close(one["z_onecall"], one["z"], "one call vs resident")
