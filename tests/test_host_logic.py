"""Host-side logic of the R-API mirror that needs no GPU: input marshalling, validation order
and messages (R/compute_deviations.R:362-379, R/utils.R:3-27,289-300), BH adjustment."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from chromvar_b200 import api
from oracle import chromvar_oracle as orc


def test_convert_to_ix_list_dense_and_sparse():
    rng = np.random.default_rng(0)
    M = rng.random((40, 6)) < 0.2
    ptr, idx = api.convert_to_ix_list(M)
    ref = orc.convert_to_ix_list(M)
    assert [list(idx[ptr[k]:ptr[k + 1]]) for k in range(6)] == [list(r) for r in ref]
    ptr2, idx2 = api.convert_to_ix_list(sp.csr_matrix(M))
    assert np.array_equal(ptr, ptr2) and np.array_equal(idx, idx2)
    # quirk kept: an explicitly stored FALSE counts as a match for sparse input (summary() triplets)
    S = sp.csc_matrix((np.array([True, False, True]), (np.array([0, 3, 5]), np.array([0, 0, 1]))), shape=(8, 2))
    ptr3, idx3 = api.convert_to_ix_list(S)
    assert list(ptr3) == [0, 2, 3] and list(idx3) == [1, 4, 6]


def test_annotation_validation_messages():
    C = np.ones((5, 3))
    bg = np.ones((5, 4), np.int32)
    e = np.full(5, 0.2)
    with pytest.raises(ValueError, match="Annotations are not valid"):
        api.computeDeviations(C, [[1.5, 2]], bg, e)
    with pytest.raises(ValueError, match="Annotations are not valid"):
        api.computeDeviations(C, [[], []], bg, e)
    with pytest.raises(ValueError, match="Annotations are not valid"):
        api.computeDeviations(C, [[0, 1]], bg, e)
    with pytest.raises(ValueError, match="Annotations are not valid"):
        api.computeDeviations(C, [[1, 6]], bg, e)
    with pytest.raises(ValueError, match="nrow"):
        api.computeDeviations(C, [[1, 2]], np.ones((4, 4), np.int32), e)
    with pytest.raises(ValueError, match="length\\(expectation\\)"):
        api.computeDeviations(C, [[1, 2]], bg, np.full(4, 0.25))
    with pytest.raises(TypeError, match="background_peaks"):
        api.computeDeviations(C, [[1, 2]])        # matrix method has no default (:298-353)


def test_counts_and_matches_check():
    with pytest.raises(ValueError):
        api.counts_check(api.SummarizedExperiment({"notcounts": np.ones((2, 2))}))
    with pytest.raises(ValueError):
        api.counts_check(api.SummarizedExperiment({"counts": np.array([[1.0, -1.0], [0, 1]])}))
    with pytest.raises(ValueError):
        api.counts_check(api.SummarizedExperiment({"counts": np.ones((3, 1))}))
    with pytest.raises(ValueError):
        api.matches_check(api.SummarizedExperiment({"foo": np.ones((2, 2))}))
    se = api.SummarizedExperiment({"motifMatches": np.eye(2, dtype=bool)})
    assert api.annotationMatches(se) is se.assays["motifMatches"]
    with pytest.raises(ValueError):
        api.chromVARDeviations({"z": np.ones((1, 1))})


def test_group_argument_validation():
    C = np.ones((4, 3))
    with pytest.raises(ValueError, match="Group must be vector"):
        api.computeExpectations(C, group=[1, 2])


def test_bh_adjust_matches_oracle():
    rng = np.random.default_rng(1)
    p = rng.random(50)
    p[[3, 9]] = np.nan
    assert np.allclose(api._p_adjust_bh(p), orc.p_adjust_bh(p), equal_nan=True)


def test_e2m1_term_decomposition():
    """The e2m1 operands hold {0,1,2,3,4,6} only: every other count / multiplicity is a sum of such
    terms (cv_dense_fp4.cu).  The decomposition must be exact for every value a u8 count or a
    4-bit tile can hold, use the documented number of terms, and encode each term correctly."""
    import ctypes as C
    from chromvar_b200 import _lib
    lib = _lib.load()
    lib.cv_e2m1_terms.restype = C.c_int
    lib.cv_e2m1_terms.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]
    value_of_code = {0: 0.0, 1: 0.5, 2: 1.0, 3: 1.5, 4: 2.0, 5: 3.0, 6: 4.0, 7: 6.0}
    terms = (C.c_uint32 * 64)()
    codes = (C.c_uint32 * 64)()
    for v in list(range(0, 300)) + [1000, 4095, 65535]:
        n = lib.cv_e2m1_terms(v, terms, codes, 64) if v < 300 else lib.cv_e2m1_terms(v, None, None, 1 << 20)
        if v >= 300:
            assert n == v // 6 + (0 if v % 6 == 0 else 2 if v % 6 == 5 else 1)
            continue
        assert n >= 1
        t = [terms[i] for i in range(n)]
        assert sum(t) == v, (v, t)
        assert all(x in (1, 2, 3, 4, 6) for x in t) or v == 0
        assert [value_of_code[codes[i]] for i in range(n)] == [float(x) for x in t]
        assert n == (1 if v in (0, 1, 2, 3, 4, 6) else v // 6 + (0 if v % 6 == 0 else 2 if v % 6 == 5 else 1))
    assert lib.cv_e2m1_terms(65, terms, codes, 4) == -1


def test_host_narrow():
    """The host-side conversion of a dgCMatrix's double x slot to bytes (cv_hostnarrow.cpp): exact for
    integers 0..255, and anything else -- a large count, a fraction, a negative, NaN, Inf -- is reported
    so that chunk travels in its original type.  Slices of 2^18 elements on several threads."""
    from chromvar_b200 import _lib
    rng = np.random.default_rng(0)
    n = (1 << 20) + 12345                         # several slices and a ragged tail
    v = rng.integers(0, 256, n)
    for dt in (np.float64, np.float32, np.int32):
        for threads in (1, 3):
            ok, out = _lib.host_narrow(v.astype(dt), threads)
            assert ok and np.array_equal(out, v.astype(np.uint8)), (dt, threads)
    ok, out = _lib.host_narrow(np.zeros(0), 2)
    assert ok and out.size == 0
    for bad, where in ((256.0, 5), (0.5, n - 1), (-1.0, 300000), (np.nan, 7), (np.inf, n // 2), (1e300, 11), (-0.5, 1 << 18)):
        x = v.astype(np.float64)
        x[where] = bad
        ok, _ = _lib.host_narrow(x, 3)
        assert not ok, (bad, where)
    x = v.astype(np.float64)
    x[123] = -0.0                                 # negative zero is the integer 0
    ok, out = _lib.host_narrow(x, 2)
    assert ok and out[123] == 0
    ok, _ = _lib.host_narrow(np.array([7, 300, 2], np.int32), 1)
    assert not ok


def test_device_generator_of_synthetic_counts_on_cpu():
    """synthetic.make_counts_device (the atlas shard's generator, torch) on the CPU device: a valid CSC
    matrix of the requested density with the same value distribution as the numpy generator."""
    from chromvar_b200 import synthetic
    d = synthetic.make_config("small", device="cpu")
    C = d["counts"]
    P, N = C.shape
    assert (P, N) == (20_000, 2_000) and C.indptr[0] == 0 and C.indptr[-1] == C.nnz
    assert np.all(np.diff(C.indptr) >= 0) and C.indices.min() >= 0 and C.indices.max() < P
    for c in range(0, N, 101):                                   # rows strictly increasing inside a column
        r = C.indices[C.indptr[c]:C.indptr[c + 1]]
        assert np.all(np.diff(r) > 0)
    assert np.all(C.data >= 1) and np.all(C.data == np.round(C.data)) and C.data.max() <= 127
    assert np.all(np.asarray(C.sum(axis=1)).ravel() > 0)          # every peak has a fragment (R/compute_deviations.R:362)
    dens = C.nnz / (P * N)
    assert 0.02 < dens < 0.04
    ref = synthetic.make_config("small")["counts"]
    h1 = np.bincount(C.data.astype(int), minlength=6)[:6] / C.nnz
    h0 = np.bincount(ref.data.astype(int), minlength=6)[:6] / ref.nnz
    assert np.max(np.abs(h1 - h0)) < 0.03
    assert d["cfg"]["generator"].startswith("torch")


def test_bench_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` (the CPU restatement timed on the host's cores): one JSON line with
    the contract's keys; other ranks of a torchrun launch exit 0 without work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--quick",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "annotation_x_cell_zscores_per_sec_50bg" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["higher_is_better"] is True and d["config"]["workload"] == "small"
    env.update(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--quick",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_counts_fingerprint_sees_in_place_edits():
    """residency of the counts on the GPU must not be decided by object identity alone --
    an in-place edit keeps identity, shape and buffer addresses."""
    import scipy.sparse as sp
    from chromvar_b200 import api
    rng = np.random.default_rng(0)
    C = sp.random(200, 50, density=0.1, format="csc", random_state=np.random.RandomState(1),
                  data_rvs=lambda n: rng.integers(1, 5, n).astype(np.float64))
    k0 = api._fingerprint(C)
    assert api._fingerprint(C) == k0
    C.data[7] += 1.0                                  # same object, same buffers
    k1 = api._fingerprint(C)
    assert k1 != k0
    C *= 2                                            # scipy keeps the object, rescales the values
    assert api._fingerprint(C) != k1
    D = np.asfortranarray(rng.integers(0, 4, size=(30, 20)).astype(np.float64))
    kd = api._fingerprint(D)
    D[3, 3] += 2
    assert api._fingerprint(D) != kd


def test_every_background_call_draws_a_fresh_seed():
    """like R advancing its RNG, each getBackgroundPeaks call takes the next seed of the
    stream; set_seed() makes the sequence reproducible."""
    from chromvar_b200 import api
    api.set_seed(11)
    a = [api._next_sampler_seed() for _ in range(3)]
    api.set_seed(11)
    b = [api._next_sampler_seed() for _ in range(3)]
    assert a == b and len(set(a)) == 3
    api._seed_rng = None
    assert api._next_sampler_seed() != api._next_sampler_seed()
    api._seed_rng = None
