"""Small invocations of every kernel family through the C ABI, for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_driver.py

Shapes are tiny (the sanitizers slow the persistent tcgen05 kernels down by orders of magnitude);
every variant of the deviation path runs once: e2m1 / u8 CTA-pair / u8 single-CTA GEMMs with the
fixed-point and the fp64 epilogue, digit planes, the row-gather kernel, the builders behind them, the
streamed call, the sampler, kNN sampler, permuted counts, covariability, bootstrap and two shards.
Results are checked against the u8 / fp64 variant so a sanitizer-induced failure is visible too."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromvar_b200 import _lib, synthetic  # noqa: E402


def main():
    d = synthetic.make_config("tiny")                    # 2 000 peaks x 300 cells, 12 annotations
    C = d["counts"].tocsc()
    P, N = C.shape
    ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
    eng = _lib.Handle(0, seed=5)
    eng.set_counts_csc(P, N, C.indptr, C.indices, C.data)
    e = eng.expectations()
    bg = eng.background_peaks(d["bias"], 50, 0.1, 50)
    eng.background_peaks_knn(d["bias"], 10, 10)
    ref = None
    for name, opts in (("u8 pair fp64", dict(path=2, dense_fp4=0, dense_cta_pair=1, dense_epilogue=2)),
                       ("u8 pair fixed", dict(path=2, dense_fp4=0, dense_cta_pair=1, dense_epilogue=1)),
                       ("u8 single fp64", dict(path=2, dense_fp4=0, dense_cta_pair=0, dense_epilogue=2)),
                       ("u8 single fixed", dict(path=2, dense_fp4=0, dense_cta_pair=0, dense_epilogue=1)),
                       ("e2m1 fixed", dict(path=2, dense_fp4=1, dense_cta_pair=1, dense_epilogue=1)),
                       ("e2m1 fp64", dict(path=2, dense_fp4=1, dense_cta_pair=1, dense_epilogue=2)),
                       ("row gather", dict(path=1))):
        for k, v in opts.items():
            eng.set_option(k, v)
        got = eng.deviations(ptr, idx, bg, e, index_base=1, sums=True)
        st = eng.stats()
        print("[sanitize] %-16s path %d bits %d epilogue %d" % (name, st["path"], st["operand_bits"], st["epilogue"]), flush=True)
        if ref is None:
            ref = got
        else:
            assert np.array_equal(got["observed"], ref["observed"]) and np.array_equal(got["sampled"], ref["sampled"]), name
            ok = np.isfinite(ref["z"])
            assert np.array_equal(np.isnan(got["z"]), np.isnan(ref["z"])), name
            assert np.max(np.abs(got["z"][ok] - ref["z"][ok]) / np.maximum(np.abs(ref["z"][ok]), 1e-3)) < 1e-5, name
    # digit planes (counts above 255) and the streamed call
    eng.set_option("path", 2)
    eng.set_option("dense_fp4", -1)
    eng.set_option("dense_cta_pair", -1)
    eng.set_option("dense_epilogue", 0)
    C2 = C.copy()
    C2.data = C2.data.copy()
    C2.data[::97] = 700.0
    e2 = np.asarray(C2.sum(axis=1)).ravel() / C2.sum()
    eng.set_counts_csc(P, N, C2.indptr, C2.indices, C2.data)
    eng.deviations(ptr, idx, bg, e2, index_base=1)
    assert eng.stats()["planes"] == 2
    print("[sanitize] digit planes", flush=True)
    got = eng.deviations_csc(P, N, C.indptr, C.indices, C.data, ptr, idx, bg, e, index_base=1)
    ok = np.isfinite(ref["z"])
    assert np.max(np.abs(got["z"][ok] - ref["z"][ok]) / np.maximum(np.abs(ref["z"][ok]), 1e-3)) < 1e-5
    print("[sanitize] streamed call", flush=True)
    eng.deviations_compute(rebuild_counts_layout=True)
    eng.deviations_compute_async(rebuild_counts_layout=True)
    eng.deviations_wait()
    print("[sanitize] queued call", flush=True)
    # the rest of the path
    eng.permuted_counts(d["bias"])
    eng.deviations_covariability()
    eng.row_sds(got["z"], True)
    eng.row_sds_perm(got["z"], True, 8)
    eng.expectations(norm=True)
    print("[sanitize] permuted counts, covariability, bootstrap", flush=True)
    eng.close()
    m = _lib.MultiHandle([0, 0], seed=5)
    m.set_option("path", 2)
    res = m.deviations_csc([C[:, :120].tocsc(), C[:, 120:].tocsc()], ptr, idx, bg, e, index_base=1)
    assert np.max(np.abs(res["z"][ok] - ref["z"][ok]) / np.maximum(np.abs(ref["z"][ok]), 1e-3)) < 1e-5
    m.close()
    print("[sanitize] two shards", flush=True)
    print("[sanitize] done", flush=True)


if __name__ == "__main__":
    main()
