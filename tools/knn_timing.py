"""Device time of cv_background_peaks_knn at the sizes of configs 2 and 4/5 (window 10 and the
reference's default 500).  Counts: one fragment total per peak is all the sampler reads."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from chromvar_b200 import _lib

eng = _lib.Handle(0, seed=1)
rng = np.random.default_rng(0)
for P in (100_000, 500_000):
    fpp = np.maximum(1, rng.lognormal(5.0, 1.0, P).astype(np.int64))
    C = sp.csc_matrix((fpp.astype(np.float64), (np.arange(P), np.zeros(P, np.int64))), shape=(P, 1))
    bias = np.round(np.clip(rng.normal(0.52, 0.09, P), 0.25, 0.85), 3)
    eng.set_counts_csc(P, 1, C.indptr, C.indices, C.data)
    for window, niter in ((10, 1), (10, 50), (500, 50)):
        eng.background_peaks_knn(bias, niter, window)
        t0 = time.perf_counter()
        bg = eng.background_peaks_knn(bias, niter, window)
        wall = time.perf_counter() - t0
        print("P=%d window=%d niter=%d: device %.2f ms, wall %.2f ms, self-draws %d" %
              (P, window, niter, eng.stats()["ms_background"], 1e3 * wall, int((bg == (np.arange(P) + 1)[:, None]).sum())), flush=True)
