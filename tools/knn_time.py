"""Device time of cv_background_peaks_knn (whitening + sort + exact kNN + draws): python tools/knn_time.py"""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromvar_b200 import _lib
eng = _lib.Handle(0, seed=3)
rng = np.random.default_rng(0)
for P in (100000, 500000):
    tot = np.floor(np.maximum(1, rng.lognormal(3, 1, P)))
    C = sp.csc_matrix(np.stack([tot, np.ones(P)], axis=1))
    eng.set_counts_csc(P, 2, C.indptr, C.indices, C.data)
    bias = rng.normal(0.5, 0.1, P)
    for W in (10, 17, 24, 32, 64, 100, 500):
        t = []
        for ins in (0, 1):
            eng.set_option("knn_insertion", ins)
            eng.background_peaks_knn(bias, 50, W)
            a = eng.background_peaks_knn(bias, 50, W)
            t.append(eng.stats()["ms_background"])
            if ins == 0:
                ref = a
            else:
                assert np.array_equal(a, ref)
        print("P %d window %d: %.2f ms (exact insertion at every step: %.2f ms)" % (P, W, t[0], t[1]), flush=True)
