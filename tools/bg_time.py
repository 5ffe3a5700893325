"""Device time of cv_background_peaks: python tools/bg_time.py"""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromvar_b200 import _lib
eng = _lib.Handle(0, seed=3)
rng = np.random.default_rng(0)
for P in (2000, 100000, 500000):
    tot = np.floor(np.maximum(1, rng.lognormal(3, 1, P)))
    C = sp.csc_matrix(np.stack([tot, np.ones(P)], axis=1))
    eng.set_counts_csc(P, 2, C.indptr, C.indices, C.data)
    bias = rng.normal(0.5, 0.1, P)
    for bs in (50, 100):
        eng.background_peaks(bias, 50, 0.1, bs)
        eng.background_peaks(bias, 50, 0.1, bs)
        print("P %d bs %d niterations 50: %.3f ms on the device" % (P, bs, eng.stats()["ms_background"]), flush=True)
