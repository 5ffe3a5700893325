import os, sys
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from chromvar_b200 import _lib
from oracle import chromvar_oracle as orc
seed = int(sys.argv[1])
rng = np.random.default_rng(seed)
P = int(rng.choice([60, 400, 3000, 20000])); N = int(rng.choice([2, 50])); bs = int(rng.choice([5, 20, 50, 64]))
w = float(rng.choice([0.05, 0.1, 1.0])); B = int(rng.choice([1, 7, 50]))
tot = np.maximum(1, rng.lognormal(rng.uniform(1, 5), rng.uniform(0.3, 1.5), P)).astype(np.int64)
X = np.zeros((P, N)); X[:, 0] = tot // 2; X[:, N - 1] += tot - tot // 2
kind = int(rng.integers(4))
bias = (rng.normal(0.5, 0.1, P), np.round(rng.normal(0.5, 0.1, P), 2), rng.beta(2, 5, P), rng.choice(np.linspace(0.3, 0.7, 9), P))[kind]
C = sp.csc_matrix(X)
eng = _lib.Handle(0, seed=seed)
eng.set_counts_csc(P, N, C.indptr, C.indices, C.data)
bg, dbg = eng.background_peaks(bias, B, w, bs, debug=True)
fh = orc.background_front_half(orc.get_fragments_per_peak(C), bias, bs)
Pbin = orc.bin_probability_matrix(orc.bin_weight_matrix(fh["bin_data"], w), fh["bin_density"])
D = np.abs(dbg["bin_prob"] - Pbin)
i, j = np.unravel_index(np.argmax(D), D.shape)
print("P", P, "bs", bs, "w", w, "max diff", D.max(), "at", i, j, "ours", dbg["bin_prob"][i, j], "ref", Pbin[i, j])
print("row sums ours", dbg["bin_prob"][i].sum(), "ref", Pbin[i].sum(), "density i", fh["bin_density"][i], "density j", fh["bin_density"][j])
nz = np.nonzero(Pbin[i])[0]
print("ref row nonzeros", nz[:10], Pbin[i][nz][:10]); nz2 = np.nonzero(dbg["bin_prob"][i])[0]; print("our row nonzeros", nz2[:10], dbg["bin_prob"][i][nz2][:10])
print("rows with nan ref", np.isnan(Pbin).any(axis=1).sum(), "ours", np.isnan(dbg["bin_prob"]).any(axis=1).sum())
