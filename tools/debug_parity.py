"""Parity probe on a scaled configuration: GPU (resident + streamed) vs the CPU oracle, printing where they differ."""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromvar_b200 import _lib, synthetic
from oracle import chromvar_oracle as orc

name = sys.argv[1]; P = int(sys.argv[2]); N = int(sys.argv[3]); K = int(sys.argv[4])
d = synthetic.make_config(name, P=P, N=N, K=K)
C = sp.csc_matrix(d["counts"]); ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
eng = _lib.Handle(0, seed=1)
eng.set_counts_csc(P, N, C.indptr, C.indices, C.data)
e = eng.expectations()
bg = eng.background_peaks(d["bias"], 50, 0.1, 50)
print("max count", C.data.max(), "nnz", C.nnz)
res = {}
for path in (2, 1):
    eng.set_option("path", path)
    res[path] = eng.deviations(ptr, idx, bg, e, sums=True)
    print("path", path, eng.stats()["path"], "planes", eng.stats()["planes"])
eng.set_option("path", 2)
res["csc"] = eng.deviations_csc(P, N, C.indptr, C.indices, C.data, ptr, idx, bg, e)
a, b = res[2], res[1]
print("dense vs row-gather: observed equal", np.array_equal(a["observed"], b["observed"]), "sampled equal", np.array_equal(a["sampled"], b["sampled"]))
def rel(x, y):
    m = ~(np.isnan(x) | np.isnan(y))
    r = np.abs(x[m] - y[m]) / np.maximum(np.abs(y[m]), 1e-3)
    return r.max() if r.size else 0.0
print("dense vs row-gather z", rel(a["z"], b["z"]), "dev", rel(a["dev"], b["dev"]))
print("streamed vs dense z equal", np.array_equal(res["csc"]["z"], a["z"], equal_nan=True))
S = min(K, 32)
sets = [idx[ptr[k]:ptr[k + 1]].astype(np.int64) for k in range(S)]
z_ref, dev_ref = orc.deviations_task_farm(C, sets, bg, e, os.cpu_count())
for nm, r in (("dense", a), ("row-gather", b)):
    print(nm, "vs oracle: z", rel(r["z"][:S], z_ref), "dev", rel(r["dev"][:S], dev_ref))
m = ~np.isnan(z_ref)
r = np.abs(a["z"][:S] - z_ref) / np.maximum(np.abs(z_ref), 1e-3)
r[~m] = 0
k, c = np.unravel_index(np.argmax(r), r.shape)
print("worst", k, c, "gpu z", a["z"][k, c], "ref z", z_ref[k, c], "gpu dev", a["dev"][k, c], "ref dev", dev_ref[k, c], "set size", len(sets[k]))
one = orc.compute_deviations_single(sets[k], sp.csr_matrix(C), bg, e, intermediate_results=True)
print("oracle obs", one["observed"][c], "gpu obs", a["observed"][k, c], "samp equal", np.array_equal(one["sampled"][:, c].astype(np.int64), a["sampled"][k, :, c]))
sd_ref = np.std(((one["sampled"] - one["sampled_expected"]) / one["sampled_expected"])[:, c], ddof=1) if "sampled_expected" in one else None
print("bg sd (oracle)", sd_ref)
