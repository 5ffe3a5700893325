"""Print the mismatching entries of tests/test_gpu_fuzz.py cases: python tools/fuzz_debug.py seed [seed ...]"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import test_gpu_fuzz as f  # noqa: E402
from chromvar_b200 import _lib  # noqa: E402
from oracle import chromvar_oracle as orc  # noqa: E402

eng = _lib.Handle(0, seed=1)
for s in [int(a) for a in sys.argv[1:]]:
    c = f._case(s)
    rng = np.random.default_rng(s + 1)
    C, ip, ii, d = f._csc_arrays(c["X"], c["dtype"], c["shuffle"], rng)
    e = orc.compute_expectations_core(C)
    live = [k for k in range(c["K"]) if len(c["sets"][k])]
    ref = orc.compute_deviations_core(C, [c["sets"][k] for k in live], c["bg"], e, intermediate=True)
    ill, edge = f._ill_conditioned(C, [c["sets"][k] for k in live], c["bg"], e, ref)
    print("   ill-conditioned entries: %d of %d" % (ill.sum(), ill.size))
    ptr, idx = f.sets_to_csr(c["sets"])
    eng.set_counts_csc(c["P"], c["N"], ip, ii, d)
    print("== seed", s, "P", c["P"], "N", c["N"], "K", c["K"], "B", c["B"], "top", c["top"], c["dtype"], flush=True)
    for name, opts in f.VARIANTS:
        for k, v in {**f.DEFAULTS, **opts}.items():
            eng.set_option(k, v)
        try:
            got = eng.deviations(ptr, idx, c["bg"], e, index_base=1, sums=True)
        except _lib.ChromVARError as err:
            print("  ", name, "declined:", err)
            continue
        st = eng.stats()
        z, zr = got["z"][live], ref["z"]
        dv, dr = got["dev"][live], ref["dev"]
        same_sums = np.array_equal(got["observed"][live], ref["observed"]) and np.array_equal(got["sampled"][live], ref["sampled"])
        bad = ~ill & ~((z == zr) | (np.isnan(z) & np.isnan(zr)) | (np.abs(z - zr) <= 1e-5 * np.maximum(np.abs(zr), 1e-3)))
        bad |= ~edge & (np.isnan(dv) != np.isnan(dr))
        print("  %-28s path %d bits %d epi %d few %d sums_equal %s bad z %d" % (name, st["path"], st["operand_bits"], st["epilogue"],
                                                                               st["few_cells"], same_sums, bad.sum()))
        for k, n in list(zip(*np.nonzero(bad)))[:4]:
            print("      k %d cell %d: z ours %r ref %r | dev ours %r ref %r | obs %d bg sums %s" % (
                k, n, z[k, n], zr[k, n], dv[k, n], dr[k, n], ref["observed"][k, n], ref["sampled"][k, :, n]))
            members = c["sets"][live[k]]
            exp_obs = e[members - 1].sum()
            exp_bg = [e[c["bg"][members - 1, j] - 1].sum() for j in range(c["B"])]
            print("        expectation obs %r bg %s fragments %r" % (exp_obs, exp_bg, C[:, n].sum()))
