"""ctypes binding of oracle/_ref/libchromvar_ref.so -- TEST INFRASTRUCTURE ONLY.

That library is the reference's OWN native code: /root/reference/src/utils.cpp compiled
unmodified (oracle/Makefile) against stand-in Armadillo/Rcpp headers (oracle/shim/), plus R's
Mersenne-Twister `unif_rand()` / `set.seed()` (oracle/refbuild/ref_capi.cpp).  It pins the
restatements in chromvar_oracle.py (row_sds, row_sds_perm, ProbSampleReplace, bg_sample_helper,
euc_dist) and is the "reference sampler" of the chi-square tests.  Only tests/, smoke() and
bench.py's CPU legs may import this module.

`build()` needs /root/reference (present in the development container only); the built .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libchromvar_ref.so")
REFERENCE_SOURCE = "/root/reference/src/utils.cpp"

_lib = None


def build():
    """Compile the reference source where it lies (no-op when it is absent, e.g. on the GPU box)."""
    if not os.path.exists(REFERENCE_SOURCE):
        return LIB if os.path.exists(LIB) else None
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout + r.stderr)
    return LIB


def available():
    return os.path.exists(LIB)


def load():
    global _lib
    if _lib is None:
        if not available():
            build()
        if not available():
            raise RuntimeError("oracle/_ref/libchromvar_ref.so is missing and /root/reference is not here to build it")
        lib = C.CDLL(LIB)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        lib.ref_set_seed.argtypes = [C.c_uint32]
        lib.ref_set_seed.restype = None
        lib.ref_unif_rand.argtypes = [C.c_int64, dp]
        lib.ref_unif_rand.restype = None
        lib.ref_draws.restype = C.c_uint64
        lib.ref_row_sds.argtypes = [dp, C.c_int64, C.c_int64, C.c_int, dp]
        lib.ref_row_sds_perm.argtypes = [dp, C.c_int64, C.c_int64, C.c_int, dp]
        lib.ref_prob_sample_replace.argtypes = [C.c_int, dp, C.c_int64, ip]
        lib.ref_bg_sample_helper.argtypes = [ip, C.c_int64, dp, C.c_int64, dp, C.c_int64, ip]
        lib.ref_euc_dist.argtypes = [dp, C.c_int64, C.c_int64, dp]
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def set_seed(seed):
    """R's set.seed(seed) (Mersenne-Twister)."""
    load().ref_set_seed(int(seed) & 0xFFFFFFFF)


def unif_rand(n):
    """n draws of R's unif_rand() from the current state (== runif(n))."""
    out = np.empty(int(n), np.float64)
    load().ref_unif_rand(int(n), _d(out))
    return out


def draws():
    return int(load().ref_draws())


def row_sds(X, na_rm=False):
    """src/utils.cpp:9-25 (the reference's code)."""
    X = np.asfortranarray(X, dtype=np.float64)
    out = np.empty(X.shape[0])
    assert load().ref_row_sds(_d(X), X.shape[0], X.shape[1], int(bool(na_rm)), _d(out)) == 0
    return out


def row_sds_perm(X, na_rm=False):
    """src/utils.cpp:28-33; consumes N draws of unif_rand()."""
    X = np.asfortranarray(X, dtype=np.float64)
    out = np.empty(X.shape[0])
    assert load().ref_row_sds_perm(_d(X), X.shape[0], X.shape[1], int(bool(na_rm)), _d(out)) == 0
    return out


def prob_sample_replace(size, prob):
    """src/utils.cpp:38-77; 1-based indices; consumes `size` draws."""
    prob = np.ascontiguousarray(prob, dtype=np.float64)
    out = np.empty(int(size), np.int64)
    assert load().ref_prob_sample_replace(int(size), _d(prob), prob.size, _i(out)) == 0
    return out


def bg_sample_helper(bin_membership0, bin_p, bin_density, niterations):
    """src/utils.cpp:83-99; P x niterations, 1-based."""
    bm = np.ascontiguousarray(bin_membership0, dtype=np.int64)
    bp = np.asfortranarray(bin_p, dtype=np.float64)
    bd = np.ascontiguousarray(bin_density, dtype=np.float64)
    assert bp.shape == (bd.size, bd.size)
    out = np.empty((bm.size, int(niterations)), np.int64, order="F")
    assert load().ref_bg_sample_helper(_i(bm), bm.size, _d(bp), bd.size, _d(bd), int(niterations), _i(out)) == 0
    return out


def euc_dist(x):
    """src/utils.cpp:107-117."""
    x = np.asfortranarray(x, dtype=np.float64)
    out = np.empty((x.shape[0], x.shape[0]), np.float64, order="F")
    assert load().ref_euc_dist(_d(x), x.shape[0], x.shape[1], _d(out)) == 0
    return out
