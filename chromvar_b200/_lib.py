"""ctypes binding of libchromvar_b200.so (the C ABI in include/chromvar_b200.h).

This is the only place the Python host side touches native code.  There is no fallback:
if the shared library is missing it is built in-tree with nvcc; if that fails, or if no
sm_100 GPU is visible when a handle is created, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libchromvar_b200.so")

CV_OK, CV_ERR_INVALID, CV_ERR_CUDA, CV_ERR_OOM, CV_ERR_STATE, CV_ERR_UNSUPPORTED = range(6)
CV_I32, CV_I64, CV_F64, CV_U8, CV_F32 = range(5)
NA_REAL_BITS = 0x7FF00000000007A2

_DTYPE_TAG = {np.dtype(np.int32): CV_I32, np.dtype(np.int64): CV_I64,
              np.dtype(np.float64): CV_F64, np.dtype(np.uint8): CV_U8,
              np.dtype(np.float32): CV_F32}

# every symbol include/chromvar_b200.h declares (tests check the library exports them all)
EXPORTS = [
    "cv_create", "cv_destroy", "cv_last_error", "cv_version", "cv_set_seed", "cv_set_option", "cv_stream",
    "cv_set_counts_csc", "cv_set_counts_dense", "cv_fragments", "cv_peak_totals_device",
    "cv_peak_totals_commit", "cv_expectations", "cv_background_peaks", "cv_background_peaks_knn", "cv_deviations",
    "cv_deviations_csc",
    "cv_deviations_upload", "cv_deviations_compute", "cv_deviations_compute_async", "cv_deviations_wait",
    "cv_deviations_download",
    "cv_variability_partials_device", "cv_merge_variability", "cv_row_sds", "cv_row_sds_perm",
    "cv_prob_sample_replace", "cv_bg_sample_helper", "cv_euc_dist", "cv_stats", "cv_e2m1_terms", "cv_host_narrow",
    "cv_variability_merge_device", "cv_permuted_counts", "cv_permuted_counts_fetch", "cv_deviations_covariability",
    "cv_create_multi", "cv_destroy_multi", "cv_multi_last_error", "cv_multi_n_devices", "cv_multi_peer_access",
    "cv_multi_handle", "cv_multi_set_seed", "cv_multi_set_option", "cv_multi_stats", "cv_multi_set_counts",
    "cv_multi_set_counts_dense", "cv_multi_shards", "cv_multi_fragments", "cv_multi_expectations", "cv_multi_background_peaks",
    "cv_multi_deviations", "cv_multi_deviations_csc",
]


class StageStats(C.Structure):
    _fields_ = [
        ("ms_counts_layout", C.c_double), ("ms_anno_prep", C.c_double),
        ("ms_contract", C.c_double), ("ms_finalize", C.c_double),
        ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("ms_background", C.c_double),
        ("contract_flops", C.c_double),
        ("contract_updates", C.c_int64), ("contract_bytes", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("kernel_launches", C.c_int64),
        ("path", C.c_int32), ("cell_tile", C.c_int32), ("n_cell_tiles", C.c_int32),
        ("acc_bytes", C.c_int32), ("planes", C.c_int32), ("epilogue", C.c_int32),
        ("operand_bits", C.c_int32), ("overflow_columns", C.c_int32), ("few_cells", C.c_int32), ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class CscBlock(C.Structure):
    """cv_csc_block: one column block of the counts (a dgCMatrix's @p / @i / @x)."""
    _fields_ = [("n_cols", C.c_int64), ("colptr", C.c_void_p), ("colptr_type", C.c_int),
                ("rowidx", C.c_void_p), ("x", C.c_void_p), ("x_type", C.c_int)]


class ChromVARError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_lib = None


def load(build_if_missing=True):
    """dlopen the library (building it in-tree first when absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise ChromVARError(CV_ERR_STATE, "libchromvar_b200.so is not built: run "
                                "`python -m chromvar_b200.build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    sig = {
        "cv_create": (i32, [i32, u64, C.POINTER(vp)]),
        "cv_destroy": (None, [vp]),
        "cv_last_error": (C.c_char_p, [vp]),
        "cv_version": (i32, []),
        "cv_set_seed": (i32, [vp, u64]),
        "cv_stream": (vp, [vp]),
        "cv_set_option": (i32, [vp, C.c_char_p, i64]),
        "cv_set_counts_csc": (i32, [vp, i64, i64, vp, i32, vp, vp, i32]),
        "cv_set_counts_dense": (i32, [vp, i64, i64, vp, i32]),
        "cv_fragments": (i32, [vp, vp, vp, vp]),
        "cv_peak_totals_device": (i32, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "cv_peak_totals_commit": (i32, [vp]),
        "cv_expectations": (i32, [vp, i32, vp, i32, vp]),
        "cv_background_peaks": (i32, [vp, vp, i32, dbl, i32, vp, vp, vp, vp, vp]),
        "cv_background_peaks_knn": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "cv_deviations": (i32, [vp, i64, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "cv_deviations_csc": (i32, [vp, i64, i64, vp, i32, vp, vp, i32, i64, vp, vp, i32, vp, i32, vp,
                                    vp, vp, vp, vp, vp]),
        "cv_deviations_upload": (i32, [vp, i64, vp, vp, i32, vp, i32, vp]),
        "cv_deviations_compute": (i32, [vp, i32, i32]),
        "cv_deviations_compute_async": (i32, [vp, i32]),
        "cv_deviations_wait": (i32, [vp]),
        "cv_deviations_download": (i32, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "cv_variability_partials_device": (i32, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "cv_merge_variability": (i32, [vp, i32, i64, i32, vp]),
        "cv_row_sds": (i32, [vp, i64, i64, vp, i32, vp]),
        "cv_row_sds_perm": (i32, [vp, i64, i64, vp, i32, i32, vp]),
        "cv_prob_sample_replace": (i32, [vp, i64, i64, vp, vp]),
        "cv_bg_sample_helper": (i32, [vp, i64, vp, i64, vp, vp, i32, vp]),
        "cv_euc_dist": (i32, [vp, i64, i64, vp, vp]),
        "cv_stats": (i32, [vp, C.POINTER(StageStats)]),
        "cv_host_narrow": (i32, [vp, i32, i64, vp, i32]),
        "cv_variability_merge_device": (i32, [vp, vp, i32]),
        "cv_permuted_counts": (i32, [vp, vp, dbl, i32, vp, C.POINTER(i64)]),
        "cv_permuted_counts_fetch": (i32, [vp, i64, vp, vp]),
        "cv_deviations_covariability": (i32, [vp, i64, i64, vp, vp]),
        "cv_create_multi": (i32, [vp, i32, u64, C.POINTER(vp)]),
        "cv_destroy_multi": (None, [vp]),
        "cv_multi_last_error": (C.c_char_p, [vp]),
        "cv_multi_n_devices": (i32, [vp]),
        "cv_multi_peer_access": (i32, [vp]),
        "cv_multi_handle": (vp, [vp, i32]),
        "cv_multi_set_seed": (i32, [vp, u64]),
        "cv_multi_set_option": (i32, [vp, C.c_char_p, i64]),
        "cv_multi_stats": (i32, [vp, i32, C.POINTER(StageStats)]),
        "cv_multi_set_counts": (i32, [vp, i64, C.POINTER(CscBlock), i32]),
        "cv_multi_shards": (i32, [vp, vp]),
        "cv_multi_fragments": (i32, [vp, vp, vp, vp]),
        "cv_multi_expectations": (i32, [vp, i32, vp, i32, vp]),
        "cv_multi_set_counts_dense": (i32, [vp, i64, i64, vp, i32]),
        "cv_multi_background_peaks": (i32, [vp, vp, i32, dbl, i32, vp]),
        "cv_multi_deviations": (i32, [vp, i64, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]),
        "cv_multi_deviations_csc": (i32, [vp, i64, C.POINTER(CscBlock), i32, i64, vp, vp, i32, vp, i32, vp,
                                        vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError if the export is missing: loud on purpose
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_narrow(x, threads=4):
    """cv_host_narrow: (ok, bytes) -- the host-side conversion of wide count values to bytes that
    cv_deviations_csc applies before the upload.  Host only, needs no GPU."""
    x = np.ascontiguousarray(x)
    t = {np.dtype(np.float64): CV_F64, np.dtype(np.float32): CV_F32, np.dtype(np.int32): CV_I32}[x.dtype]
    out = np.empty(x.size, np.uint8)
    rc = load().cv_host_narrow(_ptr(x), t, x.size, _ptr(out), int(threads))
    if rc < 0:
        raise ChromVARError(-rc, "cv_host_narrow: bad arguments")
    return bool(rc), out


def _carr(a, dtype):
    """C-contiguous (or F-contiguous for 2-D) array of dtype without copying when possible."""
    a = np.asarray(a)
    if a.dtype != np.dtype(dtype):
        a = a.astype(dtype)
    if a.ndim <= 1:
        return np.ascontiguousarray(a)
    return a


class DeviceBuffer:
    """A device pointer owned by a Handle, exposed through __cuda_array_interface__ so that
    torch / NCCL can operate on it in place (torch.as_tensor(buf, device='cuda'))."""

    def __init__(self, ptr, n, typestr, owner):
        self.ptr, self.n, self.owner = ptr, n, owner
        self.__cuda_array_interface__ = {
            "shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3,
            "strides": None,
        }


class Handle:
    """One GPU's engine state (cv_handle).  Mirrors the ABI one to one; the R-facing API in
    chromvar_b200.api is built on top of it."""

    def __init__(self, device=0, seed=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.cv_create(int(device), int(seed) & (2 ** 64 - 1), C.byref(h))
        if rc != CV_OK:
            raise ChromVARError(rc, self.lib.cv_last_error(None).decode())
        self._h = h
        self.device = device
        self.P = self.N = self.K = self.B = 0

    @classmethod
    def borrowed(cls, ptr, device=0):
        """A view of an engine somebody else owns (cv_multi_handle): same methods, close() does not
        destroy it."""
        self = cls.__new__(cls)
        self.lib = load()
        self._h = C.c_void_p(ptr) if not isinstance(ptr, C.c_void_p) else ptr
        self._borrowed = True
        self.device = device
        self.P = self.N = self.K = self.B = 0
        return self

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                self.lib.cv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != CV_OK:
            raise ChromVARError(rc, self.lib.cv_last_error(self._h).decode())

    # ---- counts ---------------------------------------------------------------------------
    def set_counts_csc(self, P, N, colptr, rowidx, x):
        colptr = np.ascontiguousarray(colptr)
        if colptr.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            colptr = colptr.astype(np.int64)
        rowidx = _carr(rowidx, np.int32)
        x = np.ascontiguousarray(x)
        if x.dtype not in _DTYPE_TAG or x.dtype == np.dtype(np.int64):
            x = x.astype(np.float64)
        self._check(self.lib.cv_set_counts_csc(self._h, P, N, _ptr(colptr), _DTYPE_TAG[colptr.dtype],
                                               _ptr(rowidx), _ptr(x), _DTYPE_TAG[x.dtype]))
        self.P, self.N = int(P), int(N)

    def set_counts_dense(self, x):
        x = np.asarray(x)
        if x.dtype not in (np.dtype(np.float64), np.dtype(np.float32), np.dtype(np.int32)):
            x = x.astype(np.float64)
        x = np.asfortranarray(x)
        P, N = x.shape
        self._check(self.lib.cv_set_counts_dense(self._h, P, N, _ptr(x), _DTYPE_TAG[x.dtype]))
        self.P, self.N = int(P), int(N)

    def fragments(self):
        fpp = np.empty(self.P, np.float64)
        fps = np.empty(self.N, np.float64)
        tot = np.empty(1, np.float64)
        self._check(self.lib.cv_fragments(self._h, _ptr(fpp), _ptr(fps), _ptr(tot)))
        return fpp, fps, float(tot[0])

    def peak_totals_device(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.cv_peak_totals_device(self._h, C.byref(p), C.byref(n)))
        return DeviceBuffer(p.value, n.value, "<i8", self)

    def peak_totals_commit(self):
        self._check(self.lib.cv_peak_totals_commit(self._h))

    def expectations(self, norm=False, group=None, n_groups=0):
        out = np.empty(self.P, np.float64)
        g = None if group is None else _carr(group, np.int32)
        self._check(self.lib.cv_expectations(self._h, int(bool(norm)), _ptr(g), int(n_groups), _ptr(out)))
        return out

    # ---- background -----------------------------------------------------------------------
    def background_peaks(self, bias, niterations=50, w=0.1, bs=50, debug=False):
        bias = _carr(bias, np.float64)
        if bias.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "length(bias) must equal nrow(counts)")
        bg = np.empty((self.P, niterations), np.int32, order="F")
        if not debug:
            self._check(self.lib.cv_background_peaks(self._h, _ptr(bias), niterations, w, bs, _ptr(bg),
                                                     None, None, None, None))
            return bg
        nb = bs * bs
        trans = np.empty((self.P, 2), np.float64, order="F")
        memb = np.empty(self.P, np.int32)
        dens = np.empty(nb, np.float64)
        binprob = np.empty((nb, nb), np.float64)
        self._check(self.lib.cv_background_peaks(self._h, _ptr(bias), niterations, w, bs, _ptr(bg),
                                                 _ptr(trans), _ptr(memb), _ptr(dens), _ptr(binprob)))
        return bg, dict(trans_norm_mat=trans, bin_membership=memb, bin_density=dens, bin_prob=binprob)

    def background_peaks_knn(self, bias, niterations=50, window=500, debug=False):
        """get_background_peaks_alternative (R/test_sets.R:291-330): uniform draws among the `window`
        nearest other peaks in the whitened (fragments per peak, bias) plane."""
        bias = _carr(bias, np.float64)
        if bias.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "length(bias) must equal nrow(counts)")
        bg = np.empty((self.P, niterations), np.int32, order="F")
        if not debug:
            self._check(self.lib.cv_background_peaks_knn(self._h, _ptr(bias), niterations, window, _ptr(bg), None, None))
            return bg
        trans = np.empty((self.P, 2), np.float64, order="F")
        nn = np.empty((self.P, window), np.int32)
        self._check(self.lib.cv_background_peaks_knn(self._h, _ptr(bias), niterations, window, _ptr(bg),
                                                     _ptr(trans), _ptr(nn)))
        return bg, dict(tmp_vals=trans, nn_idx=nn)

    # ---- deviations -----------------------------------------------------------------------
    @staticmethod
    def _prep_dev_inputs(annoptr, annoidx, bg, expectation):
        annoptr = _carr(annoptr, np.int64)
        annoidx = _carr(annoidx, np.int32)
        bg = np.asarray(bg)
        if bg.dtype != np.dtype(np.int32) or not bg.flags.f_contiguous:
            bg = np.asfortranarray(bg.astype(np.int32, copy=False))
        expectation = _carr(expectation, np.float64)
        return annoptr, annoidx, bg, expectation

    def deviations_upload(self, annoptr, annoidx, bg, expectation, index_base=1):
        annoptr, annoidx, bg, expectation = self._prep_dev_inputs(annoptr, annoidx, bg, expectation)
        K = annoptr.shape[0] - 1
        if bg.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "nrow(background_peaks) must equal nrow(counts)")
        if expectation.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "length(expectation) must equal nrow(counts)")
        self._check(self.lib.cv_deviations_upload(self._h, K, _ptr(annoptr), _ptr(annoidx), index_base,
                                                  _ptr(bg), bg.shape[1], _ptr(expectation)))
        self.K, self.B = K, bg.shape[1]

    def deviations_compute(self, rebuild_counts_layout=False, keep_sums=False):
        self._check(self.lib.cv_deviations_compute(self._h, int(rebuild_counts_layout), int(keep_sums)))

    def deviations_compute_async(self, rebuild_counts_layout=False):
        """cv_deviations_compute without host synchronisation (a replay of the last checked call on
        the same resident inputs); deviations_wait() or any reading call settles it."""
        self._check(self.lib.cv_deviations_compute_async(self._h, int(rebuild_counts_layout)))

    def deviations_wait(self):
        self._check(self.lib.cv_deviations_wait(self._h))

    def deviations_download(self, sums=False, out=None, vectors_only=False):
        K, N, B = self.K, self.N, self.B
        if vectors_only:
            out = dict(dev=None, z=None, matches=np.empty(K, np.float64), overlap=np.empty(K, np.float64),
                       variability=np.empty(K, np.float64))
        if out is None:
            out = dict(dev=np.empty((K, N), np.float64, order="F"),
                       z=np.empty((K, N), np.float64, order="F"),
                       matches=np.empty(K, np.float64), overlap=np.empty(K, np.float64),
                       variability=np.empty(K, np.float64))
        obs = samp = None
        if sums:
            obs = np.empty((K, N), np.int64, order="F")
            samp = np.empty((K, B, N), np.int64)
            out["observed"], out["sampled"] = obs, samp
        self._check(self.lib.cv_deviations_download(self._h, _ptr(out["dev"]), _ptr(out["z"]),
                                                    _ptr(out["matches"]), _ptr(out["overlap"]),
                                                    _ptr(out["variability"]), _ptr(obs), _ptr(samp)))
        return out

    def deviations(self, annoptr, annoidx, bg, expectation, index_base=1, sums=False):
        """cv_deviations: the one-call form (H2D + compute + D2H)."""
        annoptr, annoidx, bg, expectation = self._prep_dev_inputs(annoptr, annoidx, bg, expectation)
        K = annoptr.shape[0] - 1
        N, B = self.N, bg.shape[1]
        if bg.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "nrow(background_peaks) must equal nrow(counts)")
        if expectation.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "length(expectation) must equal nrow(counts)")
        out = dict(dev=np.empty((K, N), np.float64, order="F"), z=np.empty((K, N), np.float64, order="F"),
                   matches=np.empty(K, np.float64), overlap=np.empty(K, np.float64),
                   variability=np.empty(K, np.float64))
        obs = samp = None
        if sums:
            obs = np.empty((K, N), np.int64, order="F")
            samp = np.empty((K, B, N), np.int64)
            out["observed"], out["sampled"] = obs, samp
        self._check(self.lib.cv_deviations(self._h, K, _ptr(annoptr), _ptr(annoidx), index_base, _ptr(bg),
                                           B, _ptr(expectation), _ptr(out["dev"]), _ptr(out["z"]),
                                           _ptr(out["matches"]), _ptr(out["overlap"]),
                                           _ptr(out["variability"]), _ptr(obs), _ptr(samp)))
        self.K, self.B = K, B
        return out

    def deviations_csc(self, P, N, colptr, rowidx, x, annoptr, annoidx, bg, expectation=None, index_base=1,
                       out=None):
        """cv_deviations_csc: compute_deviations_core as one call on host data (counts, annotations,
        background, expectation in; deviations / z out), copies overlapped with the kernels."""
        colptr = np.ascontiguousarray(colptr)
        if colptr.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            colptr = colptr.astype(np.int64)
        rowidx = _carr(rowidx, np.int32)
        x = np.ascontiguousarray(x)
        if x.dtype not in _DTYPE_TAG or x.dtype == np.dtype(np.int64):
            x = x.astype(np.float64)
        annoptr = _carr(annoptr, np.int64)
        annoidx = _carr(annoidx, np.int32)
        bg = np.asarray(bg)
        if bg.dtype != np.dtype(np.int32) or not bg.flags.f_contiguous:
            bg = np.asfortranarray(bg.astype(np.int32, copy=False))
        if expectation is not None:
            expectation = _carr(expectation, np.float64)
            if expectation.shape[0] != P:
                raise ChromVARError(CV_ERR_INVALID, "length(expectation) must equal nrow(counts)")
        if bg.shape[0] != P:
            raise ChromVARError(CV_ERR_INVALID, "nrow(background_peaks) must equal nrow(counts)")
        K, B = annoptr.shape[0] - 1, bg.shape[1]
        if out is None:
            out = dict(dev=np.empty((K, N), np.float64, order="F"), z=np.empty((K, N), np.float64, order="F"),
                       matches=np.empty(K, np.float64), overlap=np.empty(K, np.float64),
                       variability=np.empty(K, np.float64))
        self._check(self.lib.cv_deviations_csc(
            self._h, P, N, _ptr(colptr), _DTYPE_TAG[colptr.dtype], _ptr(rowidx), _ptr(x), _DTYPE_TAG[x.dtype],
            K, _ptr(annoptr), _ptr(annoidx), index_base, _ptr(bg), B, _ptr(expectation),
            _ptr(out["dev"]), _ptr(out["z"]), _ptr(out["matches"]), _ptr(out["overlap"]),
            _ptr(out["variability"])))
        self.P, self.N, self.K, self.B = int(P), int(N), K, B
        return out

    def variability_merge_device(self, d_parts_ptr, n_parts):
        """cv_variability_merge_device: merge an all-gathered device buffer of n_parts K x 4 tables into
        this handle's variability vector, on the handle's stream (no host synchronisation)."""
        self._check(self.lib.cv_variability_merge_device(self._h, C.c_void_p(int(d_parts_ptr)), int(n_parts)))

    def variability_partials_device(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.cv_variability_partials_device(self._h, C.byref(p), C.byref(n)))
        return DeviceBuffer(p.value, n.value, "<f8", self)

    # ---- the reference's registered native routines ---------------------------------------
    def row_sds(self, x, na_rm=False):
        x = np.asfortranarray(np.asarray(x, np.float64))
        K, N = x.shape
        out = np.empty(K, np.float64)
        self._check(self.lib.cv_row_sds(self._h, K, N, _ptr(x), int(bool(na_rm)), _ptr(out)))
        return out

    def row_sds_perm(self, x, na_rm=False, n_rep=1):
        x = np.asfortranarray(np.asarray(x, np.float64))
        K, N = x.shape
        out = np.empty((n_rep, K), np.float64)
        self._check(self.lib.cv_row_sds_perm(self._h, K, N, _ptr(x), int(bool(na_rm)), int(n_rep), _ptr(out)))
        return out

    def prob_sample_replace(self, size, prob):
        prob = _carr(prob, np.float64)
        out = np.empty(size, np.int32)
        self._check(self.lib.cv_prob_sample_replace(self._h, size, prob.shape[0], _ptr(prob), _ptr(out)))
        return out

    def bg_sample_helper(self, bin_membership, bin_p, bin_density, niterations):
        bm = _carr(bin_membership, np.int32)
        bin_p = np.asfortranarray(np.asarray(bin_p, np.float64))
        dens = _carr(bin_density, np.float64)
        out = np.empty((bm.shape[0], niterations), np.int32, order="F")
        self._check(self.lib.cv_bg_sample_helper(self._h, bm.shape[0], _ptr(bm), dens.shape[0], _ptr(bin_p),
                                                 _ptr(dens), niterations, _ptr(out)))
        return out

    def euc_dist(self, x):
        x = np.asfortranarray(np.asarray(x, np.float64))
        n, d = x.shape
        out = np.empty((n, n), np.float64, order="F")
        self._check(self.lib.cv_euc_dist(self._h, n, d, _ptr(x), _ptr(out)))
        return out

    def permuted_counts(self, bias, w=0.1, bs=50):
        """cv_permuted_counts: one permuted assay of the resident counts (getPermutedData), gathered on
        the device; returns (colptr int64[N+1], rowidx int32, x float64) of the CSC result."""
        bias = _carr(bias, np.float64)
        if bias.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "length(bias) must equal nrow(counts)")
        colptr = np.empty(self.N + 1, np.int64)
        nnz = C.c_int64()
        self._check(self.lib.cv_permuted_counts(self._h, _ptr(bias), float(w), int(bs), _ptr(colptr), C.byref(nnz)))
        rowidx, x = np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float64)
        self._check(self.lib.cv_permuted_counts_fetch(self._h, nnz.value, _ptr(rowidx), _ptr(x)))
        return colptr, rowidx, x

    def deviations_covariability(self, z=None):
        """cv_deviations_covariability: K x K normalised covariance of the rows of z (host matrix), or of
        the z still resident from the last deviations call (z = None)."""
        if z is None:
            K = self.K
            out = np.empty((K, K), np.float64, order="F")
            self._check(self.lib.cv_deviations_covariability(self._h, 0, 0, None, _ptr(out)))
            return out
        z = np.asfortranarray(np.asarray(z, np.float64))
        K, N = z.shape
        out = np.empty((K, K), np.float64, order="F")
        self._check(self.lib.cv_deviations_covariability(self._h, K, N, _ptr(z), _ptr(out)))
        return out

    def set_seed(self, seed):
        self._check(self.lib.cv_set_seed(self._h, int(seed) & (2 ** 64 - 1)))

    def set_option(self, key, value):
        self._check(self.lib.cv_set_option(self._h, key.encode(), int(value)))

    def stats(self):
        s = StageStats()
        self._check(self.lib.cv_stats(self._h, C.byref(s)))
        return s.as_dict()

    def stream(self):
        return self.lib.cv_stream(self._h)


class MultiHandle:
    """cv_multi: one process driving several GPUs (the call the R shim binds).  Counts are given as
    a list of column blocks -- scipy CSC matrices with the same number of rows, e.g. the pieces of an
    atlas whose nonzeros exceed one dgCMatrix -- or a single matrix."""

    def __init__(self, devices, seed=0):
        self.lib = load()
        devices = [int(d) for d in devices]
        arr = (C.c_int * len(devices))(*devices)
        m = C.c_void_p()
        rc = self.lib.cv_create_multi(arr, len(devices), int(seed) & (2 ** 64 - 1), C.byref(m))
        if rc != CV_OK:
            raise ChromVARError(rc, self.lib.cv_multi_last_error(None).decode())
        self._m = m
        self.devices = devices
        self.P = self.N = self.K = self.B = 0

    def close(self):
        if getattr(self, "_m", None):
            self.lib.cv_destroy_multi(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != CV_OK:
            raise ChromVARError(rc, self.lib.cv_multi_last_error(self._m).decode())

    @staticmethod
    def _blocks(mats):
        """(ctypes array of cv_csc_block, keep-alive list, P, N)."""
        if not isinstance(mats, (list, tuple)):
            mats = [mats]
        keep, P, N = [], None, 0
        arr = (CscBlock * len(mats))()
        for b, M in enumerate(mats):
            if P is None:
                P = int(M.shape[0])
            elif int(M.shape[0]) != P:
                raise ChromVARError(CV_ERR_INVALID, "all column blocks must have the same number of rows")
            cp = np.ascontiguousarray(M.indptr)
            if cp.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
                cp = cp.astype(np.int64)
            ri = _carr(M.indices, np.int32)
            x = np.ascontiguousarray(M.data)
            if x.dtype not in _DTYPE_TAG or x.dtype == np.dtype(np.int64):
                x = x.astype(np.float64)
            keep += [cp, ri, x]
            arr[b].n_cols = int(M.shape[1])
            arr[b].colptr = cp.ctypes.data
            arr[b].colptr_type = _DTYPE_TAG[cp.dtype]
            arr[b].rowidx = ri.ctypes.data
            arr[b].x = x.ctypes.data
            arr[b].x_type = _DTYPE_TAG[x.dtype]
            N += int(M.shape[1])
        return arr, keep, P, N

    def n_devices(self):
        return self.lib.cv_multi_n_devices(self._m)

    def handle(self, i=0):
        """Device i's engine as a (borrowed) Handle: the cell-independent routines run there."""
        h = Handle.borrowed(self.lib.cv_multi_handle(self._m, int(i)), self.devices[i])
        h.P, h.N, h.K, h.B = self.P, 0, self.K, self.B
        return h

    def peer_access(self):
        return bool(self.lib.cv_multi_peer_access(self._m))

    def set_seed(self, seed):
        self._check(self.lib.cv_multi_set_seed(self._m, int(seed) & (2 ** 64 - 1)))

    def set_option(self, key, value):
        self._check(self.lib.cv_multi_set_option(self._m, key.encode(), int(value)))

    def stats(self, i=0):
        s = StageStats()
        self._check(self.lib.cv_multi_stats(self._m, int(i), C.byref(s)))
        return s.as_dict()

    def set_counts(self, mats):
        arr, keep, P, N = self._blocks(mats)
        self._check(self.lib.cv_multi_set_counts(self._m, P, arr, len(arr)))
        self.P, self.N = P, N

    def set_counts_dense(self, x):
        x = np.asarray(x)
        if x.dtype not in (np.dtype(np.float64), np.dtype(np.float32), np.dtype(np.int32)):
            x = x.astype(np.float64)
        x = np.asfortranarray(x)
        P, N = x.shape
        self._check(self.lib.cv_multi_set_counts_dense(self._m, P, N, _ptr(x), _DTYPE_TAG[x.dtype]))
        self.P, self.N = int(P), int(N)

    def shards(self):
        b = np.zeros(len(self.devices) + 1, np.int64)
        self._check(self.lib.cv_multi_shards(self._m, _ptr(b)))
        return b

    def fragments(self):
        fpp, fps, tot = np.empty(self.P), np.empty(self.N), np.empty(1)
        self._check(self.lib.cv_multi_fragments(self._m, _ptr(fpp), _ptr(fps), _ptr(tot)))
        return fpp, fps, float(tot[0])

    def expectations(self, norm=False, group=None, n_groups=0):
        e = np.empty(self.P)
        g = None if group is None else _carr(group, np.int32)
        if g is not None and g.shape[0] != self.N:
            raise ChromVARError(CV_ERR_INVALID, "group must have one label per cell")
        self._check(self.lib.cv_multi_expectations(self._m, int(bool(norm)), _ptr(g), int(n_groups), _ptr(e)))
        return e

    def background_peaks(self, bias, niterations=50, w=0.1, bs=50):
        bias = _carr(bias, np.float64)
        if bias.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "length(bias) must equal nrow(counts)")
        bg = np.empty((self.P, niterations), np.int32, order="F")
        self._check(self.lib.cv_multi_background_peaks(self._m, _ptr(bias), niterations, w, bs, _ptr(bg)))
        return bg

    def _outputs(self, K, N, out):
        if out is None:
            out = dict(dev=np.empty((K, N), np.float64, order="F"), z=np.empty((K, N), np.float64, order="F"),
                       matches=np.empty(K, np.float64), overlap=np.empty(K, np.float64),
                       variability=np.empty(K, np.float64))
        return out

    def deviations(self, annoptr, annoidx, bg, expectation, index_base=1, out=None):
        annoptr, annoidx, bg, expectation = Handle._prep_dev_inputs(annoptr, annoidx, bg, expectation)
        K, B = annoptr.shape[0] - 1, bg.shape[1]
        if bg.shape[0] != self.P or expectation.shape[0] != self.P:
            raise ChromVARError(CV_ERR_INVALID, "nrow(background_peaks) / length(expectation) must equal nrow(counts)")
        out = self._outputs(K, self.N, out)
        self._check(self.lib.cv_multi_deviations(self._m, K, _ptr(annoptr), _ptr(annoidx), index_base, _ptr(bg), B,
                                                 _ptr(expectation), _ptr(out["dev"]), _ptr(out["z"]),
                                                 _ptr(out["matches"]), _ptr(out["overlap"]), _ptr(out["variability"])))
        self.K, self.B = K, B
        return out

    def deviations_csc(self, mats, annoptr, annoidx, bg, expectation=None, index_base=1, out=None, blocks=None):
        """cv_multi_deviations_csc: compute_deviations_core as ONE call on host data, all devices.
        blocks: the tuple _blocks(mats) returned earlier (a caller timing the call prepares it once)."""
        arr, keep, P, N = blocks if blocks is not None else self._blocks(mats)
        annoptr = _carr(annoptr, np.int64)
        annoidx = _carr(annoidx, np.int32)
        bg = np.asarray(bg)
        if bg.dtype != np.dtype(np.int32) or not bg.flags.f_contiguous:
            bg = np.asfortranarray(bg.astype(np.int32, copy=False))
        if expectation is not None:
            expectation = _carr(expectation, np.float64)
            if expectation.shape[0] != P:
                raise ChromVARError(CV_ERR_INVALID, "length(expectation) must equal nrow(counts)")
        if bg.shape[0] != P:
            raise ChromVARError(CV_ERR_INVALID, "nrow(background_peaks) must equal nrow(counts)")
        K, B = annoptr.shape[0] - 1, bg.shape[1]
        out = self._outputs(K, N, out)
        self._check(self.lib.cv_multi_deviations_csc(
            self._m, P, arr, len(arr), K, _ptr(annoptr), _ptr(annoidx), index_base, _ptr(bg), B, _ptr(expectation),
            _ptr(out["dev"]), _ptr(out["z"]), _ptr(out["matches"]), _ptr(out["overlap"]), _ptr(out["variability"])))
        self.P, self.N, self.K, self.B = P, N, K, B
        return out


def merge_variability(parts, na_rm=True):
    """parts: [n_parts, K, 4] (n_finite, mean, M2, n_nonfinite) -> sd[K]."""
    lib = load()
    parts = np.ascontiguousarray(parts, np.float64)
    n_parts, K = parts.shape[0], parts.shape[1]
    out = np.empty(K, np.float64)
    rc = lib.cv_merge_variability(_ptr(parts), n_parts, K, int(bool(na_rm)), _ptr(out))
    if rc != CV_OK:
        raise ChromVARError(rc, "cv_merge_variability failed")
    return out


def is_na(a):
    """True where a holds R's NA_real_ payload (as opposed to an ordinary NaN)."""
    return np.asarray(a, np.float64).view(np.uint64) == np.uint64(NA_REAL_BITS)
