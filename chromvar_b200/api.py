"""Host-side mirror of the reference's R interface for the deviation engine.

R is not installed in this image, so the host side above the C ABI is Python with the
reference's names, argument meaning and error behaviour (the R shim a maintainer would use is
under rpkg/ and described in INTEGRATION.md).  Everything numerical happens in
libchromvar_b200.so; this module only validates, marshals and wraps, like the R methods do:

  computeDeviations   R/compute_deviations.R:66-68, methods :224-353, core :355-408
  computeExpectations R/compute_deviations.R:35-36, methods :184-219, core :415-448
  getBackgroundPeaks  R/background_peaks.R:78-114, core :116-156
  computeVariability  R/compute_variability.R:30-81
  getFragmentsPerPeak / getFragmentsPerSample / getTotalFragments  R/helper_functions.R:48-89
  deviations / deviationScores                                      R/compute_deviations.R:109-119
  chromVARDeviations                                                R/AllClasses.R:17-26

Index conventions follow R: annotation index lists and background_peaks are 1-based.
"""
import os
import warnings

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import ChromVARError

__all__ = [
    "SummarizedExperiment", "chromVARDeviations", "computeDeviations", "computeExpectations",
    "getBackgroundPeaks", "computeVariability", "getFragmentsPerPeak", "getFragmentsPerSample",
    "getTotalFragments", "deviations", "deviationScores", "annotationMatches", "get_engine",
    "set_device", "set_devices", "set_seed",
]

_MATCH_ASSAYS = ("annotationMatches", "annotation_matches", "motif_matches", "motifMatches", "matches")


# --------------------------------------------------------------------------------------------
# containers (the slice of SummarizedExperiment the hot path touches)
# --------------------------------------------------------------------------------------------
class SummarizedExperiment:
    """assays: dict name -> matrix (scipy sparse or ndarray); rowData / colData: dict of columns."""

    def __init__(self, assays, rowData=None, colData=None, rownames=None, colnames=None):
        self.assays = dict(assays)
        shapes = {tuple(a.shape) for a in self.assays.values()}
        if len(shapes) != 1:
            raise ValueError("all assays must have the same dimensions")
        self.rowData = dict(rowData or {})
        self.colData = dict(colData or {})
        self.rownames = None if rownames is None else list(rownames)
        self.colnames = None if colnames is None else list(colnames)

    @property
    def shape(self):
        return next(iter(self.assays.values())).shape

    def nrow(self):
        return self.shape[0]

    def ncol(self):
        return self.shape[1]


class chromVARDeviations(SummarizedExperiment):
    """R/AllClasses.R:17-26: a SummarizedExperiment whose assays are `deviations` and `z`."""

    def __init__(self, assays, rowData=None, colData=None, rownames=None, colnames=None):
        super().__init__(assays, rowData, colData, rownames, colnames)
        if "deviations" not in self.assays or "z" not in self.assays:   # validity, :20-26
            raise ValueError("chromVARDeviations needs assays 'deviations' and 'z'")


def deviations(object):
    """R/compute_deviations.R:109-112."""
    return object.assays["deviations"]


def deviationScores(object):
    """R/compute_deviations.R:116-119."""
    return object.assays["z"]


def annotationMatches(object):
    """R/get_annotations.R:52-72: the assay-name fallback chain."""
    for name in _MATCH_ASSAYS:
        if name in object.assays:
            if name == "motif_matches":
                warnings.warn("motif_matches assay deprecated; update motifmatchr")
            return object.assays[name]
    raise ValueError("No appropriately named assay. See Details section in man page")


# --------------------------------------------------------------------------------------------
# engine (one cv_handle per process; cells shard across processes, see distributed.py)
# --------------------------------------------------------------------------------------------
_engine = None
_device = None
_seed_rng = None     # set_seed(): the analogue of R's set.seed() for the samplers


def set_seed(seed):
    """Like set.seed() in R: every later getBackgroundPeaks() call draws a fresh sampler seed from
    this stream (the R shim does the same with sample.int(.Machine$integer.max, 1))."""
    global _seed_rng
    _seed_rng = np.random.default_rng(seed)


def _next_sampler_seed():
    """A fresh sampler seed for every getBackgroundPeaks call, like R advancing its RNG: without
    set_seed() the stream starts from OS entropy (distributed.init() replaces that with one seed
    broadcast from rank 0, so every rank draws the same background matrix)."""
    global _seed_rng
    if _seed_rng is None:
        _seed_rng = np.random.default_rng()
    return int(_seed_rng.integers(1, 2 ** 31 - 1))


def set_device(device):
    """One GPU per process (the default: LOCAL_RANK, else 0)."""
    global _device, _engine
    if _engine is not None and _device != device:
        _engine.close()
        _engine = None
        invalidate_counts_cache()
    _device = device


def set_devices(devices):
    """Several GPUs behind this one process (cv_multi, what the R shim's chromVAR_b200_init(devices)
    binds): the cells are sharded across `devices`, every function of this module keeps its meaning.
    A single id is the same as set_device()."""
    devices = [int(d) for d in devices]
    if len(devices) == 0:
        raise ValueError("devices must name at least one GPU")
    set_device(devices[0] if len(devices) == 1 else tuple(devices))


class _Block:
    """The CSC arrays of one column block, as cv_multi takes them."""
    def __init__(self, P, N, indptr, indices, data):
        self.shape, self.indptr, self.indices, self.data = (int(P), int(N)), indptr, indices, data


class _MultiEngine:
    """cv_multi with the method names api.py uses on a single engine.  Sharded: counts, fragments,
    expectations (all variants), deviations.  Cell independent (bias / totals in, matrix out, or a host
    matrix in): run on the first device, which holds the reduced per-peak totals."""

    def __init__(self, devices, seed):
        self.m = _lib.MultiHandle(devices, seed=seed)

    P = property(lambda self: self.m.P)
    N = property(lambda self: self.m.N)
    K = property(lambda self: self.m.K)

    def close(self):
        self.m.close()

    def first(self):
        return self.m.handle(0)

    def set_seed(self, seed):
        self.m.set_seed(seed)

    def set_option(self, key, value):
        self.m.set_option(key, value)

    def stats(self, i=0):
        return self.m.stats(i)

    def set_counts_csc(self, P, N, indptr, indices, data):
        self.m.set_counts([_Block(P, N, indptr, indices, data)])

    def set_counts_dense(self, a):
        self.m.set_counts_dense(a)

    def fragments(self):
        return self.m.fragments()

    def expectations(self, norm=False, group=None, n_groups=0):
        return self.m.expectations(norm=norm, group=group, n_groups=n_groups)

    def background_peaks(self, bias, niterations=50, w=0.1, bs=50):
        return self.m.background_peaks(bias, niterations, w, bs)

    def background_peaks_knn(self, bias, niterations=50, window=500):
        return self.first().background_peaks_knn(bias, niterations, window)

    def deviations(self, annoptr, annoidx, bg, expectation, index_base=1):
        return self.m.deviations(annoptr, annoidx, bg, expectation, index_base=index_base)

    def deviations_csc(self, P, N, indptr, indices, data, annoptr, annoidx, bg, expectation=None, index_base=1):
        return self.m.deviations_csc([_Block(P, N, indptr, indices, data)], annoptr, annoidx, bg, expectation,
                                     index_base=index_base)

    # the three-step form synergy.py uses (row SDs only: no K x N matrices come back)
    def deviations_upload(self, annoptr, annoidx, bg, expectation, index_base=1):
        self._pending = (annoptr, annoidx, bg, expectation, index_base)

    def deviations_compute(self):
        pass

    def deviations_download(self, vectors_only=False):
        annoptr, annoidx, bg, expectation, index_base = self._pending
        out = None
        if vectors_only:
            K = len(annoptr) - 1
            out = dict(dev=None, z=None, matches=np.empty(K), overlap=np.empty(K), variability=np.empty(K))
        return self.m.deviations(annoptr, annoidx, bg, expectation, index_base=index_base, out=out)

    def row_sds(self, x, na_rm=False):
        return self.first().row_sds(x, na_rm)

    def row_sds_perm(self, x, na_rm=False, n_rep=1):
        return self.first().row_sds_perm(x, na_rm, n_rep)

    def deviations_covariability(self, z=None):
        if z is None:
            raise ChromVARError(_lib.CV_ERR_STATE, "z is sharded across the devices: pass the matrix")
        return self.first().deviations_covariability(z)

    def permuted_counts(self, bias, w=0.1, bs=50):
        raise ChromVARError(_lib.CV_ERR_UNSUPPORTED, "getPermutedData needs a single device (set_device)")


def get_engine():
    """The process-wide engine.  Raises (never falls back to the CPU) when the CUDA library or a
    B200 is missing."""
    global _engine, _device
    if _engine is None:
        if _device is None:
            _device = int(os.environ.get("LOCAL_RANK", "0"))
        seed = int(np.random.SeedSequence().entropy) & (2 ** 63 - 1)
        _engine = _MultiEngine(list(_device), seed) if isinstance(_device, tuple) else _lib.Handle(_device, seed=seed)
    return _engine


class _CountsCache:
    """Remembers which matrix is resident on the device so that the usual sequence
    getBackgroundPeaks(x) -> computeExpectations(x) -> computeDeviations(x, ...) uploads once.
    Python objects can be edited in place (mat.data[...] = ..., mat *= 2) without changing identity
    or shape -- R's copy-on-modify cannot -- so residency is keyed on a content fingerprint, not on
    the object alone: shape, number of stored entries, buffer addresses and checksums of the values
    and of the structure (one pass over the host arrays, a few ms per 1e7 entries)."""
    obj = None
    key = None


def _fingerprint(mat):
    if sp.issparse(mat):
        m = mat
        if not hasattr(m, "indptr"):
            return ("sparse-other", id(mat))
        d, i, p = m.data, m.indices, m.indptr
        return ("csx", m.format, m.shape, int(m.nnz), d.__array_interface__["data"][0],
                i.__array_interface__["data"][0], p.__array_interface__["data"][0], str(d.dtype),
                float(np.sum(d, dtype=np.float64)) if d.size else 0.0,
                float(np.dot(d[::7].astype(np.float64), (np.arange(d[::7].size) % 8191) + 1.0)) if d.size else 0.0,
                int(np.sum(i[::3], dtype=np.int64)) if i.size else 0, int(p[-1]) if p.size else 0)
    a = np.asarray(mat)
    return ("dense", a.shape, a.__array_interface__["data"][0], str(a.dtype), float(np.sum(a, dtype=np.float64)),
            float(np.sum(a[::5, ::3], dtype=np.float64)))


def _is_resident(mat, eng):
    return (_CountsCache.obj is mat and eng.P == mat.shape[0] and eng.N == mat.shape[1] and
            _CountsCache.key == _fingerprint(mat))


def _mark_resident(mat):
    _CountsCache.obj = mat    # strong reference: the identity test stays meaningful
    _CountsCache.key = _fingerprint(mat)


def _is_matrix(x):
    return sp.issparse(x) or isinstance(x, np.ndarray)


def counts_check(object):
    """R/utils.R:3-12."""
    if not isinstance(object, SummarizedExperiment):
        raise TypeError("object must be a SummarizedExperiment")
    if "counts" not in object.assays:
        raise ValueError('"counts" %in% assayNames(object) is not TRUE')
    c = object.assays["counts"]
    mn = c.min() if c.shape[0] * c.shape[1] else 0
    if mn < 0:
        raise ValueError("min(assays(object)$counts) >= 0 is not TRUE")
    if object.ncol() <= 1:
        raise ValueError("ncol(object) > 1 is not TRUE")
    return object


def matches_check(object):
    """R/utils.R:14-27."""
    if not isinstance(object, SummarizedExperiment):
        raise TypeError("annotations must be a SummarizedExperiment")
    if not any(n in object.assays for n in _MATCH_ASSAYS):
        raise ValueError("no annotation matches assay (matches / annotationMatches / ...)")
    return object


def _counts_matrix(object):
    return counts_check(object).assays["counts"] if isinstance(object, SummarizedExperiment) else object


def _load_counts(mat):
    """Make `mat` the engine's resident counts matrix (H2D + K1) unless it already is."""
    eng = get_engine()
    if _is_resident(mat, eng):
        return eng
    if sp.issparse(mat):
        m = mat if sp.isspmatrix_csc(mat) else sp.csc_matrix(mat)
        eng.set_counts_csc(m.shape[0], m.shape[1], m.indptr, m.indices, m.data)
    else:
        a = np.asarray(mat)
        if a.ndim != 2:
            raise ValueError("counts must be a matrix")
        eng.set_counts_dense(a)
    _mark_resident(mat)
    return eng


def invalidate_counts_cache():
    """Forget which matrix is resident (the next call uploads again)."""
    _CountsCache.obj = None
    _CountsCache.key = None


# --------------------------------------------------------------------------------------------
# helper_functions.R
# --------------------------------------------------------------------------------------------
def getFragmentsPerPeak(object):
    return _load_counts(_counts_matrix(object)).fragments()[0]


def getFragmentsPerSample(object):
    return _load_counts(_counts_matrix(object)).fragments()[1]


def getTotalFragments(object):
    return _load_counts(_counts_matrix(object)).fragments()[2]


# --------------------------------------------------------------------------------------------
# computeExpectations
# --------------------------------------------------------------------------------------------
def computeExpectations(object, norm=False, group=None):
    mat = _counts_matrix(object)
    ncol = mat.shape[1]
    codes = None
    n_levels = 0
    if group is not None:
        if isinstance(object, SummarizedExperiment) and isinstance(group, str) and group in object.colData:
            group = object.colData[group]                                   # :203-204
        group = np.asarray(group)
        if group.ndim != 1 or group.shape[0] != ncol:
            if isinstance(object, SummarizedExperiment):
                raise ValueError("Group must be vector of length of columns of object, or a character "
                                 "vector referring to name of column incolData of object")
            raise ValueError("Group must be vector of length of columns of object")
        levels, codes = np.unique(group, return_inverse=True)              # as.factor
        n_levels = len(levels)
    eng = _load_counts(mat)
    return eng.expectations(norm=norm, group=codes, n_groups=n_levels)


# --------------------------------------------------------------------------------------------
# getBackgroundPeaks
# --------------------------------------------------------------------------------------------
def getBackgroundPeaks(object, bias=None, niterations=50, w=0.1, bs=50):
    """Returns the P x niterations matrix of 1-based background peak indices (the reference
    returns the same matrix with storage mode double)."""
    if isinstance(object, SummarizedExperiment):
        if bias is None:
            bias = object.rowData.get("bias")                               # :85, :97
    mat = _counts_matrix(object)
    if bias is None:
        raise ValueError("bias is missing (rowData(object)$bias)")
    bias = np.asarray(bias, np.float64)
    if bias.shape[0] != mat.shape[0]:
        raise ValueError("length(bias) == length(fragments_per_peak) is not TRUE")   # :123
    eng = _load_counts(mat)
    eng.set_seed(_next_sampler_seed())
    try:
        return eng.background_peaks(bias, int(niterations), float(w), int(bs))
    except ChromVARError as e:
        raise ValueError(str(e)) from e


# --------------------------------------------------------------------------------------------
# computeDeviations
# --------------------------------------------------------------------------------------------
def convert_to_ix_list(ix):
    """R/utils.R:289-300: (annoptr, annoidx) CSR-of-sets, 1-based, from a peaks x K matrix.
    For sparse input every *stored* entry is a match (summary() triplets; quirk kept)."""
    if sp.issparse(ix):
        m = ix if sp.isspmatrix_csc(ix) else sp.csc_matrix(ix)
        m.sort_indices()
        return m.indptr.astype(np.int64), (m.indices.astype(np.int32) + 1)
    a = np.asarray(ix)
    nz = a != 0
    counts = nz.sum(axis=0)
    annoptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cols, rows = np.nonzero(nz.T)                                           # column-major order
    return annoptr, (rows.astype(np.int32) + 1)


def _ix_list_to_csr(peak_indices):
    """list of K index vectors -> (annoptr, annoidx), validating like :373-379."""
    if len(peak_indices) == 0:
        raise ValueError("Annotations are not valid")
    arrs = [np.asarray(s).ravel() for s in peak_indices]
    sizes = np.array([a.size for a in arrs], np.int64)
    if sizes.sum() == 0:
        raise ValueError("Annotations are not valid")                      # is.null(unlist(...))
    flat = np.concatenate([a.astype(np.float64) for a in arrs])
    if not np.all(flat == np.floor(flat)) or not np.all(np.isfinite(flat)):
        raise ValueError("Annotations are not valid")
    annoptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return annoptr, flat.astype(np.int64)


def compute_deviations_core(counts_mat, annoptr, annoidx, background_peaks, expectation,
                            rowData=None, colData=None, rownames=None, colnames=None):
    """R/compute_deviations.R:355-408 with the bplapply fan-out replaced by one library call."""
    P, N = counts_mat.shape
    background_peaks = np.asarray(background_peaks)
    if background_peaks.ndim != 2 or background_peaks.shape[0] != P:
        raise ValueError("nrow(counts_mat) == nrow(background_peaks) is not TRUE")    # :365
    expectation = np.asarray(expectation, np.float64).ravel()
    if expectation.shape[0] != P:
        raise ValueError("length(expectation) == nrow(counts_mat) is not TRUE")       # :366
    annoidx = np.asarray(annoidx)
    if annoidx.size == 0 or annoidx.max() > P or annoidx.min() < 1:
        raise ValueError("Annotations are not valid")                                  # :373-379
    if not np.issubdtype(background_peaks.dtype, np.integer):
        bgf = background_peaks
        background_peaks = bgf.astype(np.int32)
        if not np.array_equal(background_peaks, bgf):
            raise ValueError("background_peaks must hold integer peak indices")
    try:
        eng = get_engine()
        resident = _is_resident(counts_mat, eng)
        if not resident and sp.isspmatrix_csc(counts_mat):
            # counts not on the GPU yet: ONE library call, the H2D of the counts, the kernels and the
            # D2H of the results overlap (cv_deviations_csc); the counts are resident afterwards
            invalidate_counts_cache()
            res = eng.deviations_csc(P, N, counts_mat.indptr, counts_mat.indices, counts_mat.data, annoptr,
                                     annoidx.astype(np.int32), background_peaks, expectation, index_base=1)
            _mark_resident(counts_mat)
        else:
            eng = _load_counts(counts_mat)
            res = eng.deviations(annoptr, annoidx.astype(np.int32), background_peaks, expectation, index_base=1)
    except ChromVARError as e:
        if e.code == _lib.CV_ERR_INVALID:
            raise ValueError(str(e)) from e
        raise
    K = len(annoptr) - 1
    rowData = dict(rowData or {})
    rowData["fractionMatches"] = res["matches"]                                       # :398-401
    rowData["fractionBackgroundOverlap"] = res["overlap"]
    rowData["variability"] = res["variability"]     # fused computeVariability SD (not in the reference object)
    if rownames is None:
        rownames = [str(i + 1) for i in range(K)]                                      # :381-383
    return chromVARDeviations(assays={"deviations": res["dev"], "z": res["z"]}, rowData=rowData,
                              colData=colData, rownames=rownames, colnames=colnames)


def computeDeviations(object, annotations=None, background_peaks=None, expectation=None):
    """All eight S4 methods of R/compute_deviations.R:224-353 behind one dispatcher."""
    is_se = isinstance(object, SummarizedExperiment)
    if not is_se and not _is_matrix(object):
        raise TypeError("object must be a SummarizedExperiment, a sparse matrix or an ndarray")
    counts_mat = _counts_matrix(object)
    colData = object.colData if is_se else None
    colnames = object.colnames if is_se else None
    rowData = rownames = None
    if annotations is None:                                                            # :284-296, :340-353
        print("Annotations not provided, so chromVAR being run on individual peaks...")
        P = counts_mat.shape[0]
        annoptr = np.arange(P + 1, dtype=np.int64)
        annoidx = np.arange(1, P + 1, dtype=np.int64)
    elif isinstance(annotations, SummarizedExperiment):                                # :224-238, :298-310
        annotations = matches_check(annotations)
        annoptr, annoidx = convert_to_ix_list(annotationMatches(annotations))
        rowData = annotations.colData
        rownames = annotations.colnames
    elif _is_matrix(annotations):                                                      # :243-258, :314-326
        if annotations.shape[0] != counts_mat.shape[0]:
            raise ValueError("Annotations are not valid")
        annoptr, annoidx = convert_to_ix_list(annotations)
    elif isinstance(annotations, dict):                                                # named list
        rownames = list(annotations.keys())
        annoptr, annoidx = _ix_list_to_csr(list(annotations.values()))
    elif isinstance(annotations, (list, tuple)):                                       # :263-276, :331-337
        annoptr, annoidx = _ix_list_to_csr(annotations)
    else:
        raise TypeError("unsupported annotations type")
    if background_peaks is None:
        if not is_se:
            raise TypeError('argument "background_peaks" is missing, with no default')
        background_peaks = getBackgroundPeaks(object)                                  # lazy default, :227
    if expectation is None:
        expectation = computeExpectations(object)                                      # lazy default, :228
    return compute_deviations_core(counts_mat, annoptr, annoidx, background_peaks, expectation,
                                   rowData=rowData, colData=colData, rownames=rownames, colnames=colnames)


# --------------------------------------------------------------------------------------------
# computeVariability
# --------------------------------------------------------------------------------------------
def _p_adjust_bh(p):
    """stats::p.adjust(method = "BH") (K scalars; stays on the host like it stays in R)."""
    p = np.asarray(p, np.float64)
    out = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pv = p[ok]
    n = pv.size
    if n:
        o = np.argsort(-pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        out[ok] = np.minimum(1.0, np.minimum.accumulate(n / np.arange(n, 0, -1) * pv[o]))[ro]
    return out


def _quantile7(x, probs, na_rm):
    """quantile_helper (R/compute_variability.R) = stats::quantile type 7 per column."""
    out = np.full((len(probs), x.shape[1]), np.nan)
    for j in range(x.shape[1]):
        col = x[:, j]
        if na_rm:
            col = col[~np.isnan(col)]
        elif np.isnan(col).any():
            continue
        if col.size:
            out[:, j] = np.quantile(col, probs)
    return out


def computeVariability(object, bootstrap_error=True, bootstrap_samples=1000,
                       bootstrap_quantiles=(0.025, 0.975), na_rm=True):
    """R/compute_variability.R:30-81.  Returns a dict of columns (the reference returns a
    data.frame): name, variability, [bootstrap_lower_bound, bootstrap_upper_bound,] p_value,
    p_value_adj."""
    from scipy.stats import chi2
    if not isinstance(object, chromVARDeviations):
        raise TypeError('inherits(object, "chromVARDeviations") is not TRUE')
    z = deviationScores(object)
    eng = get_engine()
    sd = eng.row_sds(z, na_rm)                                                        # :39
    ncol = z.shape[1]
    p_sd = chi2.sf((ncol - 1) * sd ** 2, ncol - 1)                                    # :41-43
    p_adj = _p_adjust_bh(p_sd)                                                        # :45
    names = object.rowData.get("name", object.rownames)
    out = {"name": names, "variability": sd}
    if bootstrap_error:                                                               # :50-63
        if not (bootstrap_samples > 0 and float(bootstrap_samples).is_integer()):
            raise ValueError("bootstrap_samples > 0 && all_whole(bootstrap_samples) is not TRUE")
        q = np.asarray(bootstrap_quantiles, np.float64)
        if not np.all(q > 0) or not np.all(q < 1) or not q[1] > q[0]:
            raise ValueError("bootstrap_quantiles must lie in (0, 1) and be increasing")
        boot = eng.row_sds_perm(z, na_rm, int(bootstrap_samples))
        err = _quantile7(boot, q, na_rm)
        out["bootstrap_lower_bound"] = err[0]
        out["bootstrap_upper_bound"] = err[1]
    out["p_value"] = p_sd
    out["p_value_adj"] = p_adj
    return out
