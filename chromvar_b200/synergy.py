"""Callers of the single-set kernel, batched.

The reference reaches `compute_deviations_single` once per derived peak set from
`getAnnotationSynergy` / `getAnnotationCorrelation` (R/variability_synergy.R:144-204, 341-399,
405-460; "very slow" by its own documentation).  Here every derived set of a call goes into ONE
library call on the resident counts: the sets become an annotation CSR, the fused row SD of z
(`out_variability`) is compute_variability_single (R/compute_variability.R:87-98), and only
K-vectors (or the K x N deviations for the correlations) leave the GPU.

Also here: deviationsCovariability (R/kmers.R:51-57) and getPermutedData
(R/background_peaks.R:178-203), which compose the same engine calls.
"""
import numpy as np
import scipy.sparse as sp

from . import api
from ._lib import ChromVARError


def _prep(object, annotations, background_peaks, expectation):
    """Argument handling shared by the eight S4 methods of each generic
    (R/variability_synergy.R:62-140, 252-339): returns (counts matrix, list of index vectors, names,
    background matrix, expectation)."""
    is_se = isinstance(object, api.SummarizedExperiment)
    if not is_se and not api._is_matrix(object):
        raise TypeError("object must be a SummarizedExperiment, a sparse matrix or an ndarray")
    counts_mat = api._counts_matrix(object)
    names = None
    if isinstance(annotations, api.SummarizedExperiment):
        annotations = api.matches_check(annotations)
        names = annotations.colnames
        ptr, idx = api.convert_to_ix_list(api.annotationMatches(annotations))
        sets = [idx[ptr[k]:ptr[k + 1]] for k in range(len(ptr) - 1)]
    elif api._is_matrix(annotations):
        ptr, idx = api.convert_to_ix_list(annotations)
        sets = [idx[ptr[k]:ptr[k + 1]] for k in range(len(ptr) - 1)]
    elif isinstance(annotations, dict):
        names = list(annotations.keys())
        sets = [np.asarray(v) for v in annotations.values()]
    elif isinstance(annotations, (list, tuple)):
        sets = [np.asarray(v) for v in annotations]
    else:
        raise TypeError("unsupported annotations type")
    if background_peaks is None:
        if not is_se:
            raise TypeError('argument "background_peaks" is missing, with no default')
        background_peaks = api.getBackgroundPeaks(object)
    if expectation is None:
        expectation = api.computeExpectations(object)
    P = counts_mat.shape[0]
    background_peaks = np.asarray(background_peaks)
    if background_peaks.shape[0] != P:
        raise ValueError("nrow(counts_mat) == nrow(background_peaks) is not TRUE")
    expectation = np.asarray(expectation, np.float64).ravel()
    if expectation.shape[0] != P:
        raise ValueError("length(expectation) == nrow(counts_mat) is not TRUE")
    flat = np.concatenate([np.asarray(s).ravel() for s in sets]) if sets else np.zeros(0)
    if flat.size == 0 or not np.array_equal(flat, np.round(flat)) or flat.max() > P or flat.min() < 1:
        raise ValueError("Annotations are not valid")
    if names is None:
        names = [str(i + 1) for i in range(len(sets))]
    if not np.issubdtype(background_peaks.dtype, np.integer):                 # as api.compute_deviations_core
        bgf = background_peaks
        with np.errstate(invalid="ignore"):
            background_peaks = bgf.astype(np.int32)
        if not np.array_equal(background_peaks, bgf):
            raise ValueError("background_peaks must hold integer peak indices")
    return counts_mat, [np.asarray(s).astype(np.int64) for s in sets], list(names), background_peaks, expectation


def _sets_to_csr(sets):
    ptr = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    idx = np.concatenate(sets).astype(np.int32) if ptr[-1] else np.zeros(0, np.int32)
    return ptr, idx


def _batch(counts_mat, sets, background_peaks, expectation, want_dev=False):
    """One library call for all `sets` on the resident counts: row SD of z per set (and the K x N
    deviations when asked).  Empty sets are legal here (all-NA rows, :458-461)."""
    eng = api._load_counts(counts_mat)
    ptr, idx = _sets_to_csr(sets)
    if ptr[-1] == 0:
        K = len(sets)
        return np.full(K, np.nan), (np.full((K, counts_mat.shape[1]), np.nan) if want_dev else None)
    try:
        eng.deviations_upload(ptr, idx, background_peaks, expectation, index_base=1)
        eng.deviations_compute()
        out = eng.deviations_download(vectors_only=not want_dev)
    except ChromVARError as e:
        raise ValueError(str(e)) from e
    return out["variability"], (out["dev"] if want_dev else None)


def compute_variability_batch(counts_mat, sets, background_peaks, expectation):
    """compute_variability_single (R/compute_variability.R:87-98) for a list of peak sets."""
    return _batch(counts_mat, [np.asarray(s) for s in sets], background_peaks, expectation)[0]


def _variabilities_arg(variabilities, counts_mat, sets, background_peaks, expectation):
    if variabilities is None:                                                   # :172-177, :366-374
        return _batch(counts_mat, sets, background_peaks, expectation)[0]
    if isinstance(variabilities, dict) and "variability" in variabilities:     # data.frame of computeVariability
        return np.asarray(variabilities["variability"], np.float64)
    raise ValueError("Incorrect variabilities input")


def getAnnotationSynergy(object, annotations, background_peaks=None, expectation=None, variabilities=None,
                         nbg=25, rng=None):
    """R/variability_synergy.R:62-204.  Returns (l x l matrix, names in decreasing-variability order)
    -- the reference returns the matrix with those dimnames.  The `sample(peak_set, x)` draws of
    get_variability_boost_helper (:421-427) come from `rng` (numpy Generator) in the reference's
    order; all (l - i)(1 + nbg) derived sets of every i go into one library call."""
    counts_mat, sets, names, bg, e = _prep(object, annotations, background_peaks, expectation)
    rng = rng or np.random.default_rng()
    var = _variabilities_arg(variabilities, counts_mat, sets, bg, e)
    l = len(sets)
    if var.shape[0] != l:
        raise ValueError("length(variabilities) == length(peak_indices) is not TRUE")
    if l < 2:
        raise ValueError("l >= 2 is not TRUE")
    order = np.argsort(-var, kind="stable")                                     # :190
    derived, layout = [], []
    for i in range(l - 1):                                                      # :191-199
        peak_set = sets[order[i]]
        tmpsets = [s[np.isin(s, peak_set)] for s in (sets[k] for k in order[i + 1:])]      # remove_nonoverlap
        bgsets = [rng.choice(peak_set, len(s), replace=False) for s in tmpsets for _ in range(nbg)]
        layout.append((len(derived), len(tmpsets)))
        derived.extend(tmpsets)
        derived.extend(bgsets)
    v = _batch(counts_mat, derived, bg, e)[0]
    out = np.full((l, l), np.nan)
    for i, (start, n) in enumerate(layout):
        tmpvar = v[start:start + n]
        bgvar = v[start + n:start + n + n * nbg].reshape(n, nbg)
        with np.errstate(invalid="ignore", divide="ignore"):
            boost = (tmpvar - bgvar.mean(axis=1)) / bgvar.std(axis=1, ddof=1)   # :435-441
        out[i, i + 1:] = boost
        out[i + 1:, i] = boost
    np.fill_diagonal(out, 0.0)                                                  # :200
    return out, [names[k] for k in order]


def getAnnotationCorrelation(object, annotations, background_peaks=None, expectation=None, variabilities=None):
    """R/variability_synergy.R:252-399: cor of the bias-corrected deviations of A_i \\ A_j and
    A_j \\ A_i for every pair (get_cor_helper, :447-460); all l(l-1) difference sets in one call."""
    counts_mat, sets, names, bg, e = _prep(object, annotations, background_peaks, expectation)
    var = _variabilities_arg(variabilities, counts_mat, sets, bg, e)
    l = len(sets)
    if var.shape[0] != l:
        raise ValueError("length(variabilities) == length(peak_indices) is not TRUE")
    if l < 2:
        raise ValueError("l >= 2 is not TRUE")
    order = np.argsort(-var, kind="stable")
    s = [sets[k] for k in order]
    derived, pairs = [], []
    for i in range(l):
        for j in range(i + 1, l):
            pairs.append((i, j))
            derived.append(s[i][~np.isin(s[i], s[j])])
            derived.append(s[j][~np.isin(s[j], s[i])])
    dev = _batch(counts_mat, derived, bg, e, want_dev=True)[1]
    out = np.ones((l, l))
    for n, (i, j) in enumerate(pairs):
        d1, d2 = dev[2 * n], dev[2 * n + 1]
        if np.isnan(d1).any() or np.isnan(d2).any():
            c = np.nan                                                          # cor(use = "everything")
        else:
            with np.errstate(invalid="ignore", divide="ignore"):
                c = np.corrcoef(d1, d2)[0, 1]
        out[i, j] = out[j, i] = c
    return out, [names[k] for k in order]


def deviationsCovariability(object):
    """R/kmers.R:51-57: cov(t(z)) with row i divided by var_i (row_sds on the GPU)."""
    if not isinstance(object, api.chromVARDeviations):
        raise TypeError('inherits(object, "chromVARDeviations") is not TRUE')
    # the K x K covariance runs on the device (cv_deviations_covariability): z of a k-mer run is
    # 2080 x 50 000, its covariance 4e11 multiply-adds
    return api.get_engine().deviations_covariability(api.deviationScores(object))


def _bg_for_perm(object, n, w, bs):
    """getBackgroundPeaks(object, niterations = ncol(object)) with the engine seed already set."""
    mat = api._counts_matrix(object)
    bias = np.asarray(object.rowData["bias"], np.float64)
    eng = api._load_counts(mat)
    try:
        return eng.background_peaks(bias, int(n), float(w), int(bs))
    except ChromVARError as e:
        raise ValueError(str(e)) from e


def getPermutedData(object, niterations=10, w=0.1, bs=50):
    """R/background_peaks.R:178-203: adds assays perm_1 .. perm_n; column x of a permuted assay is
    counts[bg[, x], x] with bg = getBackgroundPeaks(object, niterations = ncol(object))."""
    if not isinstance(object, api.SummarizedExperiment):
        raise TypeError("object must be a SummarizedExperiment")
    mat = api.counts_check(object).assays["counts"]
    bias = object.rowData.get("bias")
    if bias is None:
        raise ValueError("bias is missing (rowData(object)$bias)")
    eng = api._load_counts(mat)
    assays = dict(object.assays)
    for it in range(int(niterations)):
        eng.set_seed(api._next_sampler_seed())                   # a fresh draw per iteration (:184-186)
        # draws and row gather on the device (cv_permuted_counts): the P x N index matrix never exists
        try:
            colptr, rowidx, x = eng.permuted_counts(np.asarray(bias, np.float64), float(w), int(bs))
        except ChromVARError as e:
            raise ValueError(str(e)) from e
        M = sp.csc_matrix((x, rowidx, colptr), shape=mat.shape)
        M.has_sorted_indices = True
        assays["perm_%d" % (it + 1)] = M
    return api.SummarizedExperiment(assays, rowData=object.rowData, colData=object.colData,
                                    rownames=object.rownames, colnames=object.colnames)


# --------------------------------------------------------------------------------------------
# makePermutedSets (R/test_sets.R:150-289) on get_background_peaks_alternative (:291-330)
# --------------------------------------------------------------------------------------------
def getBackgroundPeaksAlternative(object, bias=None, niterations=50, window=500):
    """get_background_peaks_alternative (R/test_sets.R:291-330; internal in the reference): P x
    niterations 1-based peaks, each drawn uniformly among the `window` nearest other peaks in the
    whitened (fragments per peak, bias) plane.  One kernel on the GPU instead of the bplapply over
    groups of 2000 peaks around nabor::knn."""
    if isinstance(object, api.SummarizedExperiment) and bias is None:
        bias = object.rowData.get("bias")
    mat = api._counts_matrix(object)
    if bias is None:
        raise ValueError("bias is missing (rowData(object)$bias)")
    bias = np.asarray(bias, np.float64)
    if bias.shape[0] != mat.shape[0]:
        raise ValueError("length(bias) must equal nrow(object)")
    eng = api._load_counts(mat)
    eng.set_seed(api._next_sampler_seed())
    try:
        return eng.background_peaks_knn(bias, int(niterations), int(window))
    except ChromVARError as e:
        raise ValueError(str(e)) from e


def makePermutedSets(object, annotations, bias=None, window=10):
    """R/test_sets.R:150-289: every annotation's peaks replaced by a draw among their `window`
    nearest neighbours in (fragments per peak, bias); returns a SummarizedExperiment with the
    assay `matches` (P x K boolean sparse, convert_from_ix_list: a peak hit twice counts once) and
    colData$name = "permuted.<name>".  All eight method signatures: object = SummarizedExperiment /
    matrix, annotations = SummarizedExperiment / matrix / list (dict keeps the names)."""
    is_se = isinstance(object, api.SummarizedExperiment)
    if not is_se and not api._is_matrix(object):
        raise TypeError("object must be a SummarizedExperiment, a sparse matrix or an ndarray")
    if is_se:
        object = api.counts_check(object)
    col_data = None
    if isinstance(annotations, api.SummarizedExperiment):
        annotations = api.matches_check(annotations)
        ptr, idx = api.convert_to_ix_list(api.annotationMatches(annotations))
        sets = [idx[ptr[k]:ptr[k + 1]] for k in range(len(ptr) - 1)]
        names = annotations.colnames
        col_data = dict(annotations.colData)
    elif api._is_matrix(annotations):
        ptr, idx = api.convert_to_ix_list(annotations)
        sets = [idx[ptr[k]:ptr[k + 1]] for k in range(len(ptr) - 1)]
        names = None
    elif isinstance(annotations, dict):
        names = list(annotations.keys())
        sets = [np.asarray(v) for v in annotations.values()]
    elif isinstance(annotations, (list, tuple)):
        sets, names = [np.asarray(v) for v in annotations], None
    else:
        raise TypeError("unsupported annotations type")
    P = api._counts_matrix(object).shape[0]
    bg = getBackgroundPeaksAlternative(object, bias, niterations=1, window=window)[:, 0]     # :275-276
    perm = [bg[np.asarray(s).astype(np.int64) - 1].astype(np.int64) for s in sets]            # :277
    # names(sets) <- paste("permuted.", name): unnamed lists give "permuted." for every set (:278-281)
    pnames = ["permuted.%s" % (n if n is not None else "") for n in (names if names is not None else [None] * len(sets))]
    rows = np.concatenate(perm) - 1 if perm else np.zeros(0, np.int64)
    cols = np.concatenate([np.full(len(v), k, np.int64) for k, v in enumerate(perm)]) if perm else np.zeros(0, np.int64)
    M = sp.csc_matrix((np.ones(rows.size, bool), (rows, cols)), shape=(P, len(perm)), dtype=bool)
    M.data[:] = True                                                                          # duplicates: TRUE
    if col_data is None:
        col_data = {"name": np.asarray(pnames)}                                               # :283-284
    elif "name" in col_data:
        col_data["name"] = np.asarray(["permuted.%s" % n for n in col_data["name"]])          # :286-290
    else:
        col_data["name"] = np.asarray(pnames)
    return api.SummarizedExperiment({"matches": M}, colData=col_data, colnames=pnames)
