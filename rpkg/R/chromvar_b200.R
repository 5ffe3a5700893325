# R shim: replacement bodies for the attachment points of chromVAR's deviation engine.
# Signatures, validation and stop() messages are the reference's; only the arithmetic moves to
# libchromvar_b200.so through .Call (rpkg/src/cv_rglue.c).  No R interpreter exists in the build
# image: the C side of every .Call below is compiled and driven through its registered table by
# tests/test_gpu_rglue.py (same argument types and order as here); this file itself is checked
# only statically (tests/test_rglue_cpu.py: every .Call names a registered routine with the right
# number of arguments).  See INTEGRATION.md.
#
#   compute_deviations_core    replaces R/compute_deviations.R:355-408 (bplapply fan-out + vapply)
#   compute_expectations_core  replaces R/compute_deviations.R:415-448
#   get_background_peaks_core  replaces R/background_peaks.R:116-156
#   row_sds / row_sds_perm     replace the .Call wrappers in R/RcppExports.R:20-26
#
# counts_mat may be ONE matrix (dgCMatrix / Matrix / base matrix) or a LIST of dgCMatrix column
# blocks with equal nrow (an atlas whose nonzeros exceed 2^31 - 1 cannot be one dgCMatrix: @p is
# integer).  cvb_init(devices = 0:7) lets one R process drive eight GPUs: the library shards the
# cells over the devices and reduces per-peak totals / variability partials itself.

.cvb <- new.env()

cvb_init <- function(devices = 0L) {
  seed <- sample.int(.Machine$integer.max, 1)           # set.seed() keeps governing the sampler
  .Call("cvb_init", as.integer(devices), as.numeric(seed), PACKAGE = "chromVARb200")
  .cvb$counts <- NULL
  .cvb$n_devices <- length(devices)
  invisible(TRUE)
}

.cvb_blocks <- function(counts_mat) {
  # list of dgCMatrix column blocks (one element for an ordinary matrix)
  if (is.list(counts_mat)) {
    blocks <- lapply(counts_mat, function(b) if (is(b, "dgCMatrix")) b else as(b, "dgCMatrix"))
    stopifnot(length(unique(vapply(blocks, nrow, 0L))) == 1)
    blocks
  } else if (is(counts_mat, "dgCMatrix")) {
    list(counts_mat)
  } else {
    list(as(counts_mat, "dgCMatrix"))
  }
}

.cvb_nrow <- function(counts_mat) if (is.list(counts_mat)) nrow(counts_mat[[1]]) else nrow(counts_mat)
.cvb_ncol <- function(counts_mat) if (is.list(counts_mat)) sum(vapply(counts_mat, ncol, 0L)) else ncol(counts_mat)
.cvb_colnames <- function(counts_mat)
  if (is.list(counts_mat)) unlist(lapply(counts_mat, colnames), use.names = FALSE) else colnames(counts_mat)

.cvb_load_counts <- function(counts_mat) {
  # identical() returns at once for the same object (pointer-equal slots), so the usual sequence
  # getBackgroundPeaks(x) -> computeExpectations(x) -> computeDeviations(x, ...) uploads once
  if (!is.null(.cvb$counts) && identical(.cvb$counts, counts_mat)) return(invisible(FALSE))
  if (!is.list(counts_mat) && !is(counts_mat, "Matrix") && .cvb$n_devices == 1L) {
    .Call("cvb_set_counts_dense", as.matrix(counts_mat), PACKAGE = "chromVARb200")
  } else {
    b <- .cvb_blocks(counts_mat)
    .Call("cvb_set_counts", lapply(b, slot, "p"), lapply(b, slot, "i"), lapply(b, slot, "x"),
          nrow(b[[1]]), PACKAGE = "chromVARb200")
  }
  .cvb$counts <- counts_mat
  invisible(TRUE)
}

.cvb_fragments <- function(counts_mat) {
  .cvb_load_counts(counts_mat)
  .Call("cvb_fragments", .cvb_nrow(counts_mat), .cvb_ncol(counts_mat), PACKAGE = "chromVARb200")
}

compute_expectations_core <- function(object, norm = FALSE, group = NULL) {
  .cvb_load_counts(object)
  codes <- if (is.null(group)) NULL else as.integer(as.factor(group))
  .Call("cvb_expectations", .cvb_nrow(object), isTRUE(norm), codes,
        if (is.null(group)) 0L else nlevels(as.factor(group)), PACKAGE = "chromVARb200")
}

get_background_peaks_core <- function(object, bias, niterations = 50, w = 0.1, bs = 50) {
  counts_mat <- if (is(object, "SummarizedExperiment")) counts(object) else object
  fragments_per_peak <- .cvb_fragments(counts_mat)[[1]]
  if (min(fragments_per_peak) <= 0)                                  # R/background_peaks.R:119-121
    stop("All peaks must have at least one fragment in one sample")
  stopifnot(length(bias) == length(fragments_per_peak))              # :123
  .Call("cvb_set_seed", as.numeric(sample.int(.Machine$integer.max, 1)), PACKAGE = "chromVARb200")
  .Call("cvb_background_peaks", as.numeric(bias), as.integer(niterations), as.numeric(w),
        as.integer(bs), PACKAGE = "chromVARb200")
}

# replaces R/test_sets.R:291-330 (the bplapply over groups of 2000 peaks around nabor::knn);
# makePermutedSets_core (R/test_sets.R:270-289) calls it unchanged
get_background_peaks_alternative <- function(object, bias, niterations = 50, window = 500) {
  counts_mat <- if (is(object, "SummarizedExperiment")) counts(object) else object
  fragments_per_peak <- .cvb_fragments(counts_mat)[[1]]
  if (min(fragments_per_peak) <= 0)
    stop("All peaks must have at least one fragment in one sample")
  .Call("cvb_set_seed", as.numeric(sample.int(.Machine$integer.max, 1)), PACKAGE = "chromVARb200")
  .Call("cvb_background_peaks_knn", as.numeric(bias), as.integer(niterations), as.integer(window),
        PACKAGE = "chromVARb200")
}

compute_deviations_core <- function(counts_mat, peak_indices, background_peaks, expectation,
                                    rowData = NULL, colData = NULL) {
  P <- .cvb_nrow(counts_mat)
  # The reference checks, in this order (R/compute_deviations.R:362-379): every peak has a
  # fragment; nrow(background_peaks); length(expectation); the annotations.  The library reports the
  # first one itself, after the kernels ran, so the cheap R-level checks are evaluated first WITHOUT
  # stopping; only if one of them fails is the zero-peak check forced ahead of it (per-peak totals
  # from the GPU), which keeps the reference's precedence at no cost on the good path.
  tmp <- unlist(peak_indices, use.names = FALSE)
  later <- NULL
  if (P != nrow(background_peaks)) {
    later <- "nrow(counts_mat) == nrow(background_peaks) is not TRUE"
  } else if (length(expectation) != P) {
    later <- "length(expectation) == nrow(counts_mat) is not TRUE"
  } else if (is.null(tmp) || !isTRUE(all.equal(tmp, as.integer(tmp))) || max(tmp) > P || min(tmp) < 1) {
    later <- "Annotations are not valid"
  }
  if (!is.null(later)) {
    if (min(.cvb_fragments(counts_mat)[[1]]) <= 0)
      stop("All peaks must have at least one fragment in one sample")
    stop(later)
  }
  if (is.null(colData))
    colData <- DataFrame(seq_len(.cvb_ncol(counts_mat)), row.names = .cvb_colnames(counts_mat))[, FALSE]
  if (is.null(names(peak_indices))) names(peak_indices) <- seq_along(peak_indices)
  annoptr <- c(0, cumsum(lengths(peak_indices)))
  fresh <- is.null(.cvb$counts) || !identical(.cvb$counts, counts_mat)
  sparse <- is.list(counts_mat) || is(counts_mat, "Matrix") || .cvb$n_devices > 1L
  if (fresh && sparse) {
    # counts not on the GPU yet: one call, H2D of the counts / kernels / D2H of the results overlap
    b <- .cvb_blocks(counts_mat)
    res <- .Call("cvb_deviations_csc", lapply(b, slot, "p"), lapply(b, slot, "i"), lapply(b, slot, "x"),
                 P, as.numeric(annoptr), as.integer(tmp), background_peaks, as.numeric(expectation),
                 PACKAGE = "chromVARb200")
    .cvb$counts <- counts_mat
  } else {
    .cvb_load_counts(counts_mat)
    res <- .Call("cvb_deviations", as.numeric(annoptr), as.integer(tmp), background_peaks,
                 as.numeric(expectation), .cvb_ncol(counts_mat), PACKAGE = "chromVARb200")
  }
  dev <- res[[1]]; z <- res[[2]]
  rownames(z) <- rownames(dev) <- names(peak_indices)
  colnames(z) <- colnames(dev) <- .cvb_colnames(counts_mat)
  rowData$fractionMatches <- res[[3]]
  rowData$fractionBackgroundOverlap <- res[[4]]
  out <- SummarizedExperiment(assays = list(deviations = dev, z = z), colData = colData, rowData = rowData)
  new("chromVARDeviations", out)
}

row_sds <- function(x, na_rm = FALSE) .Call("cvb_row_sds", x, na_rm, PACKAGE = "chromVARb200")

# one call for all bootstrap replicates (the reference loops bplapply over row_sds_perm,
# R/compute_variability.R:57-60); returns the bootstrap_samples x K matrix boot_sd
row_sds_perm_batch <- function(x, na_rm = FALSE, n_rep = 1000L)
  .Call("cvb_row_sds_perm", x, na_rm, as.integer(n_rep), PACKAGE = "chromVARb200")
