# R shim: replacement bodies for the four attachment points of chromVAR's deviation engine.
# Signatures, validation and stop() messages are the reference's; only the arithmetic moves to
# libchromvar_b200.so through .Call.  Written without an R interpreter at hand (none in the build
# image) -- see INTEGRATION.md.
#
#   compute_deviations_core    replaces R/compute_deviations.R:355-408 (bplapply fan-out + vapply)
#   compute_expectations_core  replaces R/compute_deviations.R:415-448
#   get_background_peaks_core  replaces R/background_peaks.R:116-156
#   row_sds / row_sds_perm     replace the .Call wrappers in R/RcppExports.R:20-26

.cvb <- new.env()

cvb_init <- function(device = 0L) {
  seed <- sample.int(.Machine$integer.max, 1)           # set.seed() keeps governing the sampler
  .Call("cvb_init", as.integer(device), as.numeric(seed), PACKAGE = "chromVARb200")
  .cvb$counts <- NULL
  invisible(TRUE)
}

.cvb_load_counts <- function(counts_mat) {
  if (!is.null(.cvb$counts) && identical(.cvb$counts, counts_mat)) return(invisible(FALSE))
  if (is(counts_mat, "dgCMatrix")) {
    .Call("cvb_set_counts_csc", counts_mat@p, counts_mat@i, counts_mat@x, counts_mat@Dim,
          PACKAGE = "chromVARb200")
  } else if (is(counts_mat, "Matrix")) {
    .cvb_load_counts(as(counts_mat, "dgCMatrix")); .cvb$counts <- counts_mat; return(invisible(TRUE))
  } else {
    .Call("cvb_set_counts_dense", as.matrix(counts_mat), PACKAGE = "chromVARb200")
  }
  .cvb$counts <- counts_mat
  invisible(TRUE)
}

compute_expectations_core <- function(object, norm = FALSE, group = NULL) {
  .cvb_load_counts(object)
  codes <- if (is.null(group)) NULL else as.integer(group)
  .Call("cvb_expectations", nrow(object), isTRUE(norm), codes,
        if (is.null(group)) 0L else length(levels(group)), PACKAGE = "chromVARb200")
}

get_background_peaks_core <- function(object, bias, niterations = 50, w = 0.1, bs = 50) {
  stopifnot(length(bias) == nrow(object))
  .cvb_load_counts(if (is(object, "SummarizedExperiment")) counts(object) else object)
  .Call("cvb_set_seed", as.numeric(sample.int(.Machine$integer.max, 1)), PACKAGE = "chromVARb200")
  # the library stop()s with "All peaks must have at least one fragment in one sample"
  .Call("cvb_background_peaks", as.numeric(bias), as.integer(niterations), as.numeric(w),
        as.integer(bs), PACKAGE = "chromVARb200")
}

# replaces R/test_sets.R:291-330 (the bplapply over groups of 2000 peaks around nabor::knn);
# makePermutedSets_core (R/test_sets.R:270-289) calls it unchanged
get_background_peaks_alternative <- function(object, bias, niterations = 50, window = 500) {
  fragments_per_peak <- getFragmentsPerPeak(object)
  if (min(fragments_per_peak) <= 0)
    stop("All peaks must have at least one fragment in one sample")
  .cvb_load_counts(if (is(object, "SummarizedExperiment")) counts(object) else object)
  .Call("cvb_set_seed", as.numeric(sample.int(.Machine$integer.max, 1)), PACKAGE = "chromVARb200")
  .Call("cvb_background_peaks_knn", as.numeric(bias), as.integer(niterations), as.integer(window),
        PACKAGE = "chromVARb200")
}

compute_deviations_core <- function(counts_mat, peak_indices, background_peaks, expectation,
                                    rowData = NULL, colData = NULL) {
  stopifnot(nrow(counts_mat) == nrow(background_peaks))
  stopifnot(length(expectation) == nrow(counts_mat))
  if (is.null(colData))
    colData <- DataFrame(seq_len(ncol(counts_mat)), row.names = colnames(counts_mat))[, FALSE]
  tmp <- unlist(peak_indices, use.names = FALSE)
  if (is.null(tmp) || !(all.equal(tmp, as.integer(tmp))) || max(tmp) > nrow(counts_mat) || min(tmp) < 1)
    stop("Annotations are not valid")
  if (is.null(names(peak_indices))) names(peak_indices) <- seq_along(peak_indices)
  annoptr <- c(0, cumsum(lengths(peak_indices)))
  storage.mode(background_peaks) <- "integer"
  fresh <- is.null(.cvb$counts) || !identical(.cvb$counts, counts_mat)
  if (fresh && is(counts_mat, "dgCMatrix")) {
    # counts not on the GPU yet: one call, H2D of the counts / kernels / D2H of the results overlap
    res <- .Call("cvb_deviations_csc", counts_mat@p, counts_mat@i, counts_mat@x, counts_mat@Dim,
                 as.numeric(annoptr), as.integer(tmp), background_peaks, as.numeric(expectation),
                 PACKAGE = "chromVARb200")
    .cvb$counts <- counts_mat
  } else {
    .cvb_load_counts(counts_mat)
    res <- .Call("cvb_deviations", as.numeric(annoptr), as.integer(tmp), background_peaks,
                 as.numeric(expectation), ncol(counts_mat), PACKAGE = "chromVARb200")
  }
  dev <- res[[1]]; z <- res[[2]]
  rownames(z) <- rownames(dev) <- names(peak_indices)
  colnames(z) <- colnames(dev) <- colnames(counts_mat)
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
