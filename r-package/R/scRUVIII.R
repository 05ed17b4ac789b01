## Drop-in bodies for the remaining RUV-III closures of scMerge (same signatures as R/scRUVIII.R:41-44,
## R/scMerge2_utilis.R:17-26, R/scMerge2.R:368-372).  As in fastRUVIII.R only argument resolution lives in R; the
## arithmetic is behind the .Call entry points of src/rshim.c.  NOT executed in this repository (no R in the image): the
## tested twin of every function below is scmerge_b200/ruv.py, which drives the same C ABI in the same order.
##
## Helpers shared with fastRUVIII.R: .keys() (replicate structure -> 0-based int keys), .rows_cap().

.ctl_mask <- function(ctl, n) { out <- rep(FALSE, n); out[ctl] <- TRUE; out }   # tological(), R/fastRUVIII.R:113-117

## scRUVIII: standardize2 + fastRUVIII per k + de-standardise + choice of ruvK (R/scRUVIII.R:41-137).
## Y is genes x cells, which is byte-identical to the device layout (row-major cells x genes): uploaded once, never rewritten.
## The standardisation is folded into the fit (batch keys) and into the adjust coefficients (center / pre / post).
scRUVIII <- function(Y = Y, M = M, ctl = ctl, fullalpha = NULL, k = k, cell_type = NULL, batch = NULL,
                     return_all_RUV = TRUE, BPPARAM = SerialParam(), BSPARAM = ExactParam(), svd_k = 50) {
  n <- nrow(Y); m <- ncol(Y)
  g <- .keys(M); b <- .keys(batch); ctl2 <- .ctl_mask(ctl, n)
  exact <- inherits(BSPARAM, "ExactParam")
  s <- if (exact) min(m - g$G, sum(ctl2), svd_k, na.rm = TRUE) else min(m - g$G, sum(ctl2), na.rm = TRUE)
  if (g$G >= m | k[1] == 0) {                          # fastRUVIII's early-out, then mean and sd are added back: the input
    res <- lapply(k, function(kk) list(newY = t(as.matrix(Y)), M = M, fullalpha = NULL))
    names(res) <- k
    if (return_all_RUV) { res$optimal_ruvK <- k[1]; return(res) }
    one <- res[[1]]; one$optimal_ruvK <- k[1]; return(one)
  }
  cap <- .rows_cap(s, max(k), exact, n)
  select <- length(k) > 1
  if (!select && is.null(fullalpha) && is.matrix(Y) && is.double(Y)) {
    ## a plain R matrix and one ruvK: the whole closure in ONE call on REAL(Y) itself (scm_ruv3_host / scm_multi_ruv3_host:
    ## upload, fit, adjust and download overlapped; all GPUs of useGPUs()).  genes x cells column-major IS the device layout.
    r <- .Call("scm_host", Y, g$key, g$G, b$key, b$G, ctl2, as.integer(k[1]), cap$s, if (cap$exact) 0L else 1L, NULL, 0L, NULL)
    fa <- t(r[[2]]); colnames(fa) <- rownames(Y); attr(fa, "conv_rank") <- r[[6]]
    newY <- t(r[[1]]); dimnames(newY) <- rev(dimnames(Y))
    one <- list(newY = newY, M = M, fullalpha = fa)
    if (return_all_RUV) { res <- list(one); names(res) <- k[1]; res$optimal_ruvK <- k[1]; return(res) }
    one$optimal_ruvK <- k[1]; return(one)
  }
  dY <- .Call("scm_upload", as.matrix(Y), 0L)
  ## one fit serves every k (R/scRUVIII.R:64-75); scm_fit_gram also keeps the centred Gram for the silhouette PCA
  mode <- if (cap$exact) 0L else 1L
  fit <- if (select) .Call("scm_fit_gram", dY, g$key, g$G, b$key, b$G, cap$s, min(k[1], cap$s), mode, as.integer(sort(unique(k[-1]))))
         else .Call("scm_fit", dY, g$key, g$G, b$key, b$G, cap$s, min(k[1], cap$s), mode)
  if (is.null(fullalpha)) { fullalpha <- t(fit[[1]]); colnames(fullalpha) <- rownames(Y) }
  stand_mean <- fit[[3]]; stand_sd <- fit[[4]]
  adjust_one <- function(kk) {
    kk <- min(kk, nrow(fullalpha))
    out <- .Call("scm_adjust_", dY, t(fullalpha[seq_len(kk), , drop = FALSE]), kk, ctl2, stand_mean, 1 / stand_sd, stand_sd, NULL)
    newY <- t(out); dimnames(newY) <- rev(dimnames(Y))  # cells x genes, like the reference's t(t(newY) * sd + mean)
    list(newY = newY, M = M, fullalpha = fullalpha)
  }
  f_score <- 1; names(f_score) <- k[1]
  if (select) {
    if (is.null(cell_type)) ct <- g else ct <- .keys(cell_type)          # R/scRUVIII.R:89-92
    sil <- vapply(k, function(kk) {
      kk <- min(kk, nrow(fullalpha))
      .Call("scm_sil", dY, fit$gram, fit$colmean, t(fullalpha[seq_len(kk), , drop = FALSE]), kk, ctl2, stand_mean, 1 / stand_sd,
            stand_sd, ct$key, ct$G, b$key, b$G, 10L)
    }, numeric(2))
    zo <- (sil + 1) / 2                                                   # zeroOneScale, R/scRUVIII.R:186-190
    f_score <- 2 * zo[1, ] * (1 - zo[2, ]) / (zo[1, ] + (1 - zo[2, ]))    # f_measure, R/scRUVIII.R:177-180
    names(f_score) <- k
    message("optimal ruvK:", k[which.max(f_score)])
  }
  best <- which.max(f_score)
  if (return_all_RUV) {
    res <- lapply(k, adjust_one); names(res) <- k
    res$optimal_ruvK <- k[best]
    return(res)
  }
  one <- adjust_one(k[best]); one$optimal_ruvK <- k[best]
  one
}

## pseudoRUVIII: fit on the pseudobulk matrix, adjust every cell (R/scMerge2_utilis.R:17-150).  Y / Y_pseudo are cells x genes.
## byChunk / chunkSize / BPPARAM are accepted and unused: the fused adjust kernel streams all cells in one launch.
pseudoRUVIII <- function(Y, Y_pseudo, M, ctl, k = NULL, eta = NULL, include.intercept = TRUE, inputcheck = TRUE,
                         fullalpha = NULL, return.info = FALSE, subset = NULL, BPPARAM = BiocParallel::SerialParam(),
                         BSPARAM = BiocSingular::ExactParam(), normalised = TRUE, verbose = TRUE, byChunk = TRUE,
                         chunkSize = 50000) {
  n <- ncol(Y); P <- nrow(Y_pseudo)
  g <- .keys(M); ctl2 <- .ctl_mask(ctl, n)
  exact <- inherits(BSPARAM, "ExactParam")
  s <- if (exact) min(P - g$G, sum(ctl2), k, na.rm = TRUE) else min(P - g$G, sum(ctl2), na.rm = TRUE)   # :60-65 (cap is k here)
  if (g$G >= P | k == 0) {                                                                                # :72-76 (after RUV1, :56)
    if (is.null(eta)) return(Y_pseudo)
    E <- .eta_rows(eta, n, include.intercept)
    out <- t(.Call("scm_adjust_", .Call("scm_upload", as.matrix(Y_pseudo), 1L), t(E), nrow(E), ctl2, NULL, NULL, NULL, NULL))
    dimnames(out) <- dimnames(Y_pseudo)
    return(out)
  }
  if (is.null(fullalpha)) {
    dP <- .Call("scm_upload", as.matrix(Y_pseudo), 1L)
    if (!is.null(eta)) dP <- .ruv1(dP, eta, ctl2, include.intercept)                                      # :56: the pseudobulks only
    cap <- .rows_cap(s, k, exact, n, rows = P)
    fit <- .Call("scm_fit", dP, g$key, g$G, NULL, 0L, cap$s, min(k, cap$s), if (cap$exact) 0L else 1L)
    fullalpha <- t(fit[[1]]); colnames(fullalpha) <- colnames(Y)
  }
  if (!normalised) return(list(M = M, fullalpha = fullalpha))
  if (is.null(sub) && methods::is(Y, "dgCMatrix")) {
    ## sparse-native: the three slots of t(Y) (genes x cells, compressed by cell) are the only copy of the cells on the device;
    ## column means come from one pseudobulk pass with a single key, newY is streamed straight into the R matrix
    Yt <- Matrix::t(Y)
    cs <- .Call("scm_upload_csc", Yt@p, Yt@i, Yt@x, dim(Yt), FALSE)
    means <- .Call("scm_csc_pseudobulk", cs, integer(nrow(Y)), 1L)[[3]]
    kk <- min(k, nrow(fullalpha))
    out <- .Call("scm_csc_adjust_", cs, t(fullalpha[seq_len(kk), , drop = FALSE]), kk, ctl2, means)
    newY <- DelayedArray::DelayedArray(t(out)); dimnames(newY) <- dimnames(Y)
    return(if (!return.info) newY else list(newY = newY, M = M, fullalpha = fullalpha))
  }
  ## a dgCMatrix goes up as its three slots (genes x cells expected by scm_upload_sparse: pass t(Y) when Y is cells x genes)
  dY <- if (methods::is(Y, "dgCMatrix")) { Yt <- Matrix::t(Y); .Call("scm_upload_sparse", Yt@p, Yt@i, Yt@x, dim(Yt), FALSE) }
        else .Call("scm_upload", as.matrix(Y), 1L)
  means <- .Call("scm_segmean", dY, integer(nrow(Y)), 1L)[[1]][, 1]       # adjusted_means <- colMeans(Y), :97
  kk <- min(k, nrow(fullalpha))
  ## Y[, subset] of the reference (R/scMerge2_utilis.R:117-121): names are matched against colnames(Y) IN THE GIVEN ORDER
  ## (scMerge2 passes chosen.hvg), integers are positions, a logical vector is a mask; anything not in Y is an error as in R
  sub <- if (is.null(subset)) NULL else {
    pos <- if (is.character(subset)) match(subset, colnames(Y)) else if (is.logical(subset)) which(subset) else as.integer(subset)
    if (anyNA(pos) || any(pos < 1L) || any(pos > n)) stop("subscript out of bounds: `subset` names genes that are not columns of Y")
    as.integer(pos - 1L)
  }
  out <- .Call("scm_adjust_", dY, t(fullalpha[seq_len(kk), , drop = FALSE]), kk, ctl2, means, NULL, NULL, sub)
  newY <- DelayedArray::DelayedArray(t(out))
  rownames(newY) <- rownames(Y); colnames(newY) <- if (is.null(sub)) colnames(Y) else colnames(Y)[sub + 1L]
  if (!return.info) newY else list(newY = newY, M = M, fullalpha = fullalpha)
}

## getAdjustedMat: adjust-only from a stored fullalpha (R/scMerge2.R:368-410).  exprsMat is genes x cells; ctl and
## return_subset_genes are matched by NAME and keep the matrix order, like the reference's intersect(rownames(.), .).
getAdjustedMat <- function(exprsMat, fullalpha, ctl = rownames(exprsMat), adjusted_means = NULL, ruvK = 20,
                           return_subset_genes = NULL) {
  ctl2 <- rownames(exprsMat) %in% ctl
  if (!any(ctl2)) stop("No provided ctl genes in the data")
  sub <- NULL
  if (!is.null(return_subset_genes)) {
    sub <- which(rownames(exprsMat) %in% return_subset_genes)
    if (length(sub) == 0) stop("No provided return_subset_genes genes in the data")
  }
  kk <- min(ruvK, nrow(fullalpha))
  cr <- attr(fullalpha, "conv_rank")
  if (!is.null(cr) && !is.na(cr) && kk > cr) warning("ruvK = ", kk, " exceeds the rank this fullalpha was converged for (", cr, ")")
  if (is.null(sub) && is.matrix(exprsMat) && is.double(exprsMat)) {
    ## plain matrix, all genes: ONE call on REAL(exprsMat) (column means accumulate during the chunked upload)
    fa <- fullalpha[seq_len(kk), rownames(exprsMat), drop = FALSE]
    r <- .Call("scm_host", exprsMat, NULL, 0L, NULL, 0L, ctl2, kk, 0L, 0L, t(fa), if (is.null(adjusted_means)) 1L else 2L, adjusted_means)
    out <- r[[1]]; dimnames(out) <- dimnames(exprsMat)
    return(out)
  }
  dY <- if (methods::is(exprsMat, "dgCMatrix")) .Call("scm_upload_sparse", exprsMat@p, exprsMat@i, exprsMat@x, dim(exprsMat), FALSE)
        else .Call("scm_upload", as.matrix(exprsMat), 0L)
  if (is.null(adjusted_means)) adjusted_means <- .Call("scm_segmean", dY, integer(ncol(exprsMat)), 1L)[[1]][, 1]
  fa <- fullalpha[seq_len(kk), rownames(exprsMat), drop = FALSE]         # name indexing, R/scMerge2.R:399-405
  out <- .Call("scm_adjust_", dY, t(fa), kk, ctl2, adjusted_means, NULL, NULL, if (is.null(sub)) NULL else as.integer(sub - 1L))
  rownames(out) <- if (is.null(sub)) rownames(exprsMat) else rownames(exprsMat)[sub]
  colnames(out) <- colnames(exprsMat)
  out
}
