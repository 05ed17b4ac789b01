## Drop-in bodies for scMerge's RUV-III closures (same signatures as R/fastRUVIII.R:41-44, R/scRUVIII.R:41-44,
## R/scMerge2_utilis.R:17-26, R/scMerge2.R:368-372).  Only argument resolution lives here; all arithmetic is in
## libscmerge_b200 behind src/rshim.c.  scMerge()/scMerge2() call these exactly as they call the originals.

.keys <- function(M) {            # ruv::replicate.matrix semantics -> 0-based int keys (one-hot rows required)
  if (is.matrix(M) && is.numeric(M)) {
    stopifnot(all(M == 0 | M == 1), all(rowSums(M) == 1))
    list(key = as.integer(max.col(M, ties.method = "first") - 1L), G = ncol(M))
  } else {
    f <- if (is.data.frame(M)) factor(do.call(paste, c(M, sep = "_"))) else factor(M)
    list(key = as.integer(f) - 1L, G = nlevels(f))
  }
}
## Rows of fullalpha the device returns.  Up to 88 (SCM_MAX_TOPK_ROWS) the subspace iteration serves.  More (the non-exact
## branch asks for min(m - G, n_ctl) rows) come from the whole spectrum: of the n x n Gram when n <= 4096
## (SCM_MAX_FULL_EIG_N), of the rows x rows side for a short matrix (pseudobulks; any n).  Otherwise, or with
## options(scMergeB200.fullalpha_rows = "truncate"): max(k, 50) converged rows and a warning.
.rows_cap <- function(s, k, exact, n = NA_integer_, rows = NA_integer_) {
  cap <- 88L
  if (s <= cap) return(list(s = s, exact = exact))
  full <- !identical(getOption("scMergeB200.fullalpha_rows", "reference"), "truncate") && !is.na(n) &&
    (n <= 4096L || (!is.na(rows) && rows < n && rows <= 4096L))
  if (full) return(list(s = s, exact = TRUE))
  want <- min(s, max(k, 50L), cap)
  warning(sprintf("fullalpha truncated to %d rows (reference rule gives %d)", want, s))
  list(s = want, exact = TRUE)
}

fastRUVIII <- function(Y, M, ctl, k = NULL, eta = NULL, svd_k = 50, include.intercept = TRUE, average = FALSE,
                       BPPARAM = SerialParam(), BSPARAM = ExactParam(), fullalpha = NULL, return.info = FALSE,
                       inputcheck = TRUE) {
  if (!is.null(eta)) stop("eta is not supported on the device path")
  m <- nrow(Y); n <- ncol(Y)
  g <- .keys(M); ctl2 <- rep(FALSE, n); ctl2[ctl] <- TRUE
  exact <- inherits(BSPARAM, "ExactParam")
  s <- if (exact) min(m - g$G, sum(ctl2), svd_k, na.rm = TRUE) else min(m - g$G, sum(ctl2), na.rm = TRUE)
  if (g$G >= m | k == 0) {
    newY <- Y; fullalpha <- NULL
  } else {
    dY <- .Call("scm_upload", as.matrix(Y), 1L)                     # cells x genes column-major -> device transpose
    if (is.null(fullalpha)) {
      cap <- .rows_cap(s, k, exact, n)
      fit <- .Call("scm_fit", dY, g$key, g$G, NULL, 0L, cap$s, min(k, cap$s), if (cap$exact) 0L else 1L)
      fullalpha <- t(fit[[1]]); colnames(fullalpha) <- colnames(Y)
    }
    kk <- min(k, nrow(fullalpha))
    out <- .Call("scm_adjust_", dY, t(fullalpha[seq_len(kk), , drop = FALSE]), kk, ctl2, NULL, NULL, NULL, NULL)
    newY <- DelayedArray::DelayedArray(t(out)); dimnames(newY) <- dimnames(Y)
  }
  if (!return.info) newY else list(newY = newY, M = M, fullalpha = fullalpha)
}

## scRUVg (R/scRUVg.R:30-85) with the default Z = 1, eta = NULL
scRUVg <- function(Y, ctl, k, Z = 1, eta = NULL, include.intercept = TRUE, fullW = NULL, svdyc = NULL) {
  stopifnot(is.null(eta), identical(Z, 1), is.null(fullW), is.null(svdyc))
  n <- ncol(Y); ctl2 <- rep(FALSE, n); ctl2[ctl] <- TRUE
  if (k > sum(ctl2)) stop("k must not be larger than the number of controls")
  res <- .Call("scm_ruvg", .Call("scm_upload", as.matrix(Y), 1L), ctl2, as.integer(k))
  list(newY = DelayedArray::DelayedArray(t(res[[1]])), W = t(res[[2]]), alpha = t(res[[3]]))
}

## inside scRUVIII (R/scRUVIII.R:83-105), for length(k) > 1:
##   sil_res <- sapply(k, function(kk) .Call("scm_sil", dY, fit$gram, fit$colmean, t(fullalpha[seq_len(kk), ]), kk, ctl2,
##                                           stand_mean, 1 / stand_sd, stand_sd, ct_key, n_ct, batch_key, n_batch, 10L))
##   f_score <- f_measure(zeroOneScale(sil_res[1, ]), 1 - zeroOneScale(sil_res[2, ]))
## inside scMerge2 (R/scMerge2.R:164-198): dY <- .Call("scm_upload_sparse", exprsMat@p, exprsMat@i, exprsMat@x, dim(exprsMat), cosineNorm)
