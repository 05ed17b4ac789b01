## Drop-in bodies for scMerge's RUV-III closures (same signatures as R/fastRUVIII.R:41-44, R/scRUVIII.R:41-44,
## R/scMerge2_utilis.R:17-26, R/scMerge2.R:368-372).  Only argument resolution lives here; all arithmetic is in
## libscmerge_b200 behind src/rshim.c.  scMerge()/scMerge2() call these exactly as they call the originals.

.keys <- function(M) {            # ruv::replicate.matrix semantics -> 0-based int keys (one-hot rows required)
  if (methods::is(M, "DelayedArray") || methods::is(M, "Matrix")) M <- as.matrix(M)   # the reference's tests pass DelayedArray(M)
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

## Gene-wise covariates of ruv::RUV1 -> the q x n design rows t(design.matrix(eta)) (tested twin: scmerge_b200.ruv.eta_design_rows)
.eta_rows <- function(eta, n, include.intercept) {
  if (is.numeric(eta) && length(eta) == 1L) { stopifnot(eta == 1); eta <- matrix(1, n, 1) }
  A <- as.matrix(eta)
  if (nrow(A) != n) A <- t(A)
  if (nrow(A) != n) stop("eta must have one row (or one column) per gene")
  if (include.intercept && sum(stats::lm.fit(A, rep(1, n))$residuals^2) > 1e-8) A <- cbind(1, A)
  d <- svd(scale(A, center = FALSE, scale = sqrt(colSums(A^2))))$d
  if (ncol(A) > n || d[length(d)] < 1e-9 * d[1]) stop("eta is collinear: remove the redundant covariates")
  t(A)
}
## Y <- ruv::RUV1(Y, eta, ctl) on a device handle (in place unless copy = TRUE): the rank-q update kernels with eta as alpha
.ruv1 <- function(dY, eta, ctl2, include.intercept, copy = FALSE) {
  E <- .eta_rows(eta, length(ctl2), include.intercept)
  if (nrow(E) > sum(ctl2)) stop("eta has more covariates than there are control genes")
  .Call("scm_ruv1_", dY, t(E), nrow(E), ctl2, as.integer(copy))
}

fastRUVIII <- function(Y, M, ctl, k = NULL, eta = NULL, svd_k = 50, include.intercept = TRUE, average = FALSE,
                       BPPARAM = SerialParam(), BSPARAM = ExactParam(), fullalpha = NULL, return.info = FALSE,
                       inputcheck = TRUE) {
  m <- nrow(Y); n <- ncol(Y)
  g <- .keys(M); ctl2 <- rep(FALSE, n); ctl2[ctl] <- TRUE
  exact <- inherits(BSPARAM, "ExactParam")
  s <- if (exact) min(m - g$G, sum(ctl2), svd_k, na.rm = TRUE) else min(m - g$G, sum(ctl2), na.rm = TRUE)
  if ((g$G >= m | k == 0) && is.null(eta)) {
    newY <- Y; fullalpha <- NULL
  } else if (g$G >= m | k == 0) {                                    # the early-out comes after RUV1 (R/fastRUVIII.R:64,78-80)
    E <- .eta_rows(eta, n, include.intercept)
    newY <- t(.Call("scm_adjust_", .Call("scm_upload", as.matrix(Y), 1L), t(E), nrow(E), ctl2, NULL, NULL, NULL, NULL))
    dimnames(newY) <- dimnames(Y); fullalpha <- NULL
  } else {
    dY <- .Call("scm_upload", as.matrix(Y), 1L)                     # cells x genes column-major -> device transpose
    if (!is.null(eta)) dY <- .ruv1(dY, eta, ctl2, include.intercept)  # R/fastRUVIII.R:63-64, in place on the private copy
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

## scRUVg (R/scRUVg.R:30-85) with Z = 1.  eta: the decomposition and alpha use RUV1(Y), W alpha is subtracted from Y itself (:59,80);
## fullW: W = fullW[, 1:k], alpha by least squares (:64,77-78).  svdyc without fullW fails in the reference too (:77).
scRUVg <- function(Y, ctl, k, Z = 1, eta = NULL, include.intercept = TRUE, fullW = NULL, svdyc = NULL) {
  stopifnot(identical(Z, 1))
  if (!is.null(svdyc) && is.null(fullW)) stop("svdyc without fullW: pass fullW = svdyc$u (the reference never derives fullW from svdyc)")
  n <- ncol(Y); ctl2 <- rep(FALSE, n); ctl2[ctl] <- TRUE
  if (k > sum(ctl2)) stop("k must not be larger than the number of controls")
  dY <- .Call("scm_upload", as.matrix(Y), 1L)
  if (is.null(eta) && is.null(fullW)) {
    res <- .Call("scm_ruvg", dY, ctl2, as.integer(k))
  } else {
    dY0 <- if (is.null(eta)) dY else .ruv1(dY, eta, ctl2, include.intercept, copy = TRUE)
    Wt <- if (is.null(fullW)) NULL else t(as.matrix(fullW)[, seq_len(k), drop = FALSE])
    res <- .Call("scm_ruvg2", dY, dY0, ctl2, as.integer(k), Wt)
  }
  list(newY = DelayedArray::DelayedArray(t(res[[1]])), W = t(res[[2]]), alpha = t(res[[3]]))
}

## inside scRUVIII (R/scRUVIII.R:83-105), for length(k) > 1:
##   sil_res <- sapply(k, function(kk) .Call("scm_sil", dY, fit$gram, fit$colmean, t(fullalpha[seq_len(kk), ]), kk, ctl2,
##                                           stand_mean, 1 / stand_sd, stand_sd, ct_key, n_ct, batch_key, n_batch, 10L))
##   f_score <- f_measure(zeroOneScale(sil_res[1, ]), 1 - zeroOneScale(sil_res[2, ]))
## inside scMerge2 (R/scMerge2.R:164-198): dY <- .Call("scm_upload_sparse", exprsMat@p, exprsMat@i, exprsMat@x, dim(exprsMat), cosineNorm)
