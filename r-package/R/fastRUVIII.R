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
.rows_cap <- function(s, k, exact) { cap <- if (exact) 88L else 88L; if (s > cap) list(s = min(s, max(k, 50L), cap), exact = TRUE) else list(s = s, exact = exact) }

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
      cap <- .rows_cap(s, k, exact)
      fit <- .Call("scm_fit", dY, g$key, g$G, NULL, 0L, cap$s, min(k, cap$s), if (cap$exact) 0L else 1L)
      fullalpha <- t(fit[[1]]); colnames(fullalpha) <- colnames(Y)
    }
    kk <- min(k, nrow(fullalpha))
    out <- .Call("scm_adjust_", dY, t(fullalpha[seq_len(kk), , drop = FALSE]), kk, ctl2, NULL, NULL, NULL, NULL)
    newY <- DelayedArray::DelayedArray(t(out)); dimnames(newY) <- dimnames(Y)
  }
  if (!return.info) newY else list(newY = newY, M = M, fullalpha = fullalpha)
}
