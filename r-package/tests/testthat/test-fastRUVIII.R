## Mirror of the reference's tests/testthat/test-fastRUVIII.R:3-16, run against the B200 closures, plus the equality with
## ruv::RUVIII that the reference documents (R/fastRUVIII.R:30-38) but leaves commented out (:15).
context("Test fastRUVIII (B200)")
library(BiocSingular)
skip_if_not_installed("scMerge"); skip_if_not_installed("ruv")
set.seed(12345)
L = scMerge::ruvSimulate(m = 400, n = 500, nc = 400, nCelltypes = 3, nBatch = 2, lambda = 0.1, sce = FALSE)
Y = L$Y; M = L$M; ctl = L$ctl

old = ruv::RUVIII(Y = Y, M = M, ctl = ctl, k = 20)
ref1 = scMerge::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 20, BSPARAM = ExactParam())
improved1 = scMergeB200::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 20, BSPARAM = ExactParam())
improved2 = scMergeB200::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 20, BSPARAM = RandomParam(), svd_k = 100)

rel_frob <- function(a, b) sqrt(sum((as.matrix(a) - as.matrix(b))^2) / sum(as.matrix(b)^2))
test_that("exact and randomized agree like in the reference", expect_equal(as.matrix(improved1), as.matrix(improved2), tol = 0.02))
test_that("the device path reproduces the reference's exact fastRUVIII", expect_lt(rel_frob(improved1, ref1), 1e-8))
test_that("and ruv::RUVIII", expect_lt(rel_frob(improved1, old), 1e-8))
test_that("fullalpha spans the same subspace", {
  a <- scMerge::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 20, return.info = TRUE)$fullalpha[1:20, ]
  b <- scMergeB200::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 20, return.info = TRUE)$fullalpha[1:20, ]
  angles <- acos(pmin(1, svd(crossprod(qr.Q(qr(t(a))), qr.Q(qr(t(b)))))$d))
  expect_lt(max(angles), 1e-6)
})
