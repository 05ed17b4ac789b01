## Mirror of the reference's tests/testthat/test-scRUVg.R:1-7 against the B200 closure, plus equality with the reference's own scRUVg.
context("Test scRUVg (B200)")
skip_if_not_installed("scMerge"); skip_if_not_installed("ruv")
set.seed(1234)
L = scMerge::ruvSimulate(m = 80, n = 1000, nc = 50, nCelltypes = 10)
Y = L$Y; ctl = L$ctl; X = L$X
ruvgRes = scMergeB200::scRUVg(Y = Y, ctl = ctl, k = 20)
old = ruv::RUV2(Y = Y, X = X, ctl = ctl, k = 20)
test_that("W equals RUV2's up to sign (the reference's assertion)", expect_equal(abs(ruvgRes$W), abs(old$W)))
test_that("newY equals the reference's scRUVg", {
  ref = scMerge:::scRUVg(Y = Y, ctl = ctl, k = 20)
  expect_lt(sqrt(sum((as.matrix(ruvgRes$newY) - as.matrix(ref$newY))^2) / sum(as.matrix(ref$newY)^2)), 1e-8)
})
test_that("k larger than the number of controls is an error, as in the reference", expect_error(scMergeB200::scRUVg(Y = Y, ctl = ctl, k = 51)))
