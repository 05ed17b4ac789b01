## eta (ruv::RUV1), scRUVg's fullW re-use and the device kmeans against the reference's own functions (needs a B200, scMerge, ruv).
context("eta, fullW, kmeans on the device")

skip_if_not_installed("scMerge"); skip_if_not_installed("ruv")

L <- scMerge::ruvSimulate(m = 200, n = 500, nc = 100, nCelltypes = 3, nBatch = 2, lambda = 0.1, sce = FALSE)
Y <- L$Y; M <- L$M; ctl <- L$ctl
eta <- cbind(colMeans(Y), colMeans(Y == 0))

test_that("fastRUVIII with eta equals the reference", {
  ref <- scMerge::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 5, eta = eta, BSPARAM = BiocSingular::ExactParam())
  got <- scMergeB200::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 5, eta = eta, BSPARAM = BiocSingular::ExactParam())
  expect_lt(sqrt(sum((as.matrix(got) - as.matrix(ref))^2)) / sqrt(sum(as.matrix(ref)^2)), 1e-8)
  expect_equal(as.matrix(scMergeB200::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 0, eta = eta)), ruv::RUV1(Y, eta, ctl), tolerance = 1e-10)
})

test_that("scRUVg with eta and with a re-used fullW equals the reference", {
  ref <- scMerge:::scRUVg(Y = Y, ctl = ctl, k = 4, eta = eta)
  got <- scMergeB200::scRUVg(Y = Y, ctl = ctl, k = 4, eta = eta)
  expect_lt(sqrt(sum((as.matrix(got$newY) - as.matrix(ref$newY))^2)) / sqrt(sum(as.matrix(ref$newY)^2)), 1e-8)
  expect_equal(abs(got$W), abs(ref$W), tolerance = 1e-6)            # the reference's own sign-invariant idiom (test-scRUVg.R:7)
  base <- scMerge:::scRUVg(Y = Y, ctl = ctl, k = 8)
  ref2 <- scMerge:::scRUVg(Y = Y, ctl = ctl, k = 3, fullW = base$W)
  got2 <- scMergeB200::scRUVg(Y = Y, ctl = ctl, k = 3, fullW = base$W)
  expect_lt(sqrt(sum((as.matrix(got2$newY) - as.matrix(ref2$newY))^2)) / sqrt(sum(as.matrix(ref2$newY)^2)), 1e-8)
  expect_error(scMergeB200::scRUVg(Y = Y, ctl = ctl, k = 3, svdyc = svd(Y[, ctl])))
})

test_that("kmeansB200 reaches the optimum of stats::kmeans(nstart = 1000)", {
  set.seed(1)
  x <- rbind(matrix(rnorm(300 * 10, 0), 300), matrix(rnorm(300 * 10, 3), 300), matrix(rnorm(300 * 10, -3), 300))
  ref <- stats::kmeans(x, centers = 3, nstart = 1000)
  got <- scMergeB200::kmeansB200(x, centers = 3, nstart = 1000)
  expect_equal(got$tot.withinss, ref$tot.withinss, tolerance = 1e-10)
  expect_equal(sort(got$size), sort(ref$size))
  expect_true(all(table(got$cluster, ref$cluster) %in% c(0L, 300L)))   # the same partition up to the cluster ids
})
