## Mirror of the reference's tests/testthat/test-resampling-columns.R:13-37 with scMerge() running on the B200 closures
## (useB200() patches fastRUVIII / scRUVIII / pseudoRUVIII / getAdjustedMat into scMerge's namespace): cell-permutation
## equivariance of the adjusted matrix at expect_equal's default tolerance (1.5e-8), and agreement with the CPU reference.
context("Testing identical results when permuting columns (B200)")
skip_if_not_installed("scMerge")
library(SingleCellExperiment); library(scMerge)
data("example_sce", package = "scMerge"); data("segList_ensemblGeneID", package = "scMerge")
ctl <- segList_ensemblGeneID$mouse$mouse_scSEG

cpu <- scMerge(sce_combine = example_sce, ctl = ctl, kmeansK = c(3, 3), assay_name = "scMerge1", replicate_prop = 1)
scMergeB200::useB200(TRUE)
res1 <- scMerge(sce_combine = example_sce, ctl = ctl, kmeansK = c(3, 3), assay_name = "scMerge1", replicate_prop = 1)
set.seed(2)
index <- sample(seq_len(ncol(example_sce)))
res2 <- scMerge(sce_combine = example_sce[, index], ctl = ctl, kmeansK = c(3, 3), assay_name = "scMerge2", replicate_prop = 1)
scMergeB200::useB200(FALSE)

test_that("output matrices are numerically equal after un-permuting",
          expect_equal(as(assay(res1, "scMerge1"), "dgCMatrix"), as(assay(res2, "scMerge2"), "dgCMatrix")[, order(index)]))
test_that("and equal to the unpatched CPU run",
          expect_equal(as.matrix(assay(res1, "scMerge1")), as.matrix(assay(cpu, "scMerge1")), tolerance = 1e-8))
