#!/usr/bin/env Rscript
## Golden vectors produced BY THE REFERENCE ITSELF (SydneyBioX/scMerge, unmodified) for the RUV-III path.
##
## This repository's build image has no R, so its parity tests are pinned on a NumPy restatement of the reference
## (oracle/ruv_oracle.py).  On any box with R >= 4.2 and Bioconductor's scMerge installed,
##
##     Rscript tools/r/make_reference_golden.R            # writes tests/golden/r_reference/*.csv (+ MANIFEST.txt)
##     python -m pytest tests/test_r_reference.py -q      # oracle vs reference (CPU);  add `-m gpu` on a B200 for the CUDA path
##
## turns "parity unpinned" into a check against values computed by R/fastRUVIII.R, R/scRUVIII.R, R/scMerge2_utilis.R,
## R/scMerge2.R and R/scRUVg.R.  Inputs: the package's bundled example_sce (1047 genes x 200 cells) with the supervised
## replicate structure M = one-hot(cellTypes) (what scMerge(cell_type = ...) builds, R/scReplicate.R:768-789) and the mouse
## scSEG control genes -- exactly tests/golden/example_sce.npz.  Everything here is deterministic except cvTools::cvFolds
## (seeded, and its fold assignment is written out so the device path can consume the same draw).
suppressPackageStartupMessages({
  library(scMerge); library(SingleCellExperiment); library(BiocSingular); library(ruv)
})
args <- commandArgs(trailingOnly = TRUE)
out_dir <- if (length(args) >= 1) args[1] else file.path("tests", "golden", "r_reference")
dir.create(out_dir, recursive = TRUE, showWarnings = FALSE)

num <- function(x) formatC(x, digits = 17, format = "g", flag = "-")          # exact float64 round trip
write_mat <- function(x, name) {
  x <- as.matrix(x)
  write.table(apply(x, 2, num), file.path(out_dir, paste0(name, ".csv")), sep = ",", quote = FALSE, row.names = FALSE, col.names = FALSE)
}
write_vec <- function(x, name) writeLines(if (is.numeric(x)) num(x) else as.character(x), file.path(out_dir, paste0(name, ".txt")))

data("example_sce", package = "scMerge")
data("segList_ensemblGeneID", package = "scMerge")
exprs <- as.matrix(SummarizedExperiment::assay(example_sce, "logcounts"))      # genes x cells
ctl <- which(rownames(exprs) %in% segList_ensemblGeneID$mouse$mouse_scSEG)     # 1-based gene indices
cell_types <- as.character(example_sce$cellTypes)
batch <- as.character(example_sce$batch)
Y <- t(exprs)                                                                  # cells x genes
M <- ruv::replicate.matrix(cell_types)
rows <- 1:16                                                                   # cells whose adjusted rows are stored in full
write_vec(rownames(exprs), "gene_names"); write_vec(colnames(exprs), "cell_names")
write_vec(ctl - 1L, "ctl_index0"); write_vec(cell_types, "cell_types"); write_vec(batch, "batch")
write_vec(c(dim(exprs), sum(exprs)), "input_dims_and_sum")                     # guards against a different example_sce

## 1. fastRUVIII, exact SVD (R/fastRUVIII.R:41-111)
r1 <- scMerge::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 20, BSPARAM = BiocSingular::ExactParam(), return.info = TRUE)
write_mat(as.matrix(r1$newY)[rows, ], "fastRUVIII_k20_newY_rows"); write_mat(r1$fullalpha, "fastRUVIII_fullalpha")
write_vec(sqrt(sum(as.matrix(r1$newY)^2)), "fastRUVIII_k20_newY_frobenius")
## 1b. ruv::RUVIII, the equality the reference documents but does not assert (R/fastRUVIII.R:30-38, tests/testthat/test-fastRUVIII.R:15)
r1b <- ruv::RUVIII(Y = Y, M = M, ctl = ctl, k = 20)
write_mat(r1b[rows, ], "ruv_RUVIII_k20_newY_rows")

## 2. scRUVIII, single k and the ruvK selection (R/scRUVIII.R:41-137); standardize2 is internal (R/scRUVIII.R:158-175)
st <- scMerge:::standardize2(exprs, batch)
write_vec(st$stand_mean, "stand_mean"); write_vec(sqrt(st$stand_var), "stand_sd")
r2 <- scMerge::scRUVIII(Y = exprs, M = M, ctl = ctl, k = 20, batch = batch, cell_type = cell_types, BSPARAM = BiocSingular::ExactParam())
write_mat(as.matrix(r2[["20"]]$newY)[rows, ], "scRUVIII_k20_newY_rows")
grDevices::pdf(NULL)                                                           # scRUVIII plots when length(k) > 1 (R/scRUVIII.R:109-112)
r3 <- scMerge::scRUVIII(Y = exprs, M = M, ctl = ctl, k = c(5, 10, 20), batch = batch, cell_type = cell_types,
                        BSPARAM = BiocSingular::ExactParam())
write_vec(r3$optimal_ruvK, "scRUVIII_k5_10_20_optimal_ruvK")
for (kk in c(5, 10, 20)) write_mat(as.matrix(r3[[as.character(kk)]]$newY)[rows, ], paste0("scRUVIII_multi_k", kk, "_newY_rows"))

## 3. create_pseudoBulk per batch + pseudoRUVIII + getAdjustedMat (R/scMerge2_utilis.R:17-150,461-486; R/scMerge2.R:368-410)
##    with the supervised replicate structure (pseudobulks of the same cell type are replicates) and NO cosine normalisation,
##    so that the inputs stay the bundled logcounts.
batch_list <- sort(unique(batch)); k_fold <- 6
which_fold <- integer(ncol(exprs)); bulks <- list()
for (i in seq_along(batch_list)) {
  sel <- which(batch == batch_list[i])
  set.seed(100 + i); cv <- cvTools::cvFolds(length(sel), K = min(length(sel), k_fold))     # the draw create_pseudoBulk will make
  which_fold[sel[cv$subsets[, 1]]] <- cv$which
  set.seed(100 + i); res <- scMerge:::create_pseudoBulk(exprs[, sel], cell_types[sel], k_fold = k_fold)
  rownames(res) <- paste(i, rownames(res), sep = "_")                                       # R/scMerge2_utilis.R:266
  bulks[[i]] <- as.matrix(res)
}
bulk <- do.call(rbind, bulks)                                                  # pseudobulks x genes
write_vec(which_fold, "pseudobulk_fold_of_cell"); write_vec(rownames(bulk), "pseudobulk_names"); write_mat(bulk, "pseudobulk_means")
rep_vec <- vapply(strsplit(rownames(bulk), "_"), function(p) p[2], "")
Mp <- ruv::replicate.matrix(rep_vec)
r4 <- scMerge:::pseudoRUVIII(Y = Y, Y_pseudo = bulk, M = Mp, ctl = seq_len(ncol(Y)) %in% ctl, k = 20,
                             BSPARAM = BiocSingular::ExactParam(), return.info = TRUE, normalised = TRUE, verbose = FALSE)
write_mat(as.matrix(r4$newY)[rows, ], "pseudoRUVIII_k20_newY_rows"); write_mat(r4$fullalpha, "pseudoRUVIII_fullalpha")
fa <- r4$fullalpha; colnames(fa) <- rownames(exprs)
r5 <- scMerge::getAdjustedMat(exprs, fa, ctl = rownames(exprs)[ctl], ruvK = 10)
write_mat(t(as.matrix(r5))[rows, ], "getAdjustedMat_k10_newY_rows")

## 4. scRUVg (R/scRUVg.R:30-85)
r6 <- scMerge:::scRUVg(Y = Y, ctl = ctl, k = 5)
write_mat(as.matrix(r6$newY)[rows, ], "scRUVg_k5_newY_rows"); write_mat(abs(r6$W), "scRUVg_k5_absW")

## 5. gene-wise covariates eta (ruv::RUV1 at R/fastRUVIII.R:63-64 and R/scRUVg.R:59) and a re-used fullW (R/scRUVg.R:64,77-80).
##    eta: two deterministic covariates per gene (its mean expression and the fraction of zero cells), intercept added by RUV1.
eta <- cbind(rowMeans(exprs), rowMeans(exprs == 0))
write_mat(eta, "eta_genes_x_2")
r7 <- scMerge::fastRUVIII(Y = Y, M = M, ctl = ctl, k = 10, eta = eta, BSPARAM = BiocSingular::ExactParam(), return.info = TRUE)
write_mat(as.matrix(r7$newY)[rows, ], "fastRUVIII_eta_k10_newY_rows")
write_mat(ruv::RUV1(Y, eta, seq_len(ncol(Y)) %in% ctl)[rows, ], "RUV1_eta_rows")
r8 <- scMerge:::scRUVg(Y = Y, ctl = ctl, k = 5, eta = eta)
write_mat(as.matrix(r8$newY)[rows, ], "scRUVg_eta_k5_newY_rows")
r9 <- scMerge:::scRUVg(Y = Y, ctl = ctl, k = 3, fullW = r6$W)
write_mat(as.matrix(r9$newY)[rows, ], "scRUVg_fullW_k3_newY_rows"); write_mat(r9$alpha, "scRUVg_fullW_k3_alpha")

## 6. kmeans as identifyCluster calls it (R/scReplicate.R:394-395) on the first 10 PCs of every batch.  The labels depend on R's
##    RNG stream; what is compared is the optimum: tot.withinss and the partition (as a co-membership checksum).
for (i in seq_along(batch_list)) {
  sel <- which(batch == batch_list[i])
  pca <- BiocSingular::runPCA(t(exprs[, sel]), rank = 10, scale = TRUE, center = TRUE, BSPARAM = BiocSingular::ExactParam())$x
  set.seed(1); km <- stats::kmeans(pca[, seq_len(10)], centers = 3, nstart = 1000)
  write_mat(pca, paste0("kmeans_batch", i, "_pca_scores")); write_vec(km$cluster, paste0("kmeans_batch", i, "_cluster"))
  write_vec(c(km$tot.withinss, sort(km$size)), paste0("kmeans_batch", i, "_totwithinss_and_sizes"))
}

writeLines(c(paste("scMerge", as.character(utils::packageVersion("scMerge"))), paste("ruv", as.character(utils::packageVersion("ruv"))),
             paste("BiocSingular", as.character(utils::packageVersion("BiocSingular"))), R.version.string,
             paste("generated", format(Sys.time(), tz = "UTC", usetz = TRUE))), file.path(out_dir, "MANIFEST.txt"))
cat("wrote", length(list.files(out_dir)), "files to", out_dir, "\n")
