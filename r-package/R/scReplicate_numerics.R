## Drop-in bodies for the ARITHMETIC helpers of scReplicate (R/scReplicate.R): centroidDist (:458-464), compute_dist_mat (:470-490),
## compute_dist_res (:493-520), computePCA_byHVGMarker (:793-819).  kmeans, findMNC's graph logic and igraph stay as they are; with
## useB200() these four internal functions of scMerge are replaced too.  Tested twin: scmerge_b200/replicate.py.
.metric <- function(dist) switch(dist, cosine = 1L, cor = 2L, 0L)

centroidDist <- function(exprs_mat) {
  point_dist <- .Call("scm_centroid_sqdist", as.matrix(exprs_mat), integer(ncol(exprs_mat)), 1L)
  point_rank <- rank(point_dist)
  point_rank / length(point_rank)
}

compute_dist_mat <- function(k, exprs_mat, clustering_list, combine_pair, dist) {
  res1 <- clustering_list[[combine_pair[1, k]]]; res2 <- clustering_list[[combine_pair[2, k]]]
  t(.Call("scm_dist", as.matrix(exprs_mat[, names(res1), drop = FALSE]), as.matrix(exprs_mat[, names(res2), drop = FALSE]),
          .metric(dist), NULL, NULL))
}

compute_dist_res <- function(k, exprs_mat, clustering_list, combine_pair, dist) {
  res1 <- clustering_list[[combine_pair[1, k]]]; res2 <- clustering_list[[combine_pair[2, k]]]
  o1 <- order(res1); o2 <- order(res2)                       # cells by cluster: every cluster pair is a rectangle of the distance matrix
  roff <- c(0, cumsum(tabulate(res1, nbins = max(res1)))); coff <- c(0, cumsum(tabulate(res2, nbins = max(res2))))
  t(.Call("scm_dist", as.matrix(exprs_mat[, names(res1)[o1], drop = FALSE]), as.matrix(exprs_mat[, names(res2)[o2], drop = FALSE]),
          .metric(dist), as.numeric(roff), as.numeric(coff)))
}

computePCA_byHVGMarker <- function(this_batch_list, batch, batch_oneType, marker, exprs_mat, HVG_list, BSPARAM) {
  if (this_batch_list %in% batch_oneType) return(NULL)
  genes <- if (!is.null(marker)) marker else HVG_list[[this_batch_list]]
  sub <- as.matrix(exprs_mat[genes, batch == this_batch_list])
  result <- t(.Call("scm_pca", sub, 10L))
  rownames(result) <- colnames(sub)
  result
}
