## Drop-in bodies for the ARITHMETIC helpers of scReplicate (R/scReplicate.R): centroidDist (:458-464), compute_dist_mat (:470-490),
## compute_dist_res (:493-520), computePCA_byHVGMarker (:793-819), and identifyCluster (:358-447) with its kmeans(nstart = 1000) on the
## device.  findMNC's graph logic and igraph stay as they are; with useB200() these internal functions of scMerge are replaced too.
## Tested twin: scmerge_b200/replicate.py.
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


## stats::kmeans(x, centers = K, nstart) on the device (scm_kmeans): same fields as the "kmeans" object identifyCluster reads.  The
## starts come from a counter-based hash of `seed`, not from R's RNG stream (the reference's clustering depends on the RNG state too).
kmeansB200 <- function(x, centers, iter.max = 100L, nstart = 1L, seed = 1) {
  x <- as.matrix(x)
  if (length(centers) != 1L) stop("kmeansB200 takes the NUMBER of centers")
  res <- .Call("scm_kmeans_", t(x), as.integer(centers), as.integer(nstart), as.integer(iter.max), as.numeric(seed))
  names(res) <- c("cluster", "centers", "withinss", "tot.withinss", "size", "iter", "ifault")
  names(res$cluster) <- rownames(x)
  res$centers <- t(res$centers); colnames(res$centers) <- colnames(x); rownames(res$centers) <- seq_len(centers)
  if (any(res$size == 0L)) stop("empty cluster: try a better set of initial centers")
  if (res$ifault == 2L) warning(sprintf("did not converge in %d iterations", as.integer(iter.max)))
  res
}

## identifyCluster (R/scReplicate.R:358-447): per batch PCA -> kmeans(first 10 PCs, kmeansK[j], nstart = 1000) -> centroidDist of
## every cluster.  Same structure and names as the reference's; kmeans and centroidDist run on the device.
identifyCluster <- function(exprs_mat, batch, marker = NULL, HVG_list, kmeansK, BPPARAM, BSPARAM) {
  batch_list <- as.list(as.character(unique(batch)))
  batch_oneType <- unlist(batch_list)[which(kmeansK == 1)]
  batch_num <- table(batch)[as.character(unique(batch))]
  pca <- lapply(batch_list, function(b) computePCA_byHVGMarker(this_batch_list = b, batch = batch, batch_oneType = batch_oneType,
                                                              marker = marker, exprs_mat = exprs_mat, HVG_list = HVG_list, BSPARAM = BSPARAM))
  names(pca) <- unlist(batch_list)
  res <- lapply(seq_along(pca), function(j) {
    genes <- if (!is.null(marker)) marker else HVG_list[[j]]
    exprs_current <- as.matrix(exprs_mat[genes, batch == batch_list[j], drop = FALSE])
    if (!is.null(pca[[j]])) {
      cl <- kmeansB200(pca[[j]][, seq_len(10)], centers = kmeansK[j], nstart = 1000, seed = j)$cluster
    } else {
      cl <- rep(1, batch_num[j]); names(cl) <- colnames(exprs_current)
    }
    ## centroidDist of every cluster in one pass: squared distances to the cluster's per-gene medians, ranked within the cluster
    d <- .Call("scm_centroid_sqdist", exprs_current, as.integer(cl) - 1L, as.integer(max(cl)))
    pt <- stats::ave(d, cl, FUN = function(v) rank(v) / length(v)); names(pt) <- names(cl)
    list(cl, pt)
  })
  list(clustering_list = lapply(res, function(x) x[[1]]), clustering_distProp = lapply(res, function(x) x[[2]]))
}
