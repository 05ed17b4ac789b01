## Wiring into scMerge: scMerge() (R/scMerge.R:140-143 -> scRUVIII -> fastRUVIII) and scMerge2() (R/scMerge2.R:298-309 ->
## pseudoRUVIII; getAdjustedMat for return_matrix = FALSE) resolve these closures in scMerge's own namespace, so replacing the
## bindings there makes the unmodified drivers run the B200 path.  useB200(FALSE) restores the originals.
.saved <- new.env(parent = emptyenv())
.closures <- c("fastRUVIII", "scRUVIII", "pseudoRUVIII", "getAdjustedMat", "scRUVg",
               "centroidDist", "compute_dist_mat", "compute_dist_res", "computePCA_byHVGMarker",   # scReplicate's arithmetic helpers
               "identifyCluster")                                                                   # ... and its kmeans(nstart = 1000)

useB200 <- function(on = TRUE) {
  if (!requireNamespace("scMerge", quietly = TRUE)) stop("scMerge is not installed")
  ns <- asNamespace("scMerge")
  for (f in .closures) {
    if (on) {
      if (is.null(.saved[[f]])) .saved[[f]] <- get(f, envir = ns)
      utils::assignInNamespace(f, get(f, envir = asNamespace("scMergeB200")), ns = "scMerge")
    } else if (!is.null(.saved[[f]])) {
      utils::assignInNamespace(f, .saved[[f]], ns = "scMerge")
    }
  }
  invisible(on)
}

## One R process, several GPUs (scm_multi_create: a context, stream and NCCL rank per device inside the library).  The
## host-matrix closures (scRUVIII with one ruvK, getAdjustedMat, pseudoRUVIII's adjust) then split the cells over all devices.
useGPUs <- function(n = 1L) .Call("scm_use_gpus", as.integer(n))
