"""Replicate-finding numerics of scReplicate on the device: what is ARITHMETIC in
R/scReplicate.R, including the kmeans with 1000 starts per batch.  The mutual-nearest-cluster graph and igraph stay in R (out
of scope); their inputs and the quantities they consume come from here.

    centroidDist                R/scReplicate.R:458-464      per-gene medians of a cluster, squared distances, ranks
    cluster_centroid_dist       R/scReplicate.R:396-413      the same for every kmeans cluster of a batch in one pass
    compute_dist_mat / _res     R/scReplicate.R:470-520      cell-cell distances between two batches, medians per cluster pair
    computePCA_byHVGMarker      R/scReplicate.R:793-819      runPCA(rank 10, centred, scaled) of one batch on its genes
    kmeans                      R/scReplicate.R:394-395      stats::kmeans(pca[, 1:10], centers = K, nstart = 1000)
    identifyCluster             R/scReplicate.R:330-447      per batch: PCA -> kmeans -> centroidDist of every cluster

Matrices follow the reference: `exprs_mat` is genes x cells (numpy; column-major R memory is the device's cells x genes rows).
No arithmetic happens in Python: ranks (an ordering) and cluster offsets are index plumbing.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib as L
from .device import Context, DeviceMatrix, default_context

__all__ = ["centroidDist", "cluster_centroid_dist", "compute_dist_mat", "compute_dist_res", "computePCA_byHVGMarker", "run_pca_scores",
           "kmeans", "identifyCluster"]

_METRIC = {"euclidean": 0, "cosine": 1, "cor": 2}


def _cells_by_genes(ctx: Context, exprs_mat, genes=None, cells=None) -> DeviceMatrix:
    """exprs_mat[genes, cells] (genes x cells) as a device matrix cells x genes (scm_gather on the uploaded matrix)."""
    X = np.asarray(exprs_mat, dtype=np.float64)
    full = DeviceMatrix.from_host(ctx, X.T)
    if genes is None and cells is None:
        return full
    rows = None if cells is None else ctx.to_device(np.ascontiguousarray(cells, dtype=np.int32))
    cols = None if genes is None else ctx.to_device(np.ascontiguousarray(genes, dtype=np.int32))
    msub = full.m if cells is None else int(np.shape(cells)[0])
    nsub = full.n if genes is None else int(np.shape(genes)[0])
    out = DeviceMatrix(ctx, msub, nsub)
    L.check(L.lib().scm_gather(ctx.handle, full.ptr, full.m, full.n, full.ld, rows.ptr if rows is not None else None, msub,
                               cols.ptr if cols is not None else None, nsub, out.ptr, out.ld))
    ctx.sync()
    for a in (rows, cols):
        if a is not None:
            a.free()
    full.free()
    return out


def _rank_fraction(point_dist: np.ndarray) -> np.ndarray:
    """rank(point_dist) / length(point_dist) with R's default ties.method = "average" (R/scReplicate.R:461-462)."""
    order = np.argsort(point_dist, kind="stable")
    sd = point_dist[order]
    ranks = np.empty(sd.shape[0])
    i = 0
    while i < sd.shape[0]:
        j = i
        while j + 1 < sd.shape[0] and sd[j + 1] == sd[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks / ranks.shape[0]


def cluster_centroid_dist(exprs_mat, cluster, ctx: Context | None = None):
    """centroidDist for every cluster of a batch at once (what scReplicate assembles cluster by cluster, R/scReplicate.R:396-413):
    `cluster` = kmeans label per cell (any labels).  Returns (rank fraction per cell within its cluster, squared distances)."""
    ctx = ctx or default_context()
    lib = L.lib()
    levels, key = np.unique(np.asarray(cluster), return_inverse=True)
    key = np.ascontiguousarray(key, dtype=np.int32)
    G = int(levels.shape[0])
    dX = _cells_by_genes(ctx, exprs_mat)
    d_key = ctx.to_device(key)
    d_med, d_dist = ctx.empty((G, dX.n)), ctx.empty((dX.m,))
    try:
        L.check(lib.scm_group_median(ctx.handle, dX.ptr, dX.m, dX.n, dX.ld, key.ctypes.data_as(C.c_void_p), d_key.ptr, G, d_med.ptr))
        L.check(lib.scm_sqdist_to_center(ctx.handle, dX.ptr, dX.m, dX.n, dX.ld, d_key.ptr, G, d_med.ptr, d_dist.ptr))
        dist = d_dist.to_host()
    finally:
        for a in (d_key, d_med, d_dist):
            a.free()
        dX.free()
    out = np.empty(dist.shape[0])
    for g in range(G):
        sel = key == g
        out[sel] = _rank_fraction(dist[sel])
    return out, dist


def centroidDist(exprs_mat, ctx: Context | None = None) -> np.ndarray:
    """R/scReplicate.R:458-464.  exprs_mat: genes x cells of ONE cluster; returns rank(point_dist) / length."""
    m = np.shape(exprs_mat)[1]
    return cluster_centroid_dist(exprs_mat, np.zeros(m, dtype=np.int32), ctx)[0]


def _pair_dist_device(ctx: Context, dA: DeviceMatrix, dB: DeviceMatrix, dist: str):
    if dist not in _METRIC:
        raise ValueError("dist must be 'cor', 'cosine' or 'euclidean'")
    d_D = ctx.empty((dA.m, dB.m))
    L.check(L.lib().scm_pair_dist(ctx.handle, dA.ptr, dA.m, dA.ld, dB.ptr, dB.m, dB.ld, dA.n, _METRIC[dist], d_D.ptr, dB.m))
    return d_D


def compute_dist_mat(mat1, mat2, dist: str = "cor", ctx: Context | None = None) -> np.ndarray:
    """R/scReplicate.R:470-490.  mat1 / mat2: genes x cells; returns the cells1 x cells2 distance matrix
    (1 - proxyC::simil(correlation | cosine) or proxyC::dist)."""
    ctx = ctx or default_context()
    dA, dB = _cells_by_genes(ctx, mat1), _cells_by_genes(ctx, mat2)
    try:
        d_D = _pair_dist_device(ctx, dA, dB, dist)
        out = d_D.to_host()
        d_D.free()
    finally:
        dA.free()
        dB.free()
    return out


def compute_dist_res(exprs_mat, cells1, res1, cells2, res2, dist: str = "cor", ctx: Context | None = None) -> np.ndarray:
    """R/scReplicate.R:493-520 for one pair of batches.  cells1 / cells2: column indices (0-based) of the two batches' cells in
    `exprs_mat` (genes x cells), res1 / res2: their cluster ids 1..K (clustering_list[[b]]).  Returns the max(res1) x max(res2)
    matrix of block medians (NaN for a cluster pair without cells, NaNs inside a block ignored like na.rm = TRUE)."""
    ctx = ctx or default_context()
    res1, res2 = np.asarray(res1, dtype=np.int64), np.asarray(res2, dtype=np.int64)
    cells1, cells2 = np.asarray(cells1, dtype=np.int64), np.asarray(cells2, dtype=np.int64)
    K1, K2 = int(res1.max()), int(res2.max())
    # cells ordered by cluster: every (cluster, cluster) block of the distance matrix becomes a rectangle
    o1, o2 = np.argsort(res1, kind="stable"), np.argsort(res2, kind="stable")
    roff = np.concatenate([[0], np.cumsum(np.bincount(res1 - 1, minlength=K1))]).astype(np.int64)
    coff = np.concatenate([[0], np.cumsum(np.bincount(res2 - 1, minlength=K2))]).astype(np.int64)
    X = np.asarray(exprs_mat, dtype=np.float64)
    dA, dB = _cells_by_genes(ctx, X[:, cells1[o1]]), _cells_by_genes(ctx, X[:, cells2[o2]])
    med = np.empty(K1 * K2)
    try:
        d_D = _pair_dist_device(ctx, dA, dB, dist)
        L.check(L.lib().scm_block_median(ctx.handle, d_D.ptr, dB.m, roff.ctypes.data_as(C.c_void_p), K1, coff.ctypes.data_as(C.c_void_p),
                                         K2, med.ctypes.data_as(C.c_void_p)))
        d_D.free()
    finally:
        dA.free()
        dB.free()
    return med.reshape(K1, K2)


def run_pca_scores(x, rank: int = 10, ctx: Context | None = None) -> np.ndarray:
    """BiocSingular::runPCA(x, rank, scale = TRUE, center = TRUE)$x for a cells x genes matrix: centred Gram (scm_centred_gram) +
    top-`rank` eigenpairs of the correlation matrix + one projection pass (scm_pca_scores with no adjustment).  Columns are defined
    up to sign.  A zero-variance gene is dropped from the PCA (the reference would produce NaN)."""
    ctx = ctx or default_context()
    lib = L.lib()
    dX = x if isinstance(x, DeviceMatrix) else DeviceMatrix.from_host(ctx, np.asarray(x, dtype=np.float64))
    n, m = dX.n, dX.m
    r = int(min(rank, n, m - 1))
    d_gram, d_mean, d_scores = ctx.empty((n, n)), ctx.empty((n,)), ctx.empty((m, r))
    mt = C.c_int64(0)
    iters, resid = C.c_int32(0), C.c_double(0.0)
    try:
        L.check(lib.scm_centred_gram(ctx.handle, dX.ptr, m, n, dX.ld, d_gram.ptr, d_mean.ptr, C.byref(mt)))
        rc = lib.scm_pca_scores(ctx.handle, d_gram.ptr, n, mt.value, None, d_mean.ptr, r, dX.ptr, m, dX.ld, d_scores.ptr, None, 1e-9, 500,
                                C.byref(iters), C.byref(resid))
        if rc == L.ERR_NOCONV:
            warnings.warn("runPCA: " + lib.scm_last_error().decode(), L.NotConvergedWarning, stacklevel=2)
        else:
            L.check(rc)
        out = d_scores.to_host()
    finally:
        for a in (d_gram, d_mean, d_scores):
            a.free()
        if dX is not x:
            dX.free()
    return out


def computePCA_byHVGMarker(this_batch, batch, batch_oneType, marker, exprs_mat, HVG_list, ctx: Context | None = None):
    """R/scReplicate.R:793-819.  exprs_mat: genes x cells; `marker` / HVG_list[this_batch]: 0-based gene indices; returns the
    cells-of-the-batch x 10 score matrix, or None when the batch has kmeansK == 1 (`this_batch in batch_oneType`)."""
    if this_batch in batch_oneType:
        return None
    ctx = ctx or default_context()
    genes = np.asarray(marker if marker is not None else HVG_list[this_batch], dtype=np.int64)
    cells = np.nonzero(np.asarray(batch) == this_batch)[0]
    dX = _cells_by_genes(ctx, exprs_mat, genes=genes, cells=cells)
    try:
        return run_pca_scores(dX, rank=10, ctx=ctx)
    finally:
        dX.free()


def kmeans(x, centers: int, iter_max: int = 100, nstart: int = 1, seed: int = 1, ctx: Context | None = None):
    """stats::kmeans(x, centers = K, nstart = nstart) the way identifyCluster calls it (R/scReplicate.R:394-395: the first 10 PCs
    of a batch, K = kmeansK[j], nstart = 1000), on the device (scm_kmeans).  x: points x coordinates (numpy or DeviceMatrix, at
    most 16 coordinates).  Every start draws K distinct rows, the start with the smallest total within-cluster sum of squares
    wins and is finished with Hartigan-Wong transfers, so the partition is stable under the reference's own rule.  R's RNG
    stream is not reproduced (the reference's clustering is seed-dependent itself): `seed` selects the starts, the result is
    deterministic for a given seed.  `iter_max` bounds the Lloyd iterations of one start (R's iter.max = 10 counts
    Hartigan-Wong passes, a different unit).  Returns R's fields: cluster (1-based), centers, withinss, tot_withinss, size,
    iter, ifault (0; 2 = iteration / transfer cap hit, R's "did not converge")."""
    ctx = ctx or (x.ctx if isinstance(x, DeviceMatrix) else default_context())
    K = int(centers)
    dX = x if isinstance(x, DeviceMatrix) else DeviceMatrix.from_host(ctx, np.asarray(x, dtype=np.float64))
    m, d = dX.m, dX.n
    if K > m:
        raise ValueError("more cluster centers than data points")  # stats::kmeans, same condition on distinct points
    d_label, d_cent = ctx.empty((m,), np.int32), ctx.empty((K, d))
    wss, size = np.empty(K), np.empty(K, dtype=np.int64)
    tot, info = C.c_double(0.0), np.zeros(4, dtype=np.int32)
    try:
        L.check(L.lib().scm_kmeans(ctx.handle, dX.ptr, m, d, dX.ld, K, int(nstart), int(iter_max), int(seed) & (2 ** 64 - 1), d_label.ptr,
                                   d_cent.ptr, wss.ctypes.data_as(C.c_void_p), size.ctypes.data_as(C.c_void_p), C.byref(tot),
                                   info.ctypes.data_as(C.c_void_p)))
        label, cent = d_label.to_host(), d_cent.to_host()
    finally:
        d_label.free()
        d_cent.free()
        if dX is not x:
            dX.free()
    if (size == 0).any():
        raise RuntimeError("empty cluster: try a better set of initial centers")  # stats::kmeans' message (ifault 1)
    ifault = 2 if (info[1] >= iter_max or info[3]) else 0
    if ifault:
        warnings.warn(f"did not converge in {iter_max} iterations", stacklevel=2)
    return {"cluster": label.astype(np.int64) + 1, "centers": cent, "withinss": wss, "tot_withinss": tot.value, "size": size,
            "iter": int(info[1]), "ifault": ifault, "start": int(info[0]), "transfers": int(info[2])}


def identifyCluster(exprs_mat, batch, marker, HVG_list, kmeansK, nstart: int = 1000, seed: int = 1, ctx: Context | None = None):
    """identifyCluster (R/scReplicate.R:330-447).  exprs_mat: genes x cells; batch: label per cell; marker / HVG_list[b]: 0-based
    gene indices; kmeansK: K per batch in the order of unique(batch) (first appearance).  Per batch: runPCA (rank 10) on its genes ->
    kmeans(first 10 PCs, K, nstart) -> centroidDist of every cluster on the same genes; a batch with K == 1 is one cluster.
    Returns (clustering_list, clustering_distProp): per batch the cluster id (1-based) and the rank fraction of every cell."""
    ctx = ctx or default_context()
    batch = np.asarray(batch)
    _, first = np.unique(batch, return_index=True)
    batch_list = batch[np.sort(first)]  # unique(batch): order of first appearance (R/scReplicate.R:361)
    if len(kmeansK) != batch_list.shape[0]:
        raise ValueError("length of kmeansK needs to be the same as the number of batch")
    one_type = [b for b, K in zip(batch_list, kmeansK) if K == 1]
    clustering, dist_prop = [], []
    X = np.asarray(exprs_mat, dtype=np.float64)
    for j, b in enumerate(batch_list):
        cells = np.nonzero(batch == b)[0]
        genes = np.asarray(marker if marker is not None else HVG_list[b], dtype=np.int64)
        pca = computePCA_byHVGMarker(b, batch, one_type, marker, X, HVG_list, ctx=ctx)
        if pca is None:
            lab = np.ones(cells.shape[0], dtype=np.int64)
        else:
            lab = kmeans(pca[:, :10], int(kmeansK[j]), nstart=nstart, seed=seed + j, ctx=ctx)["cluster"]
        clustering.append(lab)
        dist_prop.append(cluster_centroid_dist(X[np.ix_(genes, cells)], lab, ctx=ctx)[0])
    return clustering, dist_prop
