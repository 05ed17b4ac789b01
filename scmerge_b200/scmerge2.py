"""The arithmetic core of scMerge2() (R/scMerge2.R:193-314) on the device: cosine normalisation of the cells,
pseudobulk construction, RUV-III fit on the pseudobulks and adjustment of every cell.

What stays outside (out of scope): HVG selection, SNN/Louvain clustering within batches and the
mutual-nearest-cluster matching that turns pseudobulks into replicate groups.  Their OUTPUT enters here as plain
integer labels: `clust` (cluster / cell type of every cell, what `group_wtihin_batch` returns) and
`replicate_of` (a function giving every pseudobulk its replicate id; default: the cluster label, i.e. the
supervised case where clusters with the same label are replicates of each other across batches).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .device import Context, DeviceMatrix, default_context
from .ruv import (_adjust, _device_rows_cap, _fit, _is_exact, _svd_k, column_means, genes_by_name, replicate_keys,
                  segmented_sums, tological)

__all__ = ["cosineNorm", "csc_to_device", "DeviceCSC", "pseudobulk_keys", "scMerge2_core"]


def cosineNorm(dY: DeviceMatrix, out: DeviceMatrix | None = None, min_norm: float = 1e-8) -> DeviceMatrix:
    """batchelor::cosineNorm (R/scMerge2.R:197) on a cells x genes device matrix; in place unless `out` is given."""
    out = out or dY
    L.check(L.lib().scm_cosine_rows(dY.ctx.handle, dY.ptr, dY.m, dY.n, dY.ld, out.ptr, out.ld, float(min_norm)))
    return out


def _is_sparse(x) -> bool:
    return hasattr(x, "indptr") and hasattr(x, "indices") and hasattr(x, "data") and hasattr(x, "shape")


class DeviceCSC:
    """The three dgCMatrix slots (@p widened to int64, @i, @x) of a cells-compressed expression matrix, resident in HBM."""

    def __init__(self, ctx: Context, indptr, indices, data, m: int, n: int):
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        if indptr.shape[0] != m + 1:
            raise ValueError("indptr must have one entry per cell plus one")
        self.ctx, self.m, self.n = ctx, int(m), int(n)
        self.nnz = int(indptr[-1]) if m > 0 else 0
        self.indptr = ctx.to_device(indptr)
        self.indices = ctx.to_device(np.ascontiguousarray(indices[:self.nnz], dtype=np.int32))
        self.data = ctx.to_device(np.ascontiguousarray(data[:self.nnz], dtype=np.float64))
        self._canonical = None

    @classmethod
    def from_scipy(cls, ctx: Context, X) -> "DeviceCSC":
        """`X`: scipy.sparse CSR of cells x genes or, equivalently, CSC of genes x cells (the reference's dgCMatrix)."""
        fmt = getattr(X, "format", None)
        if fmt == "csc":  # genes x cells compressed by column == cells x genes compressed by row
            n, m = X.shape
        elif fmt == "csr":
            m, n = X.shape
        else:
            raise TypeError("expected a scipy.sparse csr (cells x genes) or csc (genes x cells) matrix")
        if hasattr(X, "has_canonical_format") and not X.has_canonical_format:
            X = X.copy()
            X.sum_duplicates()
        return cls(ctx, X.indptr, X.indices, X.data, m, n)

    def to_dense(self, cosine: bool = True, min_norm: float = 1e-8, out: DeviceMatrix | None = None,
                 check_finite: bool = True) -> DeviceMatrix:
        """K8 (scm_csc_to_dense): dense cell rows on the device, fused with batchelor::cosineNorm when `cosine`."""
        ctx = self.ctx
        out = out or DeviceMatrix(ctx, self.m, self.n)
        if (out.m, out.n) != (self.m, self.n):
            raise ValueError("out has the wrong shape")
        d_flag = ctx.to_device(np.zeros(1, dtype=np.int32)) if check_finite else None
        try:
            L.check(L.lib().scm_csc_to_dense(ctx.handle, self.indptr.ptr, self.indices.ptr, self.data.ptr, self.m, self.n, self.nnz,
                                             int(bool(cosine)), float(min_norm), out.ptr, out.ld, None,
                                             d_flag.ptr if d_flag is not None else None))
            if d_flag is not None and int(d_flag.to_host()[0]):
                raise L.NonFiniteError(L.ERR_NONFINITE, "Y contains missing values or infinities.  This is not supported.")
        finally:
            if d_flag is not None:
                d_flag.free()
        return out

    def is_canonical(self) -> bool:
        """scm_csc_check: raises on a corrupt index structure; True when every cell's gene indices are strictly increasing
        (sorted, no duplicates -- every dgCMatrix), which the sparse-native kernels require."""
        if self._canonical is None:
            flag = C.c_int32(0)
            L.check(L.lib().scm_csc_check(self.ctx.handle, self.indptr.ptr, self.indices.ptr, self.m, self.n, self.nnz, C.byref(flag)))
            self._canonical = bool(flag.value)
        return self._canonical

    def row_scales(self, cosine: bool = True, min_norm: float = 1e-8, check_finite: bool = True):
        """scm_csc_rownorm: device vector of 1 / max(||x_i||, min_norm) (all ones when not `cosine`)."""
        ctx = self.ctx
        d_inv = ctx.empty((self.m,))
        d_flag = ctx.to_device(np.zeros(1, dtype=np.int32)) if check_finite else None
        L.check(L.lib().scm_csc_rownorm(ctx.handle, self.indptr.ptr, self.data.ptr, self.m, int(bool(cosine)), float(min_norm), d_inv.ptr,
                                        d_flag.ptr if d_flag is not None else None))
        if d_flag is not None:
            bad = int(d_flag.to_host()[0])
            d_flag.free()
            if bad:
                d_inv.free()
                raise L.NonFiniteError(L.ERR_NONFINITE, "Y contains missing values or infinities.  This is not supported.")
        return d_inv

    def free(self):
        for a in (self.indptr, self.indices, self.data):
            a.free()


def csc_to_device(ctx: Context, X, cosine: bool = True, min_norm: float = 1e-8, out: DeviceMatrix | None = None,
                  check_finite: bool = True) -> DeviceMatrix:
    """Sparse ingest (K8, scm_csc_to_dense).  `X` is the expression matrix compressed by CELL: a scipy.sparse CSR matrix of
    cells x genes, or equivalently the CSC of genes x cells -- the dgCMatrix scMerge2 holds (R/scMerge2.R:164-166), whose
    (@p, @i, @x) slots are exactly (indptr, indices, data).  Only the three sparse arrays cross PCIe; the dense cell rows
    are produced on the device, fused with batchelor::cosineNorm when `cosine` (R/scMerge2.R:193-198)."""
    d = DeviceCSC.from_scipy(ctx, X)
    try:
        return d.to_dense(cosine=cosine, min_norm=min_norm, out=out, check_finite=check_finite)
    finally:
        d.free()


def pseudobulk_keys(batch, clust, k_fold: int = 30, which=None, seed=None, ctx=None):
    """Composite pseudobulk key of every cell, mirroring create_pseudoBulk applied batch by batch
    (R/scMerge2_utilis.R:259-263,461-486): within batch b, cvTools::cvFolds assigns fold f = 1..K (K = min(n_b, k_fold))
    to a random order of the cells (`which = rep(1:K, length.out = n_b)`); the pseudobulk of (b, f, t) is the mean of the
    cells of cluster t in that fold.  Keys are ordered batch-major, then fold, then cluster level, which is the row
    order of the reference's `bulkExprs`.  `which` (fold id per cell, 1-based) may be passed to reproduce an R draw.
    `ctx` with several ranks: batch / cluster levels and the set of populated (batch, fold, cluster) combinations are taken
    over ALL ranks, so keys, P and names are the same everywhere; folds are drawn among the LOCAL cells of a batch (each rank
    deals its cells of batch b round-robin over the K folds in a random order -- the union is again a partition of the batch
    into K folds of equal size up to the number of ranks).
    Returns (keys int32, P, names 'i_t_f' (i = 1-based batch index), cluster level index per pseudobulk, batch index per pseudobulk)."""
    multi = ctx is not None and getattr(ctx, "world", 1) > 1
    batch, clust = np.asarray(batch), np.asarray(clust)
    if multi:
        from .ruv import _global_levels
        blev = _global_levels(ctx, np.unique(batch))
        tlev = _global_levels(ctx, np.unique(clust))
        b, t = np.searchsorted(blev, batch), np.searchsorted(tlev, clust)
    else:
        blev, b = np.unique(batch, return_inverse=True)
        tlev, t = np.unique(clust, return_inverse=True)
    m, T = b.shape[0], tlev.shape[0]
    if which is None:
        rng = np.random.default_rng(seed if not multi or seed is None else [int(seed), int(ctx.rank)])
        which = np.empty(m, dtype=np.int64)
        for bi in range(blev.shape[0]):
            idx = np.nonzero(b == bi)[0]
            if idx.shape[0] == 0:
                continue
            K = min(idx.shape[0], int(k_fold))
            order = rng.permutation(idx.shape[0])
            w = np.empty(idx.shape[0], dtype=np.int64)
            w[order] = (np.arange(idx.shape[0]) % K) + 1
            which[idx] = w
    which = np.asarray(which, dtype=np.int64)
    K = int(k_fold)
    raw = (b.astype(np.int64) * K + (which - 1)) * T + t  # dense id over (batch, fold, cluster)
    count = np.bincount(raw, minlength=blev.shape[0] * K * T).astype(np.int64)
    if multi:
        d = ctx.to_device(count)
        ctx.allreduce(d)
        count = d.to_host()
        d.free()
    present = count > 0  # droplevels(): empty combinations give no row
    remap = np.cumsum(present) - 1
    keys = remap[raw].astype(np.int32)
    ids = np.nonzero(present)[0]
    pb_t, pb_f, pb_b = ids % T, (ids // T) % K, ids // (T * K)
    # rownames(res) <- paste(i, rownames(res), sep = "_") with i the 1-based position in sort(unique(batch))
    # (R/scMerge2_utilis.R:266, R/scMerge2.R:139): the batch INDEX, not its label
    names = [f"{bb + 1}_{tlev[tt]}_{ff + 1}" for bb, tt, ff in zip(pb_b, pb_t, pb_f)]
    return keys, int(ids.shape[0]), names, pb_t.astype(np.int32), pb_b.astype(np.int32)


def _fit_pseudobulk(ctx: Context, dP: DeviceMatrix, rkeys, G: int, s: int, ruvK: int, exact: bool):
    """RUV-III fit on the P x n pseudobulk matrix.  Several ranks: after the global means every rank holds the SAME matrix;
    each fits its own contiguous block of rows through the ordinary multi-rank fit (group sums and Gram all-reduced), so the
    work is shared and every rank gets the identical fullalpha."""
    if getattr(ctx, "world", 1) <= 1:
        return _fit(ctx, dP, rkeys, G, s, ruvK, exact)
    P, w, r = dP.m, ctx.world, ctx.rank
    if P < w:
        raise ValueError(f"{P} pseudobulks cannot be shared by {w} ranks")
    base, rem = P // w, P % w
    lo = r * base + min(r, rem)
    hi = lo + base + (1 if r < rem else 0)

    class _Rows:  # rows [lo, hi) of dP, in place
        pass
    v = _Rows()
    v.ctx, v.m, v.n, v.ld = ctx, hi - lo, dP.n, dP.ld
    v.ptr = C.c_void_p(dP.buf.address + lo * dP.ld * 8)
    return _fit(ctx, v, np.asarray(rkeys)[lo:hi], G, s, ruvK, exact)


def _adjust_to_host(ctx: Context, dY: DeviceMatrix, fullalpha, k: int, ctl_mask, center, host_out: np.ndarray) -> np.ndarray:
    """scm_coef_create + scm_adjust_to_host: adjust the resident matrix in place chunk by chunk, downloads overlapped."""
    lib = L.lib()
    n = dY.n
    assert host_out.shape == (dY.m, n) and host_out.dtype == np.float64 and host_out.flags.c_contiguous
    kk = int(min(k, fullalpha.shape[0]))
    d_alpha = ctx.to_device(np.ascontiguousarray(fullalpha[:kk], dtype=np.float64))
    d_ctl = ctx.to_device(np.ascontiguousarray(ctl_mask, dtype=np.uint8))
    d_cen = ctx.to_device(np.ascontiguousarray(center, dtype=np.float64))
    coef = C.c_void_p()
    try:
        L.check(lib.scm_coef_create(ctx.handle, d_alpha.ptr, n, kk, d_ctl.ptr, d_cen.ptr, None, None, C.byref(coef)))
        L.check(lib.scm_adjust_to_host(ctx.handle, dY.ptr, dY.m, n, dY.ld, coef, host_out.ctypes.data_as(C.c_void_p), n))
    finally:
        if coef:
            lib.scm_coef_destroy(ctx.handle, coef)
        for a in (d_alpha, d_ctl, d_cen):
            a.free()
    return host_out


def _scmerge2_sparse(ctx: Context, dcsc: "DeviceCSC", batch, clust, ctl, ruvK: int, k_pseudoBulk, cosine, BSPARAM, which, seed,
                     replicate_of, pb, return_bulk, out: DeviceMatrix | None, host_out, check_finite=True, to_host=False):
    """scMerge2 core on the sparse rows themselves (scm_csc_rownorm / scm_csc_seg_colmean / scm_csc_adjust[_to_host])."""
    lib = L.lib()
    m, n = dcsc.m, dcsc.n
    ctl_mask = np.ones(n, dtype=bool) if ctl is None else tological(ctl, n)
    if pb is None:
        pb = pseudobulk_keys(batch, clust, k_pseudoBulk, which=which, seed=seed, ctx=ctx)
    keys, P, names, pb_type, _pb_batch = pb
    d_inv = dcsc.row_scales(cosine=cosine, check_finite=check_finite)
    dP = DeviceMatrix(ctx, P, n)
    d_keys, d_cnt = ctx.to_device(np.ascontiguousarray(keys, dtype=np.int32)), ctx.empty((P,), np.int64)
    coef = None
    side = None  # the whole-spectrum fit running beside the download (see below)
    try:
        L.check(lib.scm_csc_seg_colmean(ctx.handle, dcsc.indptr.ptr, dcsc.indices.ptr, dcsc.data.ptr, d_inv.ptr, m, n, d_keys.ptr, P,
                                        dP.ptr, dP.ld, d_cnt.ptr))
        d_mean = ctx.empty((n,))
        L.check(lib.scm_group_colmean(ctx.handle, dP.ptr, d_cnt.ptr, P, n, dP.ld, d_mean.ptr))
        means = d_mean.to_host()
        d_mean.free()
        bulk = dP.to_host() if (return_bulk or replicate_of is not None) else None
        rep = pb_type if replicate_of is None else np.asarray(replicate_of(names, bulk))
        rkeys, G = replicate_keys(rep)
        exact = _is_exact(BSPARAM) if BSPARAM is not None else False
        s = _svd_k(P, G, int(ctl_mask.sum()), exact, ruvK)
        if G >= P or ruvK == 0:
            raise ValueError("replicate structure leaves no residual degrees of freedom (ncol(M) >= number of pseudobulks)")
        s_dev, exact_dev = _device_rows_cap(s, ruvK, exact, n, rows=P if getattr(ctx, "world", 1) <= 1 else None)
        from .ruv import _Coef
        if (host_out is not None and s_dev > L.MAX_TOPK_ROWS and getattr(ctx, "world", 1) <= 1
                and os.environ.get("SCM_SCM2_OVERLAP", "1") != "0"):
            # The cells only need the first ruvK rows of fullalpha; the hundreds of further rows the reference's rule returns come
            # from the whole-spectrum solve (~110 ms at 2000 genes).  With a host result the adjust is a ~300 ms PCIe download, so
            # the top rows are fitted first (subspace iteration, ~2 ms), the cells start leaving at once, and the whole-spectrum
            # solve runs meanwhile on a second context (own stream, own host thread) of the same device.
            import threading
            fit_top = _fit(ctx, dP, rkeys, G, min(s, max(int(ruvK), 50), L.MAX_TOPK_ROWS), ruvK, True)
            ctx.sync()
            aux = getattr(ctx, "_aux_ctx", None)
            if aux is None:
                aux = ctx._aux_ctx = Context(ctx.device, background=True)  # least-priority stream: yields the SMs to the adjust

            class _View:  # dP as seen from the second context (same device memory)
                pass
            v = _View()
            v.ctx, v.m, v.n, v.ld, v.ptr = aux, dP.m, dP.n, dP.ld, dP.ptr
            side = {}

            def _whole_spectrum():
                try:
                    side["fit"] = _fit(aux, v, rkeys, G, s_dev, ruvK, exact_dev)
                    aux.sync()
                except BaseException as exc:  # noqa: BLE001 -- re-raised in the calling thread
                    side["error"] = exc
            side["thread"] = threading.Thread(target=_whole_spectrum, name="scm-whole-spectrum")
            side["thread"].start()
            fit = fit_top
        else:
            fit = _fit_pseudobulk(ctx, dP, rkeys, G, s_dev, ruvK, exact_dev)
        with _Coef(ctx, n, fit.fullalpha, ruvK, ctl_mask, center=means) as coef:
            if host_out is not None:
                assert host_out.shape == (m, n) and host_out.dtype == np.float64 and host_out.flags.c_contiguous
                L.check(lib.scm_csc_adjust_to_host(ctx.handle, dcsc.indptr.ptr, dcsc.indices.ptr, dcsc.data.ptr, d_inv.ptr, m, n,
                                                   coef.handle, host_out.ctypes.data_as(C.c_void_p), n))
                newY = host_out
            else:
                dout = out or DeviceMatrix(ctx, m, n)
                L.check(lib.scm_csc_adjust(ctx.handle, dcsc.indptr.ptr, dcsc.indices.ptr, dcsc.data.ptr, d_inv.ptr, m, n, coef.handle,
                                           dout.ptr, dout.ld))
                ctx.sync()
                newY = dout
                if to_host:  # host input without a caller-supplied buffer: hand back numpy like every other host path
                    newY = dout.to_host()
                    dout.free()
        if side is not None:  # every row of fullalpha from the whole-spectrum solve that ran beside the download
            side["thread"].join()
            side.pop("thread")
            if "error" in side:
                raise side["error"]
            fit = side["fit"]
    finally:
        if side is not None and "thread" in side:  # an error on the main path: the side fit still reads dP
            side["thread"].join()
        for a in (d_inv, d_keys, d_cnt):
            a.free()
        dP.free()
    return {"newY": newY, "fullalpha": fit.fullalpha, "bulkExprs": bulk, "pseudobulk_names": names, "replicate_vector": rep,
            "adjusted_means": means}


def scMerge2_core(exprs, batch, clust, ctl=None, ruvK: int = 20, k_pseudoBulk: int = 30, cosine: bool = True, BSPARAM=None,
                  which=None, seed=1, replicate_of=None, return_subset_genes=None, ctx: Context | None = None,
                  pb=None, return_bulk=True, out: DeviceMatrix | None = None, sparse_native: bool = True, gene_names=None):
    """cosineNorm -> construct_pseudoBulk -> pseudoRUVIII (R/scMerge2.R:193-309) for cells x genes `exprs`
    (numpy: copied to the device; scipy.sparse CSR cells x genes / CSC genes x cells, i.e. the reference's dgCMatrix:
    only the sparse arrays are uploaded and the dense normalised rows are built on the device (K8);
    DeviceMatrix: normalised IN PLACE, result returned as a DeviceMatrix).  For host inputs `out` may be a (page-locked)
    numpy cells x genes array: the adjusted chunks are then downloaded while the next chunk is computed.
    Sparse input with canonical cells (sorted unique indices: every dgCMatrix) takes the sparse-native kernels
    (`sparse_native`): norms, pseudobulk means and the rank-k update read 12 bytes per non-zero, the dense matrix is only
    ever written (newY), and with a host `out` it never exists on the device at all.
    With `out` (a DeviceMatrix of the same shape) a device-resident input is left untouched: the normalised cells are
    written to `out` and adjusted there.  `pb` may carry a precomputed `pseudobulk_keys(...)` result (the key construction is host-side index work).
    `ctl` / `return_subset_genes` may be masks, indices or gene names (with `gene_names` = rownames of the reference's
    genes x cells matrix; `rownames(exprsMat) %in% ctl`, R/scMerge2.R:304).
    Returns {'newY', 'fullalpha', 'bulkExprs' (P x genes), 'pseudobulk_names', 'replicate_vector', 'adjusted_means'}."""
    if gene_names is not None:
        n_genes = len(gene_names)
        ctl = None if ctl is None else genes_by_name(ctl, n_genes, gene_names)
        return_subset_genes = None if return_subset_genes is None else genes_by_name(return_subset_genes, n_genes, gene_names)
    ctx = ctx or (exprs.ctx if isinstance(exprs, (DeviceMatrix, DeviceCSC)) else default_context())
    host_in = not isinstance(exprs, (DeviceMatrix, DeviceCSC))
    host_out = out if isinstance(out, np.ndarray) else None
    if host_out is not None and isinstance(exprs, DeviceMatrix):
        raise TypeError("a numpy `out` needs a host or DeviceCSC input (a DeviceMatrix input is adjusted on the device)")
    if host_in or host_out is not None:
        out = None
    if isinstance(exprs, DeviceCSC) or _is_sparse(exprs):
        dcsc = exprs if isinstance(exprs, DeviceCSC) else DeviceCSC.from_scipy(ctx, exprs)
        try:
            if sparse_native and ruvK <= 32 and return_subset_genes is None and dcsc.is_canonical():
                # the dense normalised matrix is never an input: norms, pseudobulks and the rank-k update read the sparse rows
                return _scmerge2_sparse(ctx, dcsc, batch, clust, ctl, int(ruvK), k_pseudoBulk, cosine, BSPARAM, which, seed,
                                        replicate_of, pb, return_bulk, out if not host_in else None, host_out,
                                        check_finite=host_in, to_host=host_in)
            dY = dcsc.to_dense(cosine=cosine, out=out if not host_in else None, check_finite=host_in)
        finally:
            if dcsc is not exprs:
                dcsc.free()
        cosine, out = False, None  # already applied during the densification
    else:
        dY = DeviceMatrix.from_host(ctx, np.asarray(exprs, dtype=np.float64)) if host_in else exprs
    n = dY.n
    ctl_mask = np.ones(n, dtype=bool) if ctl is None else tological(ctl, n)  # scMerge2's default ctl = all genes
    if out is not None and not host_in:
        if cosine:
            cosineNorm(dY, out)
        else:
            L.check(L.lib().scm_copy2d(ctx.handle, out.ptr, out.ld * 8, dY.ptr, dY.ld * 8, n * 8, dY.m, 2))
        dY = out
    elif cosine:
        cosineNorm(dY)
    if pb is None:
        pb = pseudobulk_keys(batch, clust, k_pseudoBulk, which=which, seed=seed, ctx=ctx)
    keys, P, names, pb_type, _pb_batch = pb
    # K1 + device division straight into a device-resident P x genes matrix (the fit never leaves the GPU)
    dP = DeviceMatrix(ctx, P, n)
    d_keys, d_cnt = ctx.to_device(np.ascontiguousarray(keys, dtype=np.int32)), ctx.empty((P,), np.int64)
    assert dP.ld == n or n % 2 == 1
    tmp = dP if dP.ld == n else DeviceMatrix(ctx, P, n, ld=n)
    L.check(L.lib().scm_seg_colmean(ctx.handle, dY.ptr, dY.m, n, dY.ld, d_keys.ptr, P, tmp.ptr, d_cnt.ptr, None))
    if tmp is not dP:  # odd n: re-pitch to the even leading dimension
        L.check(L.lib().scm_copy2d(ctx.handle, dP.ptr, dP.ld * 8, tmp.ptr, n * 8, n * 8, P, 2))
        tmp.free()
    d_keys.free()
    # colMeans(Y) of all cells from the pseudobulk means and counts (no second pass over Y); skipped keys would break it
    d_mean = ctx.empty((n,))
    L.check(L.lib().scm_group_colmean(ctx.handle, dP.ptr, d_cnt.ptr, P, n, dP.ld, d_mean.ptr))
    means = d_mean.to_host()
    d_mean.free()
    d_cnt.free()
    bulk = dP.to_host() if (return_bulk or replicate_of is not None) else None
    rep = pb_type if replicate_of is None else np.asarray(replicate_of(names, bulk))
    rkeys, G = replicate_keys(rep)
    exact = _is_exact(BSPARAM) if BSPARAM is not None else False  # scMerge2 default use_bsparam = RandomParam()
    s = _svd_k(P, G, int(ctl_mask.sum()), exact, ruvK)
    if G >= P or ruvK == 0:
        raise ValueError("replicate structure leaves no residual degrees of freedom (ncol(M) >= number of pseudobulks)")
    s_dev, exact_dev = _device_rows_cap(s, ruvK, exact, n, rows=P if getattr(ctx, "world", 1) <= 1 else None)
    fit = _fit_pseudobulk(ctx, dP, rkeys, G, s_dev, int(ruvK), exact_dev)
    dP.free()
    sub = None if return_subset_genes is None else np.nonzero(tological(return_subset_genes, n))[0]
    if host_out is not None and sub is None:
        newY = _adjust_to_host(ctx, dY, fit.fullalpha, int(ruvK), ctl_mask, means, host_out)
        dY.free()
        return {"newY": newY, "fullalpha": fit.fullalpha, "bulkExprs": bulk, "pseudobulk_names": names, "replicate_vector": rep,
                "adjusted_means": means}
    dout, _ = _adjust(ctx, dY, fit.fullalpha, int(ruvK), ctl_mask, center=means, subset=sub, out=dY if sub is None else None)
    if host_in:
        newY = dout.to_host(host_out)
        if dout is not dY:
            dout.free()
        dY.free()
    else:
        newY = dout
    return {"newY": newY, "fullalpha": fit.fullalpha, "bulkExprs": bulk, "pseudobulk_names": names, "replicate_vector": rep,
            "adjusted_means": means}
