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

import numpy as np

from . import _lib as L
from .device import Context, DeviceMatrix, default_context
from .ruv import _adjust, _device_rows_cap, _fit, _is_exact, _svd_k, column_means, replicate_keys, segmented_sums, tological

__all__ = ["cosineNorm", "pseudobulk_keys", "scMerge2_core"]


def cosineNorm(dY: DeviceMatrix, out: DeviceMatrix | None = None, min_norm: float = 1e-8) -> DeviceMatrix:
    """batchelor::cosineNorm (R/scMerge2.R:197) on a cells x genes device matrix; in place unless `out` is given."""
    out = out or dY
    L.check(L.lib().scm_cosine_rows(dY.ctx.handle, dY.ptr, dY.m, dY.n, dY.ld, out.ptr, out.ld, float(min_norm)))
    return out


def pseudobulk_keys(batch, clust, k_fold: int = 30, which=None, seed=None):
    """Composite pseudobulk key of every cell, mirroring create_pseudoBulk applied batch by batch
    (R/scMerge2_utilis.R:259-263,461-486): within batch b, cvTools::cvFolds assigns fold f = 1..K (K = min(n_b, k_fold))
    to a random order of the cells (`which = rep(1:K, length.out = n_b)`); the pseudobulk of (b, f, t) is the mean of the
    cells of cluster t in that fold.  Keys are ordered batch-major, then fold, then cluster level, which is the row
    order of the reference's `bulkExprs`.  `which` (fold id per cell, 1-based) may be passed to reproduce an R draw.
    Returns (keys int32, P, names 'b_t_f', cluster level index per pseudobulk, batch index per pseudobulk)."""
    blev, b = np.unique(np.asarray(batch), return_inverse=True)
    tlev, t = np.unique(np.asarray(clust), return_inverse=True)
    m, T = b.shape[0], tlev.shape[0]
    if which is None:
        rng = np.random.default_rng(seed)
        which = np.empty(m, dtype=np.int64)
        for bi in range(blev.shape[0]):
            idx = np.nonzero(b == bi)[0]
            K = min(idx.shape[0], int(k_fold))
            order = rng.permutation(idx.shape[0])
            w = np.empty(idx.shape[0], dtype=np.int64)
            w[order] = (np.arange(idx.shape[0]) % K) + 1
            which[idx] = w
    which = np.asarray(which, dtype=np.int64)
    K = int(k_fold)
    raw = (b.astype(np.int64) * K + (which - 1)) * T + t  # dense id over (batch, fold, cluster)
    present = np.bincount(raw, minlength=blev.shape[0] * K * T) > 0  # droplevels(): empty combinations give no row
    remap = np.cumsum(present) - 1
    keys = remap[raw].astype(np.int32)
    ids = np.nonzero(present)[0]
    pb_t, pb_f, pb_b = ids % T, (ids // T) % K, ids // (T * K)
    names = [f"{blev[bb]}_{tlev[tt]}_{ff + 1}" for bb, tt, ff in zip(pb_b, pb_t, pb_f)]
    return keys, int(ids.shape[0]), names, pb_t.astype(np.int32), pb_b.astype(np.int32)


def scMerge2_core(exprs, batch, clust, ctl=None, ruvK: int = 20, k_pseudoBulk: int = 30, cosine: bool = True, BSPARAM=None,
                  which=None, seed=1, replicate_of=None, return_subset_genes=None, ctx: Context | None = None,
                  pb=None, return_bulk=True, out: DeviceMatrix | None = None):
    """cosineNorm -> construct_pseudoBulk -> pseudoRUVIII (R/scMerge2.R:193-309) for cells x genes `exprs`
    (numpy: copied to the device; DeviceMatrix: normalised IN PLACE, result returned as a DeviceMatrix).
    With `out` (a DeviceMatrix of the same shape) a device-resident input is left untouched: the normalised cells are
    written to `out` and adjusted there.  `pb` may carry a precomputed `pseudobulk_keys(...)` result (the key construction is host-side index work).
    Returns {'newY', 'fullalpha', 'bulkExprs' (P x genes), 'pseudobulk_names', 'replicate_vector', 'adjusted_means'}."""
    ctx = ctx or (exprs.ctx if isinstance(exprs, DeviceMatrix) else default_context())
    host_in = not isinstance(exprs, DeviceMatrix)
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
        pb = pseudobulk_keys(batch, clust, k_pseudoBulk, which=which, seed=seed)
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
    s_dev, exact_dev = _device_rows_cap(s, ruvK, exact)
    fit = _fit(ctx, dP, rkeys, G, s_dev, int(ruvK), exact_dev)
    dP.free()
    sub = None if return_subset_genes is None else np.nonzero(tological(return_subset_genes, n))[0]
    dout, _ = _adjust(ctx, dY, fit.fullalpha, int(ruvK), ctl_mask, center=means, subset=sub, out=dY if sub is None else None)
    if host_in:
        newY = dout.to_host()
        if dout is not dY:
            dout.free()
        dY.free()
    else:
        newY = dout
    return {"newY": newY, "fullalpha": fit.fullalpha, "bulkExprs": bulk, "pseudobulk_names": names, "replicate_vector": rep,
            "adjusted_means": means}
