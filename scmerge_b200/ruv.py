"""Host-side mirror of scMerge's RUV-III interface (same names, argument meaning and error behaviour as the
R closures), driving the B200 kernels through the C ABI.  No numerics are done in Python: this module only
resolves arguments (replicate matrix -> int32 keys, ctl -> mask, the svd_k rule), moves buffers and calls
libscmerge_b200.

    fastRUVIII      R/fastRUVIII.R:41-111        scRUVIII        R/scRUVIII.R:41-137
    pseudoRUVIII    R/scMerge2_utilis.R:17-150   getAdjustedMat  R/scMerge2.R:368-410
    create_pseudoBulk / aggregate_matrix   R/scMerge2_utilis.R:325-346,461-486 (arithmetic only)

Matrices are numpy float64 arrays (host; copied to the GPU inside the call) or `DeviceMatrix` objects
(already resident; results are returned as `DeviceMatrix` too).  Orientation follows the reference:
`fastRUVIII`/`pseudoRUVIII` take cells x genes, `scRUVIII`/`getAdjustedMat` take genes x cells.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib as L
from .device import Context, DeviceArray, DeviceMatrix, MultiContext, default_context

__all__ = ["eta_design_rows", "FullAlpha", "fastRUVIII", "scRUVIII", "scRUVg", "genes_by_name", "pseudoRUVIII", "getAdjustedMat", "create_pseudoBulk", "aggregate_matrix",
           "replicate_matrix", "replicate_keys", "tological", "createChunk", "RUVFit", "ExactParam", "RandomParam",
           "IrlbaParam"]


# ---- BiocSingular parameter objects (only their class matters to the reference, R/fastRUVIII.R:66) --------
class ExactParam:
    exact = True


class RandomParam:
    exact = False


class IrlbaParam:
    exact = False


def _is_exact(BSPARAM) -> bool:
    if BSPARAM is None:
        return True
    if isinstance(BSPARAM, str):
        return BSPARAM.lower().startswith("exact")
    return bool(getattr(BSPARAM, "exact", False))


# ---- argument resolution (host logic, no arithmetic) ---------------------------------------------------
def replicate_matrix(a) -> np.ndarray:
    """ruv::replicate.matrix (call site R/fastRUVIII.R:48): numeric matrix -> unchanged; label vector ->
    one-hot with columns in sorted level order; 2-D labels (data.frame) -> rows pasted with '_'."""
    keys, G = replicate_keys(a)
    M = np.zeros((keys.shape[0], G))
    M[np.arange(keys.shape[0]), keys] = 1.0
    return M


def _active_ctx(Y, ctx):
    """The context a call will run on, if one already exists (never creates one: the early-outs stay device-free)."""
    if isinstance(ctx, MultiContext):
        return None  # one process holds ALL the cells: labels and m are global already
    if ctx is not None:
        return ctx
    return Y.ctx if isinstance(Y, DeviceMatrix) else None


def _rank_table(ctx, row) -> np.ndarray:
    """world x len(row) int64 table holding every rank's `row` (one sum all-reduce of one-hot rows)."""
    row = np.asarray(row, dtype=np.int64)
    if ctx is None or ctx.world <= 1:
        return row[None, :]
    if ctx.rank is None:
        raise ValueError("multi-rank host logic needs the rank: Context.set_allreduce(fn, world_size, rank) or init_nccl")
    tab = np.zeros((ctx.world, row.shape[0]), dtype=np.int64)
    tab[ctx.rank] = row
    d = ctx.to_device(tab)
    ctx.allreduce(d)
    tab = d.to_host()
    d.free()
    return tab


def _total_cells(ctx, m_local: int) -> int:
    """Cells over ALL ranks that share them (the m of the reference's formulas: svd_k rule, early-out, PCA rank)."""
    if ctx is None or ctx.world <= 1:
        return int(m_local)
    cached = getattr(ctx, "_m_total_once", None)  # left by _global_int_keys of the same call (one-shot: every rank takes the same path)
    ctx._m_total_once = None
    if cached is not None and cached[0] == int(m_local):
        return cached[1]
    return int(_rank_table(ctx, [m_local])[:, 0].sum())


def _global_levels(ctx, levels: np.ndarray) -> np.ndarray:
    """Sorted union of the label levels seen by any rank.  Shards are contiguous row blocks and data usually arrives ordered
    by batch, so a rank routinely lacks some batch / cell type: coding labels per rank would give every rank its own
    key -> group mapping (and its own G, Bt, s), and the all-reduced sums would mix different groups."""
    import pickle
    parts = ctx.allgather_bytes(pickle.dumps(np.asarray(levels)))
    got = [pickle.loads(p) for p in parts]
    got = [g for g in got if g.shape[0] > 0]
    if not got:
        return np.asarray(levels)
    return np.unique(np.concatenate(got))


def _global_int_keys(ctx, a: np.ndarray, spare):
    """Integer labels on several ranks without sorting a million of them per call: min / max / row count of every rank in one
    small all-reduce; with enough slack (see replicate_keys: `spare`) the labels ARE the keys, otherwise a presence histogram
    (O(m), summed over the ranks) gives the global level table.  Returns None when the labels are too sparse for that."""
    m = a.shape[0]
    lo, hi = (int(a.min()), int(a.max())) if m else (2 ** 62, -1)
    tab = _rank_table(ctx, [m, lo, hi])
    m_tot, lo, hi = int(tab[:, 0].sum()), int(tab[:, 1].min()), int(tab[:, 2].max())
    ctx._m_total_once = (int(m), m_tot)  # saves _total_cells its own round trip
    if hi < 0:
        return a.astype(np.int32, copy=False), 0
    if lo < 0 or hi >= 4 * m_tot + 1024:
        return None
    if spare is not None and m_tot - (hi + 1) >= spare:
        return a.astype(np.int32, copy=False), hi + 1
    present = np.bincount(a, minlength=hi + 1).astype(np.int64)
    d = ctx.to_device(present)
    ctx.allreduce(d)
    present = d.to_host() > 0
    d.free()
    if present.all():
        return a.astype(np.int32, copy=False), hi + 1
    remap = np.cumsum(present) - 1
    return remap[a].astype(np.int32), int(present.sum())


def replicate_keys(M, spare: int | None = None, ctx=None):
    """Replicate structure -> (int32 group id per cell, number of groups).
    `ctx` with several ranks: the level table is built over ALL ranks (see _global_levels), so ids and G are global.  A numeric matrix must be a one-hot
    partition (every row exactly one 1): the device path implements the residual operator as a segmented
    mean, which is what my_residop (R/fastRUVIII.R:121-129) computes for such M; a general M is rejected.
    `spare`: for non-negative integer labels with m - (max + 1) >= spare the O(m) histogram that drops unused label
    values (R's factor levels) is skipped and G = max + 1 is returned: unused ids are empty groups, which change nothing
    in the arithmetic, and with that much slack the svd_k rule min(m - G, n_ctl, cap) does not depend on G either."""
    if isinstance(M, DeviceArray):
        raise TypeError("pass replicate keys as a host array")
    a = np.asarray(M)
    if a.ndim == 2 and np.issubdtype(a.dtype, np.number):
        if a.shape[1] == 0:
            raise ValueError("replicate matrix has no columns")
        ok = np.all((a == 0) | (a == 1)) and np.all(a.sum(axis=1) == 1)
        if not ok:
            raise ValueError("M is not a one-hot replicate matrix (each row must contain exactly one 1); "
                             "the device path does not support a general projector")
        return np.argmax(a, axis=1).astype(np.int32), int(a.shape[1])
    if a.ndim == 2:
        a = np.array(["_".join(str(x) for x in row) for row in a])
    if a.ndim != 1:
        raise ValueError("M must be a replicate matrix or a label vector")
    if ctx is not None and ctx.world > 1:
        if np.issubdtype(a.dtype, np.integer):
            got = _global_int_keys(ctx, a, spare)
            if got is not None:
                return got
        levels = _global_levels(ctx, np.unique(a))
        return np.searchsorted(levels, a).astype(np.int32), int(levels.shape[0])
    if np.issubdtype(a.dtype, np.integer) and a.shape[0] > 0:
        lo, hi = int(a.min()), int(a.max())
        if spare is not None and lo >= 0 and a.shape[0] - (hi + 1) >= spare:
            return a.astype(np.int32, copy=False), hi + 1
        if lo >= 0 and hi < 4 * a.shape[0] + 1024:  # O(m) path for integer labels (sorted level order = numeric order)
            present = np.bincount(a, minlength=hi + 1) > 0
            if present.all():
                return a.astype(np.int32, copy=False), hi + 1
            remap = np.cumsum(present) - 1
            return remap[a].astype(np.int32), int(present.sum())
    levels, inv = np.unique(a, return_inverse=True)
    return inv.astype(np.int32), int(levels.shape[0])


def tological(ctl, n: int) -> np.ndarray:
    """R/fastRUVIII.R:113-117 (0-based integer indices or a logical mask)."""
    ctl = np.asarray(ctl)
    out = np.zeros(n, dtype=bool)
    if ctl.dtype == bool:
        out[: ctl.shape[0]] = ctl[:n]
    else:
        out[ctl.astype(np.int64)] = True
    return out


def createChunk(n: int, chunkSize: int = 1000):
    """R/scMerge2_utilis.R:153-170 (1-based inclusive ranges).  Kept for interface parity: the fused adjust
    kernel makes chunking invisible, so pseudoRUVIII accepts byChunk/chunkSize but does not need them."""
    starts = list(range(1, n + 1, chunkSize))
    ends = [s + chunkSize - 1 for s in starts]
    ends[-1] = n
    if len(ends) > 1 and ends[-1] - ends[-2] < chunkSize / 2:
        starts, ends = starts[:-1], ends[:-1]
        ends[-1] = n
    return list(zip(starts, ends))


def _svd_k(m, G, n_ctl, exact, cap):
    """R/fastRUVIII.R:66-70 / R/scMerge2_utilis.R:60-65 (the non-exact branch ignores the user's svd_k)."""
    return int(min(m - G, n_ctl, cap)) if exact else int(min(m - G, n_ctl))


def eta_design_rows(eta, n: int, include_intercept: bool = True) -> np.ndarray:
    """Gene-wise covariates of ruv::RUV1 (CRAN `ruv`, un-vendored; call sites R/fastRUVIII.R:64, R/scRUVg.R:59,
    R/scMerge2_utilis.R:56) -> the q x n matrix `t(design.matrix(eta))`.  Accepted like the `ruv` documentation states: the number
    1 (an intercept), a vector of length n, a matrix with n rows (n x q) or with n columns (q x n).  `include_intercept`: a row
    of ones is put first unless the constant vector already lies in the span of the covariates (design.matrix's test:
    squared residual of 1 on eta > 1e-8).  Collinear covariates are rejected (design.matrix drops columns with a warning)."""
    a = np.asarray(eta, dtype=np.float64)
    if a.ndim == 0:
        if a != 1:
            raise ValueError("a scalar eta must be 1 (an intercept)")
        a = np.ones((n, 1))
    elif a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2 or n not in a.shape:
        raise ValueError(f"eta must have one row (or one column) per gene (n = {n})")
    if a.shape[0] != n:
        a = a.T
    if not np.all(np.isfinite(a)):
        raise ValueError("eta contains missing values or infinities")
    if include_intercept:
        one = np.ones(n)
        res = one - a @ np.linalg.lstsq(a, one, rcond=None)[0]
        if float(res @ res) > 1e-8:
            a = np.concatenate([one[:, None], a], axis=1)
    sv = np.linalg.svd(a / np.maximum(np.linalg.norm(a, axis=0), 1e-300), compute_uv=False)
    if a.shape[1] > n or sv[-1] < 1e-9 * sv[0]:
        raise ValueError("eta is collinear (ruv::design.matrix would drop columns): remove the redundant covariates")
    return np.ascontiguousarray(a.T)


def _ruv1(ctx: Context, dY: DeviceMatrix, eta, ctl: np.ndarray, include_intercept: bool, in_place: bool) -> DeviceMatrix:
    """ruv::RUV1 on the device: Y - Y[, ctl] etac' (etac etac')^-1 eta is the rank-q update of K5 + K6 with the rows of eta in
    the place of alpha (no centre, no scaling).  `in_place`: overwrite dY (a private copy), else a new device matrix."""
    rows = eta_design_rows(eta, dY.n, include_intercept)
    if rows.shape[0] > L.MAX_K:
        raise L.ScmError(L.ERR_UNSUPPORTED, f"eta with {rows.shape[0]} covariates > {L.MAX_K} is not supported by the fused adjust kernel")
    if rows.shape[0] > int(ctl.sum()):
        raise ValueError("eta has more covariates than there are control genes (etac etac' is singular)")
    out, _ = _adjust(ctx, dY, rows, rows.shape[0], ctl, out=dY if in_place else None)
    return out


FULLALPHA_ROWS = "reference"  # "truncate": keep the subspace iteration and hand back max(k, 50) rows when the reference rule gives more


def _device_rows_cap(s: int, k: int, exact: bool, n: int | None = None, rows: int | None = None):
    """How many rows of fullalpha the device fit returns, and with which solver.  Up to MAX_TOPK_ROWS rows the block subspace
    iteration serves (its Rayleigh-Ritz block is limited to MAX_BLOCK columns).  In randomized mode the reference's rule asks for
    min(m - G, n_ctl) rows (R/fastRUVIII.R:66-70, often hundreds: rsvd at nearly full rank, i.e. the exact SVD): with a Gram of
    at most MAX_FULL_EIG_N genes the device computes the WHOLE spectrum (eig_full.cuh) and returns all of them, every row
    exact.  A short matrix (`rows` < n, at most MAX_FULL_EIG_N rows, single rank: the pseudobulk fits of scMerge2) takes the
    rows x rows side instead, whatever n (fit_dual.cuh).  Larger problems (or FULLALPHA_ROWS = "truncate") fall back to
    max(k, 50) converged rows with a warning.  Returns (rows, exact)."""
    cap = L.MAX_TOPK_ROWS
    if s > cap:
        if FULLALPHA_ROWS == "reference" and n is not None and (n <= L.MAX_FULL_EIG_N or (rows is not None and rows < n and rows <= L.MAX_FULL_EIG_N)):
            return s, True
        want = min(s, max(int(k), 50), cap)
        warnings.warn(f"fullalpha truncated to {want} rows (reference rule gives {s}); the whole-spectrum solver serves Grams of "
                      f"up to {L.MAX_FULL_EIG_N} genes", stacklevel=3)
        return want, True
    return s, exact


def _as_device(ctx: Context, Y):
    if isinstance(Y, DeviceMatrix):
        return Y, False
    return DeviceMatrix.from_host(ctx, Y), True


def _dptr(a):
    return a.ptr if a is not None else None


class FullAlpha(np.ndarray):
    """fullalpha (s x n) as returned by a device fit: a plain float64 array that also remembers `conv_rank`, the number of
    leading rows the eigensolver converged as an invariant subspace (scm_ctx_fit_info).  The reference's exact SVD makes every
    row valid; here the rows beyond the ranks a fit was asked for are Ritz vectors of a subspace iteration that stopped once
    the requested ranks had converged.  Re-using the array with a larger k (fastRUVIII(fullalpha=), getAdjustedMat(ruvK=))
    than `conv_rank` raises a warning instead of silently differing from a direct fit."""

    conv_rank = None

    def __new__(cls, arr, conv_rank=None):
        obj = np.asarray(arr, dtype=np.float64).view(cls)
        obj.conv_rank = conv_rank
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self.conv_rank = getattr(obj, "conv_rank", None)

    def __reduce__(self):  # pickling keeps the rank (the fit/adjust split is the reference's checkpoint)
        fn, args, state = super().__reduce__()
        return fn, args, (state, self.conv_rank)

    def __setstate__(self, state):
        super().__setstate__(state[0])
        self.conv_rank = state[1]


def _check_conv_rank(fullalpha, k: int):
    cr = getattr(fullalpha, "conv_rank", None)
    if cr is not None and k > cr and k <= np.shape(fullalpha)[0]:
        warnings.warn(f"ruvK = {k} exceeds the rank this fullalpha was converged for ({cr}): rows {cr + 1}..{k} are unconverged Ritz "
                      "vectors, the result will differ from a direct fit with this k.  Fit with the largest k you intend to use "
                      "(or list every k in scRUVIII).", L.NotConvergedWarning, stacklevel=4)


class RUVFit:
    """Result of the fit half of RUV-III: the small, host-side 'checkpoint' (fullalpha and, for scRUVIII, the
    standardisation vectors).  `fullalpha` is s x n like the reference's (R/fastRUVIII.R:93)."""

    def __init__(self, fullalpha, sigma=None, stand_mean=None, stand_sd=None, iters=0, resid=0.0):
        self.fullalpha, self.sigma = fullalpha, sigma
        self.stand_mean, self.stand_sd = stand_mean, stand_sd
        self.iters, self.resid = iters, resid
        self.d_gram = self.d_colmean = None  # device-resident centred Gram / column means (only with keep_gram)


# ---- device drivers -----------------------------------------------------------------------------------
def _fit(ctx: Context, dY: DeviceMatrix, keys: np.ndarray, G: int, s: int, k: int, exact: bool, batch_keys=None, Bt=0,
         tol=1e-10, maxit=500, seed=1, keep_gram=False) -> RUVFit:
    """scm_ruv3_fit.  keep_gram: the centred Gram of all cells (n x n) and their column means stay on the device as
    `fit.d_gram` / `fit.d_colmean` (inputs of the ruvK selection, scm_pca_scores); the caller frees them."""
    lib = L.lib()
    n = dY.n
    d_gram = ctx.empty((n, n)) if keep_gram else None
    d_colmean = ctx.empty((n,)) if keep_gram else None
    d_keys = ctx.to_device(np.ascontiguousarray(keys, dtype=np.int32))
    d_batch = ctx.to_device(np.ascontiguousarray(batch_keys, dtype=np.int32)) if batch_keys is not None else None
    d_fa = ctx.empty((s, n))
    d_sig = ctx.empty((s,))
    d_mean = ctx.empty((n,)) if d_batch is not None else None
    d_sd = ctx.empty((n,)) if d_batch is not None else None
    iters, resid = C.c_int32(0), C.c_double(0.0)
    rc = lib.scm_ruv3_fit_gram(ctx.handle, dY.ptr, dY.m, n, dY.ld, d_keys.ptr, G, _dptr(d_batch), Bt, s, min(k, s),
                               L.MODE_EXACT if exact else L.MODE_RANDOMIZED, tol, maxit, seed, d_fa.ptr, d_sig.ptr,
                               _dptr(d_mean), _dptr(d_sd), C.byref(iters), C.byref(resid), _dptr(d_gram), _dptr(d_colmean))
    if rc == L.ERR_NOCONV:
        warnings.warn(lib.scm_last_error().decode(), L.NotConvergedWarning, stacklevel=3)
    else:
        L.check(rc)
    info = ctx.fit_info()
    fit = RUVFit(FullAlpha(d_fa.to_host(), info["conv_rank"]), d_sig.to_host(), d_mean.to_host() if d_mean is not None else None,
                 d_sd.to_host() if d_sd is not None else None, iters.value, resid.value)
    for a in (d_keys, d_batch, d_fa, d_sig, d_mean, d_sd):
        if a is not None:
            a.free()
    fit.d_gram, fit.d_colmean = d_gram, d_colmean
    return fit


class _Coef:
    """scm_coef_create / scm_coef_destroy as a context manager: the device-side (B, A, center) of
    newY = Y - ((Y - center) B) A for alpha = the first k rows of fullalpha."""

    def __init__(self, ctx: Context, n: int, fullalpha: np.ndarray, k: int, ctl: np.ndarray, center=None, pre=None, post=None):
        self.ctx, self.n = ctx, n
        self.k = int(min(k, fullalpha.shape[0]))
        _check_conv_rank(fullalpha, self.k)
        if self.k > L.MAX_K:
            raise L.ScmError(L.ERR_UNSUPPORTED, f"ruvK = {self.k} > {L.MAX_K} is not supported by the fused adjust kernel")
        self._args = (np.ascontiguousarray(fullalpha[:self.k], dtype=np.float64), np.ascontiguousarray(ctl, dtype=np.uint8),
                      center, pre, post)
        self.handle = C.c_void_p()
        self._bufs = []

    def __enter__(self):
        ctx, lib = self.ctx, L.lib()
        alpha, ctl, center, pre, post = self._args
        d_alpha, d_ctl = ctx.to_device(alpha), ctx.to_device(ctl)
        vecs = [ctx.to_device(np.ascontiguousarray(v, dtype=np.float64)) if v is not None else None for v in (center, pre, post)]
        self._bufs = [d_alpha, d_ctl] + vecs
        try:
            L.check(lib.scm_coef_create(ctx.handle, d_alpha.ptr, self.n, self.k, d_ctl.ptr, _dptr(vecs[0]), _dptr(vecs[1]),
                                        _dptr(vecs[2]), C.byref(self.handle)))
        except Exception:
            self.__exit__(None, None, None)
            raise
        return self

    def __exit__(self, *exc):
        if self.handle:
            L.lib().scm_coef_destroy(self.ctx.handle, self.handle)
            self.handle = C.c_void_p()
        for a in self._bufs:
            if a is not None:
                a.free()
        self._bufs = []
        return False


def _adjust(ctx: Context, dY: DeviceMatrix, fullalpha: np.ndarray, k: int, ctl: np.ndarray, center=None, pre=None,
            post=None, subset=None, out: DeviceMatrix | None = None, want_W=False):
    """newY = Y - ((Y - center) B) A on the device (scm_coef_create + scm_adjust)."""
    lib = L.lib()
    n = dY.n
    with _Coef(ctx, n, fullalpha, k, ctl, center, pre, post) as coef:
        d_sub, nsub = None, 0
        if subset is not None:
            sub = np.ascontiguousarray(subset, dtype=np.int32)
            if sub.ndim != 1 or (sub.size and (sub.min() < 0 or sub.max() >= n)):
                raise IndexError(f"gene subset outside [0, {n})")  # the device kernel indexes Y and A with it
            d_sub, nsub = ctx.to_device(sub), int(sub.shape[0])
        n_out = nsub if d_sub is not None else n
        if out is None:
            out = DeviceMatrix(ctx, dY.m, n_out)
        d_W = ctx.empty((dY.m, coef.k)) if want_W else None
        L.check(lib.scm_adjust(ctx.handle, dY.ptr, dY.m, n, dY.ld, coef.handle, _dptr(d_sub), nsub, out.ptr, out.ld, _dptr(d_W)))
        ctx.sync()
        W = d_W.to_host() if d_W is not None else None
        for a in (d_sub, d_W):
            if a is not None:
                a.free()
    return out, W


def _silhouette_pair(ctx: Context, fit: RUVFit, dY: DeviceMatrix, coef: _Coef | None, ct_keys: np.ndarray, n_ct: int,
                     b_keys: np.ndarray, n_b: int, rank: int = 10):
    """calculateSil (R/scRUVIII.R:182-190) for one candidate k: PC scores of the adjusted matrix (scm_pca_scores, no newY
    is materialised) and the average silhouette widths of the cell types and of the batches (scm_silhouette).
    Multi-rank: every rank computes the scores of its cells, scores and labels are gathered with the sum all-reduce
    (disjoint slices of a zero buffer), each rank takes a slice of the cells for the O(m^2) distance pass and the
    width sums are all-reduced: every rank returns the same pair."""
    from .dist import shard_rows
    lib = L.lib()
    m, n = dY.m, dY.n
    world = ctx.world
    off, m_total = 0, m
    if world > 1:
        if ctx.rank is None:
            raise ValueError("the ruvK selection on several ranks needs Context.set_allreduce(fn, world_size, rank)")
        cnt = np.zeros(world, dtype=np.int64)
        cnt[ctx.rank] = m
        d_cnt = ctx.to_device(cnt)
        ctx.allreduce(d_cnt)
        cnt = d_cnt.to_host()
        d_cnt.free()
        off, m_total = int(cnt[:ctx.rank].sum()), int(cnt.sum())
    r = int(min(rank, n, m_total - 1))
    d_scores = ctx.empty((m, r))
    iters, resid = C.c_int32(0), C.c_double(0.0)
    sa, sb = C.c_double(0.0), C.c_double(0.0)
    bufs = [d_scores]
    try:
        rc = lib.scm_pca_scores(ctx.handle, fit.d_gram.ptr, n, m_total, coef.handle if coef is not None else None, fit.d_colmean.ptr, r,
                                dY.ptr, m, dY.ld, d_scores.ptr, None, 1e-9, 500, C.byref(iters), C.byref(resid))
        if rc == L.ERR_NOCONV:
            warnings.warn("PCA for the ruvK selection: " + lib.scm_last_error().decode(), L.NotConvergedWarning, stacklevel=3)
        else:
            L.check(rc)
        ct32, b32 = np.ascontiguousarray(ct_keys, dtype=np.int32), np.ascontiguousarray(b_keys, dtype=np.int32)
        if world > 1:
            d_all = ctx.to_device(np.zeros((m_total, r)))
            bufs.append(d_all)
            if m:
                L.check(lib.scm_copy2d(ctx.handle, C.c_void_p(d_all.address + off * r * 8), r * 8, d_scores.ptr, r * 8, r * 8, m, 2))
            ctx.allreduce(d_all)
            lab = np.zeros((2, m_total), dtype=np.int64)
            lab[0, off:off + m], lab[1, off:off + m] = ct32, b32
            d_lab = ctx.to_device(lab)
            ctx.allreduce(d_lab)
            lab = d_lab.to_host()
            d_lab.free()
            ct32, b32 = lab[0].astype(np.int32), lab[1].astype(np.int32)
            i0, i1 = shard_rows(m_total, world, ctx.rank)
            d_x = d_all
        else:
            i0, i1, d_x = 0, m, d_scores
        d_ct, d_b = ctx.to_device(ct32), ctx.to_device(b32)
        bufs += [d_ct, d_b]
        L.check(lib.scm_silhouette(ctx.handle, d_x.ptr, m_total, r, r, d_ct.ptr, int(n_ct), d_b.ptr, int(n_b), i0, i1,
                                   C.byref(sa), C.byref(sb)))
        sums = np.array([sa.value, sb.value])
        if world > 1:
            d_s = ctx.to_device(sums)
            ctx.allreduce(d_s)
            sums = d_s.to_host()
            d_s.free()
    finally:
        for a in bufs:
            a.free()
    return sums[0] / m_total, sums[1] / m_total


def _is_host_matrix(Y) -> bool:
    return isinstance(Y, np.ndarray) and Y.ndim == 2 and Y.dtype == np.float64 and Y.flags.c_contiguous and Y.size > 0


def _host_pipeline(ctx: Context, Y: np.ndarray, ctl: np.ndarray, k: int, keys=None, G=0, s=0, exact=True, batch_keys=None, Bt=0,
                   fullalpha_in=None, center_mode=0, center=None, out=None, tol=1e-10, maxit=500, seed=1):
    """scm_ruv3_host: upload / fit / adjust / download of a HOST cells x genes matrix with the PCIe copies overlapped
    with the kernels.  Returns (newY, RUVFit)."""
    lib = L.lib()
    m, n = Y.shape
    if out is None:
        out = np.empty((m, n), dtype=np.float64)
    assert out.shape == (m, n) and out.dtype == np.float64 and out.flags.c_contiguous
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None  # noqa: E731
    ctl8 = np.ascontiguousarray(ctl, dtype=np.uint8)
    do_fit = fullalpha_in is None
    keys32 = np.ascontiguousarray(keys, dtype=np.int32) if do_fit else None
    b32 = np.ascontiguousarray(batch_keys, dtype=np.int32) if (do_fit and batch_keys is not None) else None
    fa_in = np.ascontiguousarray(fullalpha_in, dtype=np.float64) if not do_fit else None
    s_out = s if do_fit else fa_in.shape[0]
    if min(k, s_out) > L.MAX_K:
        raise L.ScmError(L.ERR_UNSUPPORTED, f"ruvK = {min(k, s_out)} > {L.MAX_K} is not supported by the fused adjust kernel")
    fa = np.empty((s_out, n)) if do_fit else fa_in
    sig = np.empty(s_out) if do_fit else None
    mean = np.empty(n) if (b32 is not None or center_mode == 1) else None
    sd = np.empty(n) if b32 is not None else None
    cen = np.ascontiguousarray(center, dtype=np.float64) if center is not None else None
    iters, resid = C.c_int32(0), C.c_double(0.0)
    # a MultiContext (one process, several GPUs) takes the same arguments: the library splits the cells over its devices
    multi = isinstance(ctx, MultiContext)
    call = lib.scm_multi_ruv3_host if multi else lib.scm_ruv3_host
    rc = call(ctx.handle, vp(Y), m, n, n, vp(keys32), G, vp(b32), Bt, vp(ctl8), int(k), int(s_out),
              L.MODE_EXACT if exact else L.MODE_RANDOMIZED, tol, maxit, seed, vp(fa_in), 0 if do_fit else s_out,
              center_mode, vp(cen), vp(out), n, vp(fa) if do_fit else None, vp(sig), vp(mean), vp(sd),
              C.byref(iters), C.byref(resid))
    if multi:
        ctx = ctx.ctx(0)  # fit facts are identical on every device; device 0 reports them
    if rc == L.ERR_NOCONV:
        warnings.warn(lib.scm_last_error().decode(), L.NotConvergedWarning, stacklevel=3)
    else:
        L.check(rc)
    if do_fit:
        fa = FullAlpha(fa, ctx.fit_info()["conv_rank"])
    else:
        _check_conv_rank(fullalpha_in, int(min(k, s_out)))
    return out, RUVFit(fa, sig, mean, sd, iters.value, resid.value)


def _finish(out: DeviceMatrix, made_copy_in: bool, host_out=None):
    """Return numpy when the caller gave numpy, else the DeviceMatrix."""
    if not made_copy_in:
        return out
    res = out.to_host(host_out)
    out.free()
    return res


def _scruviii_degenerate(Y, M, ks, return_all_RUV):
    """scRUVIII when fastRUVIII takes its early-out (ncol(M) >= m or k == 0, R/fastRUVIII.R:78-80): every candidate's newY is
    the standardised matrix itself, and adding the mean and sd back (R/scRUVIII.R:117-119) returns the input, cells x genes;
    fullalpha is NULL, f_score = 1 for a single k (R/scRUVIII.R:80-82).  With several k the reference would go on to the
    silhouette of identical matrices (all F-scores equal, which.max picks the first k): the first k is returned without
    touching the device."""
    Yh = Y.to_host() if isinstance(Y, DeviceMatrix) else np.ascontiguousarray(np.asarray(Y, dtype=np.float64).T)
    results = {kk: {"newY": Yh, "M": M, "fullalpha": None} for kk in (ks if return_all_RUV else ks[:1])}
    if return_all_RUV:
        results["optimal_ruvK"] = ks[0]
        return results
    one = dict(results[ks[0]])
    one["optimal_ruvK"] = ks[0]
    return one


# ---- the reference's closures ---------------------------------------------------------------------------
def fastRUVIII(Y, M, ctl, k=None, eta=None, svd_k=50, include_intercept=True, average=False, BPPARAM=None,
               BSPARAM=None, fullalpha=None, return_info=False, inputcheck=True, ctx: Context | None = None, out=None):
    """R/fastRUVIII.R:41-111.  Y: m x n (cells x genes); M: replicate matrix or labels; ctl: mask / indices.
    Returns newY (m x n), or {'newY','M','fullalpha'} with return_info.  `BPPARAM`, `average`,
    are accepted and unused like in the reference.  `eta` (gene-wise covariates, anything eta_design_rows() takes): Y is first
    replaced by ruv::RUV1(Y, eta, ctl, include_intercept) (R/fastRUVIII.R:63-64) with the same rank-k update kernel, on the
    device; the host-streaming pipeline is not used then (the matrix has to be resident for the extra pass)."""
    del BPPARAM, average
    if k is None:
        raise TypeError("k must be given (the reference fails on k = NULL, R/fastRUVIII.R:78)")
    m, n = (Y.m, Y.n) if isinstance(Y, DeviceMatrix) else np.shape(Y)
    actx = _active_ctx(Y, ctx)
    keys, G = replicate_keys(M, spare=n, ctx=actx)
    if keys.shape[0] != m:
        raise ValueError("M must have one row per row of Y")
    ctl = tological(ctl, n)
    if eta is not None:
        eta_design_rows(eta, n, include_intercept)  # argument errors before any device work
    exact = _is_exact(BSPARAM)
    m_total = _total_cells(actx, m)  # several ranks: the reference's m is the number of cells over all of them
    s = _svd_k(m_total, G, int(ctl.sum()), exact, svd_k)
    if G >= m_total or k == 0:  # R/fastRUVIII.R:78-80: RUV-III not appropriate, return the input (after RUV1, :64)
        newY, fa = Y, None
        if eta is not None:
            ectx = ctx or (Y.ctx if isinstance(Y, DeviceMatrix) else default_context())
            dY, copied = _as_device(ectx, Y)
            newY = _finish(_ruv1(ectx, dY, eta, ctl, include_intercept, in_place=copied), copied)
        return {"newY": newY, "M": M, "fullalpha": fa} if return_info else newY
    if isinstance(ctx, MultiContext) and not (_is_host_matrix(Y) and (out is None or isinstance(out, np.ndarray))):
        raise TypeError("a MultiContext drives HOST matrices (C-contiguous float64 cells x genes); device-resident shards use MultiContext.ctx(i)")
    ctx = ctx or (Y.ctx if isinstance(Y, DeviceMatrix) else default_context())
    if eta is not None and isinstance(ctx, MultiContext):
        raise L.ScmError(L.ERR_UNSUPPORTED, "eta is not supported on a MultiContext (host-streaming driver): use one context per device")
    if eta is None and _is_host_matrix(Y) and (out is None or isinstance(out, np.ndarray)):
        # host matrix: one pipelined call (H2D / fit / adjust / D2H overlapped)
        if fullalpha is None:
            s_dev, exact_dev = _device_rows_cap(s, k, exact, n)
            newY, fit = _host_pipeline(ctx, Y, ctl, int(k), keys=keys, G=G, s=s_dev, exact=exact_dev, out=out)
            fullalpha = fit.fullalpha
        else:
            fullalpha = np.asanyarray(fullalpha, dtype=np.float64)
            newY, _ = _host_pipeline(ctx, Y, ctl, int(k), fullalpha_in=fullalpha, center_mode=0, out=out)
        return {"newY": newY, "M": M, "fullalpha": fullalpha} if return_info else newY
    dY, copied = _as_device(ctx, Y)
    if eta is not None:  # R/fastRUVIII.R:63-64: Y <- RUV1(Y, eta, ctl); everything below works on that matrix
        d1 = _ruv1(ctx, dY, eta, ctl, include_intercept, in_place=copied)
        dY, copied = d1, True
    if fullalpha is None:
        # NA/Inf check (R/fastRUVIII.R:52-60) is fused into the first pass over Y; inputcheck=False only skips the raise
        del inputcheck
        s_dev, exact_dev = _device_rows_cap(s, k, exact, n)
        fit = _fit(ctx, dY, keys, G, s_dev, int(k), exact_dev)
        fullalpha = fit.fullalpha
    fullalpha = np.asanyarray(fullalpha, dtype=np.float64)
    host_result = not isinstance(Y, DeviceMatrix)
    dout, _ = _adjust(ctx, dY, fullalpha, int(k), ctl, out=out if isinstance(out, DeviceMatrix) else None)
    newY = _finish(dout, host_result, out if isinstance(out, np.ndarray) else None)
    if copied:
        dY.free()
    if not return_info:
        return newY
    return {"newY": newY, "M": M, "fullalpha": fullalpha}


def scRUVIII(Y, M, ctl, fullalpha=None, k=None, cell_type=None, batch=None, return_all_RUV=True, BPPARAM=None,
             BSPARAM=None, svd_k=50, ctx: Context | None = None):
    """R/scRUVIII.R:41-137.  Y: genes x cells (numpy) -- byte-identical to the device layout cells x genes.
    Standardisation and de-standardisation (R/scRUVIII.R:47-50,117-119) are fused into the Gram scaling and the
    adjust coefficients, so Y is read but never rewritten.  With several k the silhouette-based choice of
    R/scRUVIII.R:80-114 runs on the device (PC scores of every candidate from the centred Gram the fit already has, exact
    silhouette widths of cell types and batches; nothing is plotted): 'optimal_ruvK', 'f_score' and 'sil' (2 x len(k)) are
    returned, and with return_all_RUV=False only the optimal matrix is produced."""
    del BPPARAM
    if batch is None:
        raise ValueError("batch is required")
    ks = [int(k)] if np.isscalar(k) else [int(x) for x in k]
    actx = _active_ctx(Y, ctx)
    keys, G = replicate_keys(M, ctx=actx)
    m = Y.m if isinstance(Y, DeviceMatrix) else np.shape(Y)[1]
    if keys.shape[0] != m:
        raise ValueError("M must have one row per cell (column of Y)")
    m_total = _total_cells(actx, m)
    if G >= m_total or ks[0] == 0:  # fastRUVIII's early-out (R/fastRUVIII.R:78-80): resolved before touching the device
        return _scruviii_degenerate(Y, M, ks, return_all_RUV)
    ctx = ctx or default_context()
    if isinstance(ctx, MultiContext) and not (isinstance(Y, np.ndarray) and _is_host_matrix(Y.T) and len(ks) == 1 and fullalpha is None):
        raise TypeError("a MultiContext drives HOST matrices with a single ruvK (genes x cells in column-major order)")
    if isinstance(Y, np.ndarray) and _is_host_matrix(Y.T) and len(ks) == 1 and fullalpha is None:
        # genes x cells in column-major order (R's layout) is cells x genes row-major: pipelined host path
        Yt = Y.T
        m, n = Yt.shape
        bkeys, Bt = replicate_keys(np.asarray(batch), ctx=actx)
        ctlm = tological(ctl, n)
        exact = _is_exact(BSPARAM)
        s = _svd_k(m_total, G, int(ctlm.sum()), exact, svd_k)
        s_dev, exact_dev = _device_rows_cap(s, ks[0], exact, n)
        newY, fit = _host_pipeline(ctx, Yt, ctlm, ks[0], keys=keys, G=G, s=s_dev, exact=exact_dev, batch_keys=bkeys, Bt=Bt)
        res = {ks[0]: {"newY": newY, "M": M, "fullalpha": fit.fullalpha}, "optimal_ruvK": ks[0],
               "stand_mean": fit.stand_mean, "stand_sd": fit.stand_sd}
        if return_all_RUV:
            return res
        one = dict(res[ks[0]])
        one["optimal_ruvK"] = ks[0]
        return one
    if isinstance(Y, DeviceMatrix):
        dY, copied = Y, False
    else:
        Y = np.asarray(Y, dtype=np.float64)
        # genes x cells: C-order needs a transpose on the device, Fortran order is already cells x genes row-major
        dY, copied = DeviceMatrix.from_host(ctx, Y.T), True
    m, n = dY.m, dY.n
    bkeys, Bt = replicate_keys(np.asarray(batch), ctx=ctx)
    ctl = tological(ctl, n)
    exact = _is_exact(BSPARAM)
    s = _svd_k(m_total, G, int(ctl.sum()), exact, svd_k)
    results = {}
    s_dev, exact_dev = _device_rows_cap(s, max(ks), exact, n)
    select = len(ks) > 1
    if fullalpha is None:
        # one fit serves every k (R/scRUVIII.R:64-75): the eigensolver must converge the top-k subspace for each of them
        extra = np.ascontiguousarray(sorted(set(ks[1:])), dtype=np.int32)
        L.check(L.lib().scm_ctx_set_conv_ranks(ctx.handle, extra.ctypes.data_as(C.c_void_p), int(extra.shape[0])))
        try:
            fit = _fit(ctx, dY, keys, G, s_dev, ks[0], exact_dev, batch_keys=bkeys, Bt=Bt, keep_gram=select)
        finally:
            L.check(L.lib().scm_ctx_set_conv_ranks(ctx.handle, None, 0))
    else:  # still need stand_mean / stand_sd: run the standardisation passes only via a fit with s = 1
        fit0 = _fit(ctx, dY, keys, G, 1, 1, True, batch_keys=bkeys, Bt=Bt, maxit=2, tol=1.0, keep_gram=select)
        fit = RUVFit(np.asanyarray(fullalpha, dtype=np.float64), None, fit0.stand_mean, fit0.stand_sd)
        fit.d_gram, fit.d_colmean = fit0.d_gram, fit0.d_colmean
    f_score, sil = np.ones(1), None
    if select:
        # R/scRUVIII.R:83-105: silhouette of the cell types (the replicate structure when cell_type is NULL, :89-92) and of
        # the batches in the first 10 PCs of every candidate's adjusted matrix, F-score of the two, which.max
        ct_keys, n_ct = (keys, G) if cell_type is None else replicate_keys(np.asarray(cell_type), ctx=ctx)
        sil = np.empty((2, len(ks)))
        for i, kk in enumerate(ks):
            with _Coef(ctx, n, fit.fullalpha, kk, ctl, fit.stand_mean, 1.0 / fit.stand_sd, fit.stand_sd) as coef:
                sil[:, i] = _silhouette_pair(ctx, fit, dY, coef, ct_keys, n_ct, bkeys, Bt)
        zo = (sil + 1.0) / 2.0
        ctv, btv = zo[0], 1.0 - zo[1]
        f_score = 2.0 * ctv * btv / (ctv + btv)
        fit.d_gram.free()
        fit.d_colmean.free()
    optimal = ks[int(np.argmax(f_score))]
    for kk in (ks if return_all_RUV else [optimal]):
        dout, _ = _adjust(ctx, dY, fit.fullalpha, kk, ctl, center=fit.stand_mean, pre=1.0 / fit.stand_sd, post=fit.stand_sd)
        newY = _finish(dout, copied)
        results[kk] = {"newY": newY, "M": M, "fullalpha": fit.fullalpha}
    if copied:
        dY.free()
    extras = {"optimal_ruvK": optimal, "f_score": f_score, "sil": sil}
    if return_all_RUV:
        results.update(extras)
        results["stand_mean"], results["stand_sd"] = fit.stand_mean, fit.stand_sd
        return results
    res = dict(results[optimal])
    res.update(extras)
    return res


def scRUVg(Y, ctl, k, Z=1, eta=None, include_intercept=True, fullW=None, svdyc=None, BSPARAM=None, ctx: Context | None = None):
    """R/scRUVg.R:30-85.  Y: m x n (cells x genes, numpy or DeviceMatrix).  Returns {'newY', 'W' (m x k), 'alpha' (k x n)}.
    Z = 1 (an intercept: what every documented call uses) is the covariate structure implemented on the device.
    `eta` (gene-wise covariates): the decomposition and alpha use Y0 = RUV1(Y, eta, ctl) (:59) while W alpha is subtracted
    from Y itself (:80).  `fullW` (m x >= k, a previous decomposition): W = fullW[, 1:k], alpha = solve(W'W) W' Y0 (:64,77-78;
    scm_ls_coef + scm_rankk_subtract).  `svdyc` without `fullW` leaves fullW NULL in the reference, which then fails at :77 --
    it is rejected here with that explanation.  Columns of W (and the matching rows of alpha) are defined up to sign, exactly
    like the reference's `svd`."""
    if not (np.isscalar(Z) and Z == 1):
        raise NotImplementedError("scRUVg on the device supports Z = 1 (intercept only)")
    if svdyc is not None and fullW is None:
        raise ValueError("svdyc without fullW: the reference never derives fullW from a given svdyc (R/scRUVg.R:64-75) and fails at "
                         "fullW[, seq_len(k)]; pass fullW = svdyc$u instead")
    ctx = ctx or (Y.ctx if isinstance(Y, DeviceMatrix) else default_context())
    m, n = (Y.m, Y.n) if isinstance(Y, DeviceMatrix) else np.shape(Y)
    ctlm = tological(ctl, n)
    k = int(k)
    if k > int(ctlm.sum()):
        raise ValueError("k must not be larger than the number of controls")  # R/scRUVg.R:56-58
    if k > min(L.MAX_K, m - 1):
        raise L.ScmError(L.ERR_UNSUPPORTED, f"k = {k} outside 1..min({L.MAX_K}, m - 1)")
    if eta is not None:
        eta_design_rows(eta, n, include_intercept)  # argument errors before any device work
    if fullW is not None:
        fullW = np.asarray(fullW, dtype=np.float64)
        if fullW.ndim != 2 or fullW.shape[0] != m or fullW.shape[1] < k:
            raise ValueError("fullW must be m x (at least k)")
    lib = L.lib()
    dY, copied = _as_device(ctx, Y)
    dY0 = _ruv1(ctx, dY, eta, ctlm, include_intercept, in_place=False) if eta is not None else dY  # Y itself is needed for :80
    d_ctl = ctx.to_device(np.ascontiguousarray(ctlm, dtype=np.uint8))
    d_alpha, d_center, d_sigma = ctx.empty((k, n)), ctx.empty((n,)), ctx.empty((k,))
    d_W = ctx.empty((m, k))
    dout = DeviceMatrix(ctx, m, n)
    coef = C.c_void_p()
    iters, resid = C.c_int32(0), C.c_double(0.0)
    try:
        if fullW is None:
            rc = lib.scm_ruvg_fit(ctx.handle, dY0.ptr, m, n, dY0.ld, d_ctl.ptr, k, L.MODE_EXACT if _is_exact(BSPARAM) else L.MODE_RANDOMIZED,
                                  1e-10, 500, 1, d_alpha.ptr, d_center.ptr, d_sigma.ptr, C.byref(coef), C.byref(iters), C.byref(resid))
            if rc == L.ERR_NOCONV:
                warnings.warn(lib.scm_last_error().decode(), L.NotConvergedWarning, stacklevel=2)
            else:
                L.check(rc)
            if eta is None:
                L.check(lib.scm_adjust(ctx.handle, dY.ptr, m, n, dY.ld, coef, None, 0, dout.ptr, dout.ld, d_W.ptr))
            else:  # W = (Y0 - mean) B from the projection pass alone, then Y - W alpha on the ORIGINAL matrix
                L.check(lib.scm_adjust(ctx.handle, dY0.ptr, m, n, dY0.ld, coef, None, 0, None, 0, d_W.ptr))
                L.check(lib.scm_rankk_subtract(ctx.handle, dY.ptr, m, n, dY.ld, d_W.ptr, k, d_alpha.ptr, dout.ptr, dout.ld))
        else:
            W = np.ascontiguousarray(fullW[:, :k])
            L.check(lib.scm_h2d(ctx.handle, d_W.ptr, W.ctypes.data_as(C.c_void_p), W.nbytes))
            d_mean = ctx.to_device(column_means(dY0))
            try:
                L.check(lib.scm_ls_coef(ctx.handle, dY0.ptr, m, n, dY0.ld, d_W.ptr, k, d_mean.ptr, d_alpha.ptr))
            finally:
                d_mean.free()
            L.check(lib.scm_rankk_subtract(ctx.handle, dY.ptr, m, n, dY.ld, d_W.ptr, k, d_alpha.ptr, dout.ptr, dout.ld))
        ctx.sync()
        W, alpha = d_W.to_host(), d_alpha.to_host()
    except Exception:
        dout.free()
        raise
    finally:
        if coef:
            lib.scm_coef_destroy(ctx.handle, coef)
        for a in (d_ctl, d_alpha, d_center, d_sigma, d_W):
            a.free()
        if dY0 is not dY:
            dY0.free()
    newY = _finish(dout, copied)
    if copied:
        dY.free()
    return {"newY": newY, "W": W, "alpha": alpha}


def pseudoRUVIII(Y, Y_pseudo, M, ctl, k=None, eta=None, include_intercept=True, inputcheck=True, fullalpha=None,
                 return_info=False, subset=None, BPPARAM=None, BSPARAM=None, normalised=True, verbose=True, byChunk=True,
                 chunkSize=50000, ctx: Context | None = None, gene_names=None):
    """R/scMerge2_utilis.R:17-150.  Fit on the pseudobulk matrix (P x n), adjust every cell of Y (m x n).
    byChunk / chunkSize / BPPARAM are accepted for parity; the fused kernel streams all cells in one launch."""
    del BPPARAM, verbose, byChunk, chunkSize, inputcheck
    ctx = ctx or (Y.ctx if isinstance(Y, DeviceMatrix) else default_context())
    n = Y.n if isinstance(Y, DeviceMatrix) else np.shape(Y)[1]
    P = Y_pseudo.m if isinstance(Y_pseudo, DeviceMatrix) else np.shape(Y_pseudo)[0]
    keys, G = replicate_keys(M)
    ctl = tological(ctl, n)
    exact = _is_exact(BSPARAM)
    s = _svd_k(P, G, int(ctl.sum()), exact, k)  # Exact branch caps at k here (R/scMerge2_utilis.R:60-61)
    if G >= P or k == 0:  # R/scMerge2_utilis.R:72-76 returns the pseudobulk matrix itself (after RUV1, :56)
        if eta is not None:
            dP, pcopied = _as_device(ctx, Y_pseudo)
            return _finish(_ruv1(ctx, dP, eta, ctl, include_intercept, in_place=pcopied), pcopied)
        return Y_pseudo
    if fullalpha is None:
        dP, pcopied = _as_device(ctx, Y_pseudo)
        if eta is not None:  # R/scMerge2_utilis.R:56: Y_pseudo <- RUV1(Y_pseudo, eta, ctl) (the pseudobulks only, not the cells)
            dP, pcopied = _ruv1(ctx, dP, eta, ctl, include_intercept, in_place=pcopied), True
        s_dev, exact_dev = _device_rows_cap(s, k, exact, n, rows=P)
        try:
            fit = _fit(ctx, dP, keys, G, s_dev, int(k), exact_dev)
        except L.NonFiniteError as e:  # the reference only warns here (R/scMerge2_utilis.R:49-52) and then fails in svd
            warnings.warn(str(e))
            raise
        if pcopied:
            dP.free()
        fullalpha = fit.fullalpha
    fullalpha = np.asanyarray(fullalpha, dtype=np.float64)
    if not normalised:
        return {"M": M, "fullalpha": fullalpha}
    if _is_host_matrix(Y) and subset is None:
        # host matrix: column means accumulate during the chunked upload, adjust + download are overlapped
        newY, st = _host_pipeline(ctx, Y, ctl, int(k), fullalpha_in=fullalpha, center_mode=1)
        if not return_info:
            return newY
        return {"newY": newY, "M": M, "fullalpha": fullalpha, "adjusted_means": st.stand_mean}
    dY, copied = _as_device(ctx, Y)
    means = column_means(dY)  # adjusted_means <- colMeans(Y) over ALL cells (R/scMerge2_utilis.R:97)
    sub = _resolve_subset(subset, n, gene_names)
    dout, _ = _adjust(ctx, dY, fullalpha, int(k), ctl, center=means, subset=sub)
    newY = _finish(dout, copied)
    if copied:
        dY.free()
    if not return_info:
        return newY
    return {"newY": newY, "M": M, "fullalpha": fullalpha, "adjusted_means": means}


def _resolve_subset(subset, n: int, gene_names=None):
    """`Y[, subset]` of pseudoRUVIII (R/scMerge2_utilis.R:117-121,131-136) -> 0-based column indices IN THE GIVEN ORDER.
    Logical mask: which(); integers: as they are (order and repeats kept); names (scMerge2 passes `chosen.hvg`,
    R/scMerge2.R:283,305): match(subset, colnames(Y)), and like R's subscript an unknown name is an error."""
    if subset is None:
        return None
    a = np.asarray(subset)
    if a.dtype == bool:
        return np.nonzero(tological(a, n))[0]
    if a.dtype.kind in ("U", "S", "O"):
        if gene_names is None:
            raise ValueError("gene names given as `subset` but no gene_names (colnames of Y)")
        pos = {g: i for i, g in enumerate(np.asarray(gene_names).tolist())}
        missing = [g for g in a.tolist() if g not in pos]
        if missing:
            raise IndexError(f"subscript out of bounds: {missing[:3]} not in colnames(Y)")
        return np.array([pos[g] for g in a.tolist()], dtype=np.int64)
    a = a.astype(np.int64)
    if a.size and (a.min() < 0 or a.max() >= n):
        raise IndexError(f"subscript out of bounds: gene subset outside [0, {n})")
    return a


def genes_by_name(sel, n: int, gene_names=None) -> np.ndarray:
    """Gene selection -> boolean mask of length n.  Masks / 0-based indices as in tological(); a list of gene NAMES is
    matched against `gene_names` (the rownames of the genes x cells matrix) like the reference's
    `intersect(rownames(exprsMat), ctl)` / `rownames(exprsMat) %in% ctl` (R/scMerge2.R:304,382-390): names that are not in
    the data are dropped silently and the matrix order is kept."""
    a = np.asarray(sel)
    if a.dtype.kind in ("U", "S", "O") and a.size and isinstance(a.reshape(-1)[0], (str, bytes)):
        if gene_names is None:
            raise ValueError("gene names given but the matrix has no gene_names")
        return np.isin(np.asarray(gene_names), a)
    return tological(a, n)


def getAdjustedMat(exprsMat, fullalpha, ctl=None, adjusted_means=None, ruvK=20, return_subset_genes=None,
                   ctx: Context | None = None, gene_names=None):
    """R/scMerge2.R:368-410.  exprsMat: genes x cells; returns genes x cells (numpy) or a DeviceMatrix (cells x genes
    device layout, which *is* genes x cells column-major) when given one.  `ctl` / `return_subset_genes` may be masks,
    0-based indices or gene names (matched against `gene_names` = rownames(exprsMat), as the reference does)."""
    ctx = ctx or (exprsMat.ctx if isinstance(exprsMat, DeviceMatrix) else default_context())
    n = exprsMat.n if isinstance(exprsMat, DeviceMatrix) else np.shape(exprsMat)[0]
    ctl_mask = np.ones(n, dtype=bool) if ctl is None else genes_by_name(ctl, n, gene_names)
    if ctl_mask.sum() == 0:
        raise ValueError("No provided ctl genes in the data")
    sub = None
    if return_subset_genes is not None:
        sub = np.nonzero(genes_by_name(return_subset_genes, n, gene_names))[0]
        if sub.shape[0] == 0:
            raise ValueError("No provided return_subset_genes genes in the data")
    if sub is None and isinstance(exprsMat, np.ndarray) and _is_host_matrix(exprsMat.T):
        # genes x cells in column-major order (R's layout) is cells x genes by rows: the host-streaming driver.  With the stored
        # `adjusted_means` nothing has to be known about the cells beforehand, so every chunk goes up, is adjusted and comes
        # home in ONE pass with uploads and downloads running concurrently; without them the column means accumulate during the
        # upload first (the in-core form)
        fa = np.asanyarray(fullalpha, dtype=np.float64)
        if adjusted_means is None:
            newY, _ = _host_pipeline(ctx, exprsMat.T, ctl_mask, int(ruvK), fullalpha_in=fa, center_mode=1)
        else:
            newY, _ = _host_pipeline(ctx, exprsMat.T, ctl_mask, int(ruvK), fullalpha_in=fa, center_mode=2,
                                     center=np.asarray(adjusted_means, dtype=np.float64))
        return newY.T
    if isinstance(exprsMat, DeviceMatrix):
        dY, copied = exprsMat, False
    else:
        dY, copied = DeviceMatrix.from_host(ctx, np.asarray(exprsMat, dtype=np.float64).T), True
    means = column_means(dY) if adjusted_means is None else np.asarray(adjusted_means, dtype=np.float64)
    dout, _ = _adjust(ctx, dY, np.asanyarray(fullalpha, dtype=np.float64), int(ruvK), ctl_mask, center=means, subset=sub)
    if copied:
        dY.free()
        res = dout.to_host().T
        dout.free()
        return res
    return dout


def multi_fastRUVIII(mg: MultiContext, Y, M, ctl, k, **kw):
    """fastRUVIII of a host matrix on ALL the devices of a MultiContext (scm_multi_ruv3_host): the call a single R process makes."""
    return fastRUVIII(Y, M, ctl, k=k, ctx=mg, **kw)


# ---- K1 front-ends ----------------------------------------------------------------------------------------
def segmented_sums(dY: DeviceMatrix, keys: np.ndarray, G: int, center: np.ndarray | None = None, check_finite=False,
                   mean=False):
    """(sums | means G x n, counts G[, nonfinite flag]) of the rows of a device matrix grouped by int32 key (-1 = skip)."""
    ctx = dY.ctx
    d_keys = ctx.to_device(np.ascontiguousarray(keys, dtype=np.int32))
    d_sums, d_cnt = ctx.empty((G, dY.n)), ctx.empty((G,), np.int64)
    d_center = ctx.to_device(np.ascontiguousarray(center, dtype=np.float64)) if center is not None else None
    d_flag = ctx.to_device(np.zeros(1, dtype=np.int32)) if check_finite else None
    if mean:
        L.check(L.lib().scm_seg_colmean(ctx.handle, dY.ptr, dY.m, dY.n, dY.ld, d_keys.ptr, G, d_sums.ptr, d_cnt.ptr,
                                        _dptr(d_flag)))
    else:
        L.check(L.lib().scm_seg_colsum(ctx.handle, dY.ptr, dY.m, dY.n, dY.ld, d_keys.ptr, G, _dptr(d_center), d_sums.ptr,
                                       d_cnt.ptr, _dptr(d_flag)))
    sums, cnt = d_sums.to_host(), d_cnt.to_host()
    flag = int(d_flag.to_host()[0]) if d_flag is not None else None
    for a in (d_keys, d_sums, d_cnt, d_center, d_flag):
        if a is not None:
            a.free()
    return (sums, cnt, flag) if check_finite else (sums, cnt)


def column_means(dY: DeviceMatrix) -> np.ndarray:
    means, _cnt = segmented_sums(dY, np.zeros(dY.m, dtype=np.int32), 1, mean=True)
    return means[0]


def aggregate_matrix(x, groupings, ctx: Context | None = None):
    """aggregate.Matrix (R/scMerge2_utilis.R:325-346): rows of x (cells x genes) summed by group label;
    result rows in sorted level order.  Returns (sums, levels, counts)."""
    ctx = ctx or default_context()
    levels, inv = np.unique(np.asarray(groupings), return_inverse=True)
    dX, copied = _as_device(ctx, x)
    sums, cnt = segmented_sums(dX, inv.astype(np.int32), int(levels.shape[0]))
    if copied:
        dX.free()
    return sums, levels, cnt


def create_pseudoBulk(exprsMat, cell_info, k_fold=30, which=None, ctx: Context | None = None, seed=None):
    """create_pseudoBulk (R/scMerge2_utilis.R:461-486) for ONE batch.  exprsMat: genes x cells.  `which` is the fold
    id (1..K) per cell -- cvTools::cvFolds' random assignment; pass the R draw for bit-identical groups, otherwise a
    seeded NumPy permutation with the same structure (`rep(1:K, length.out = n)` over a random order) is used.
    One segmented-mean launch over the composite key (fold, type) replaces the per-fold sparse products.
    Returns (bulk P x genes, names 'type_fold') with rows fold-major and types in sorted level order."""
    ctx = ctx or default_context()
    X = np.asarray(exprsMat, dtype=np.float64)
    ncell = X.shape[1]
    K = min(ncell, int(k_fold))
    if which is None:
        order = np.random.default_rng(seed).permutation(ncell)
        which = np.empty(ncell, dtype=np.int64)
        which[order] = (np.arange(ncell) % K) + 1
    which = np.asarray(which)
    levels, t = np.unique(np.asarray(cell_info), return_inverse=True)
    T = levels.shape[0]
    key = ((which - 1) * T + t).astype(np.int32)
    dX = DeviceMatrix.from_host(ctx, X.T)
    means, cnt = segmented_sums(dX, key, K * T, mean=True)
    dX.free()
    keep = cnt > 0  # droplevels(): (fold, type) combinations without cells give no row
    bulk = means[keep]
    names = [f"{levels[i % T]}_{i // T + 1}" for i in np.nonzero(keep)[0]]
    return bulk, names
