"""CPU oracle: float64 NumPy restatement of scMerge's RUV-III fit-and-adjust path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`
/ `--impl reference` legs may import this module; the product path (`scmerge_b200/`) never does.

PARITY UNPINNED: the reference is a pure-R package (no golden vectors in its tests, R is not
installable here), so the absolute values produced by this restatement cannot be checked against a
run of the reference itself.  What *is* pinned (tests/test_oracle.py) are the relations the
reference's own tests/docs assert: exact ~= randomized to 2 % (`tests/testthat/test-fastRUVIII.R:11-16`),
equality with classic `ruv::RUVIII` (`R/fastRUVIII.R:30-38`), cell-permutation equivariance at
1.5e-8 (`tests/testthat/test-resampling-columns.R:26`), and agreement of the literal form with the
algebraically reduced (Gram) form.

Every function cites the reference lines it restates (paths relative to /root/reference).
Third-party arithmetic that the reference calls but does not vendor (DESCRIPTION:22-41, unpinned):
  ruv::replicate.matrix, ruv::RUV1, ruv::RUVIII  (CRAN `ruv`)      -> restated from the published algorithm
  BiocSingular::runSVD(ExactParam) = base::svd (LAPACK dgesdd)     -> numpy.linalg.svd
  BiocSingular::runSVD(RandomParam) = rsvd::rsvd(p=10, q=2)        -> `rsvd` below
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------------
# helpers restating un-vendored `ruv` functions
# --------------------------------------------------------------------------------------------------


def replicate_matrix(a) -> np.ndarray:
    """ruv::replicate.matrix (call sites R/fastRUVIII.R:48, R/scMerge2_utilis.R:41, R/ruvSimulate.R:73).

    numeric 0/1 matrix -> returned unchanged; label vector -> one-hot `model.matrix(~a-1)` with
    columns in sorted level order; 2-D array of labels (data.frame) -> rows pasted with "_" first.
    """
    a = np.asarray(a)
    if a.ndim == 2 and np.issubdtype(a.dtype, np.number):
        return a.astype(np.float64)
    if a.ndim == 2:
        a = np.array(["_".join(str(x) for x in row) for row in a])
    levels, inv = np.unique(a, return_inverse=True)
    M = np.zeros((a.shape[0], levels.shape[0]))
    M[np.arange(a.shape[0]), inv] = 1.0
    return M


def tological(ctl, n: int) -> np.ndarray:
    """R/fastRUVIII.R:113-117. Accepts a logical mask or 0-based integer indices."""
    ctl = np.asarray(ctl)
    if ctl.dtype == bool:
        out = np.zeros(n, dtype=bool)
        out[: ctl.shape[0]] = ctl
        return out
    out = np.zeros(n, dtype=bool)
    out[ctl] = True
    return out


def my_residop(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """R/fastRUVIII.R:121-129, literal: A - B (B'B)^-1 B' A (dense projector, left to right)."""
    tBB = B.T @ B
    tBB_inv = np.linalg.inv(tBB)  # Matrix::solve(tBB)
    BtBB_inv = B @ tBB_inv
    tBA = B.T @ A
    return A - BtBB_inv @ tBA


def design_matrix(a, n: int, include_intercept: bool = True) -> np.ndarray:
    """ruv::design.matrix as ruv::RUV1 uses it for gene-wise covariates (CRAN `ruv`, un-vendored; restated from the published
    source): numeric input only.  1 -> a column of ones; a vector -> one column; a matrix with n columns and not n rows is
    transposed.  include_intercept: cbind(1, A) unless sum(residop(1, A)^2) <= 1e-8.  Returns n x q."""
    A = np.asarray(a, dtype=np.float64)
    if A.ndim == 0:
        A = np.ones((n, 1))
    elif A.ndim == 1:
        A = A.reshape(-1, 1)
    if A.shape[0] != n:
        A = A.T
    assert A.shape[0] == n
    if include_intercept:
        one = np.ones((n, 1))
        if (my_residop(one, A) ** 2).sum() > 1e-8:
            A = np.concatenate([one, A], axis=1)
    return A


def RUV1(Y: np.ndarray, eta, ctl, include_intercept: bool = True) -> np.ndarray:
    """ruv::RUV1 (call sites R/fastRUVIII.R:64, R/scRUVg.R:59, R/scMerge2_utilis.R:56), literal:
    Y - Y[, ctl] etac' solve(etac etac') eta with eta = t(design.matrix(eta)) (q x n); eta = NULL returns Y."""
    if eta is None:
        return Y
    Y = np.asarray(Y, dtype=np.float64)
    ctl = tological(ctl, Y.shape[1])
    E = design_matrix(eta, Y.shape[1], include_intercept).T
    Ec = E[:, ctl]
    return Y - Y[:, ctl] @ Ec.T @ np.linalg.inv(Ec @ Ec.T) @ E


def rsvd(A: np.ndarray, k: int, p: int = 10, q: int = 2, rng=None):
    """rsvd::rsvd defaults as dispatched by BiocSingular::RandomParam (un-vendored; published algorithm:
    Erichson et al., 'Randomized matrix decompositions using R').  Returns (u, d, v) truncated to k."""
    rng = np.random.default_rng(0) if rng is None else rng
    flipped = A.shape[0] < A.shape[1]
    if flipped:
        A = A.T
    m, n = A.shape
    l = min(int(round(k)) + int(round(p)), n)
    O = rng.standard_normal((n, l))
    Y = A @ O
    for _ in range(q):
        Y, _r = np.linalg.qr(Y)
        Z, _r = np.linalg.qr(A.T @ Y)
        Y = A @ Z
    Q, _r = np.linalg.qr(Y)
    B = Q.T @ A
    ub, d, vt = np.linalg.svd(B, full_matrices=False)
    u, v = Q @ ub, vt.T
    u, d, v = u[:, :k], d[:k], v[:, :k]
    return (v, d, u) if flipped else (u, d, v)


def svd_k_bound(m: int, n_groups: int, n_ctl: int, exact: bool, cap) -> int:
    """R/fastRUVIII.R:66-70 (cap = svd_k, default 50) and R/scMerge2_utilis.R:60-65 (cap = k).
    The non-exact branch ignores the user's svd_k (reference quirk)."""
    if exact:
        return int(min(m - n_groups, n_ctl, cap))
    return int(min(m - n_groups, n_ctl))


# --------------------------------------------------------------------------------------------------
# fastRUVIII
# --------------------------------------------------------------------------------------------------


def fastRUVIII(Y, M, ctl, k, svd_k=50, exact=True, fullalpha=None, return_info=False,
               inputcheck=True, rng=None, eta=None, include_intercept=True):
    """R/fastRUVIII.R:41-111, literal.  Y is m x n (cells x genes).  `exact` <=> inherits(BSPARAM,
    "ExactParam").  eta is NULL at every scMerge call site, where ruv::RUV1 is the identity (:64)."""
    Y = np.asarray(Y, dtype=np.float64)
    m, n = Y.shape
    M = replicate_matrix(M)
    ctl = tological(ctl, n)
    if inputcheck:  # :52-60
        if np.isnan(Y).sum() > 0:
            raise ValueError("Y contains missing values.  This is not supported.")
        if np.isinf(Y).sum() > 0:
            raise ValueError("Y contains infinities.  This is not supported.")
    Y = RUV1(Y, eta, ctl, include_intercept)  # :63-64
    s = svd_k_bound(m, M.shape[1], int(ctl.sum()), exact, svd_k)
    if M.shape[1] >= m or k == 0:  # :78-80
        newY, fullalpha = Y, None
    else:
        if fullalpha is None:
            Y0 = my_residop(Y, M)  # :89
            if exact:
                u, _d, _vt = np.linalg.svd(Y0, full_matrices=False)  # :91 base::svd via ExactParam
                u = u[:, :s]
            else:
                u, _d, _v = rsvd(Y0, s, rng=rng)
            fullalpha = u.T @ Y  # :93  (Y, not Y0)
        alpha = fullalpha[: min(k, fullalpha.shape[0])]  # :97
        ac = alpha[:, ctl]  # :98
        W = Y[:, ctl] @ ac.T @ np.linalg.inv(ac @ ac.T)  # :99 explicit solve(), left to right
        newY = Y - W @ alpha  # :100-101
    if not return_info:
        return newY
    return {"newY": newY, "M": M, "fullalpha": fullalpha}


def group_ids(M: np.ndarray) -> np.ndarray:
    """int32 group id per row of a one-hot replicate matrix (what the device path consumes)."""
    M = np.asarray(M)
    if not np.all((M == 0) | (M == 1)) or not np.all(M.sum(axis=1) == 1):
        raise ValueError("M is not a one-hot partition matrix")
    return np.argmax(M, axis=1).astype(np.int32)


def fastRUVIII_reduced(Y, M, ctl, k, svd_k=50):
    """Algebraically reduced exact form (SURVEY.md section 0.3 / Appendix A 'device-side reduced form'):
    group-mean residual -> n x n Gram -> top-s eigenpairs -> fullalpha = sqrt(lambda) v' ->
    newY = Y - (Y_c V_c (V_c'V_c)^-1) V_k'.  Must agree with `fastRUVIII` to ~1e-12."""
    Y = np.asarray(Y, dtype=np.float64)
    m, n = Y.shape
    g = group_ids(replicate_matrix(M))
    ctl = tological(ctl, n)
    G = int(g.max()) + 1
    s = svd_k_bound(m, G, int(ctl.sum()), True, svd_k)
    counts = np.bincount(g, minlength=G).astype(np.float64)
    sums = np.zeros((G, n))
    np.add.at(sums, g, Y)
    Y0 = Y - (sums / counts[:, None])[g]
    lam, V = np.linalg.eigh(Y0.T @ Y0)
    order = np.argsort(lam)[::-1][:s]
    lam, V = lam[order], V[:, order]
    fullalpha = np.sqrt(np.maximum(lam, 0))[:, None] * V.T
    kk = min(k, s)
    Vk = V[:, :kk]
    Vc = Vk[ctl]
    Bm = Vc @ np.linalg.inv(Vc.T @ Vc)
    newY = Y - (Y[:, ctl] @ Bm) @ Vk.T
    return {"newY": newY, "fullalpha": fullalpha}


def ruvIII_classic(Y, M, ctl, k):
    """ruv::RUVIII (CRAN `ruv`, un-vendored), restated from the published algorithm (Molania et al. 2019):
    Y0 = residop(Y, M); fullalpha = t(eigen(Y0 Y0')$vectors[, 1:(m - ncol(M))]) %*% Y; then alpha/ac/W as
    in fastRUVIII.  The reference documents fastRUVIII == ruv::RUVIII (R/fastRUVIII.R:30-38)."""
    Y = np.asarray(Y, dtype=np.float64)
    m, n = Y.shape
    M = replicate_matrix(M)
    ctl = tological(ctl, n)
    Y0 = my_residop(Y, M)
    lam, U = np.linalg.eigh(Y0 @ Y0.T)
    U = U[:, np.argsort(lam)[::-1][: m - M.shape[1]]]
    fullalpha = U.T @ Y
    alpha = fullalpha[: min(k, fullalpha.shape[0])]
    ac = alpha[:, ctl]
    W = Y[:, ctl] @ ac.T @ np.linalg.inv(ac @ ac.T)
    return Y - W @ alpha


# --------------------------------------------------------------------------------------------------
# scRUVIII (standardise -> fastRUVIII -> de-standardise)
# --------------------------------------------------------------------------------------------------


def standardize2(Y, batch):
    """R/scRUVIII.R:158-175.  Y is genes x cells here (as in the reference)."""
    Y = np.asarray(Y, dtype=np.float64)
    num_cell = Y.shape[1]
    levels, b = np.unique(np.asarray(batch), return_inverse=True)
    num_batch = levels.shape[0]
    stand_mean = Y.mean(axis=1)
    design = np.zeros((num_cell, num_batch))
    design[np.arange(num_cell), b] = 1.0
    a = design.T @ design
    bmat = (Y @ design).T
    B_hat = np.linalg.inv(a.T @ a) @ a.T @ bmat  # solve_axb, :153-156
    B_designed = B_hat.T @ design.T
    stand_var = ((Y - B_designed) ** 2).sum(axis=1) / (num_cell - num_batch)
    stand_Y = (Y - stand_mean[:, None]) / np.sqrt(stand_var)[:, None]
    return {"stand_Y": stand_Y, "stand_mean": stand_mean, "stand_var": stand_var}


def silhouette_avg_width(X, labels):
    """summary(cluster::silhouette(labels, dist(X)))$avg.width as used by kBET_batch_sil (R/scRUVIII.R:192-199; `cluster` is
    un-vendored, this is its published definition): s_i = (b_i - a_i) / max(a_i, b_i), a_i = mean distance to the OTHER
    members of i's cluster, b_i = smallest mean distance to another cluster, s_i = 0 for singleton clusters."""
    X = np.asarray(X, dtype=np.float64)
    lev, lab = np.unique(np.asarray(labels), return_inverse=True)
    C, m = lev.shape[0], X.shape[0]
    if C < 2:
        return float("nan")
    D = np.empty((m, m))
    for r0 in range(0, m, 512):  # dist(): plain Euclidean distances from coordinate differences
        diff = X[r0:r0 + 512, None, :] - X[None, :, :]
        D[r0:r0 + 512] = np.sqrt((diff * diff).sum(axis=2))
    onehot = np.zeros((m, C))
    onehot[np.arange(m), lab] = 1.0
    S = D @ onehot  # sum of distances to every cluster
    cnt = onehot.sum(axis=0)
    own = cnt[lab]
    a = S[np.arange(m), lab] / np.maximum(own - 1, 1)
    other = S / cnt[None, :]
    other[np.arange(m), lab] = np.inf
    b = other.min(axis=1)
    s = np.where(own > 1, (b - a) / np.maximum(a, b), 0.0)
    return float(s.mean())


def calculate_sil(newY, cell_type, batch, rank=10):
    """calculateSil (R/scRUVIII.R:182-190): BiocSingular::runPCA(newY, rank = 10, center = TRUE, scale = TRUE) -- scores
    x = U D of the column-centred, unit-variance (n - 1 denominator) matrix -- then the average silhouette width of the
    cell types and of the batches in the first 10 PCs (Euclidean `dist`)."""
    Y = np.asarray(newY, dtype=np.float64)
    Xc = Y - Y.mean(axis=0)
    Xs = Xc / Y.std(axis=0, ddof=1)
    U, S, _ = np.linalg.svd(Xs, full_matrices=False)
    r = min(rank, S.shape[0])
    x = U[:, :r] * S[:r]
    return silhouette_avg_width(x, cell_type), silhouette_avg_width(x, batch)


def scRUVIII(Y, M, ctl, k, batch, fullalpha=None, exact=True, svd_k=50, rng=None, cell_type=None):
    """R/scRUVIII.R:41-137.  Y is genes x cells.  `k` may be an int or a sequence; returns {k: {newY (cells x genes),
    fullalpha}} plus 'optimal_ruvK': k[0] for a single k (f_score = 1, :80-82), otherwise the k maximising the F-score of
    the cell-type and (1 - batch) silhouette coefficients (:83-105; computed on the standardised adjusted matrices, i.e.
    BEFORE the mean / sd are added back at :117-119), with 'f_score' and 'sil' (2 x len(k))."""
    ks = [int(k)] if np.isscalar(k) else [int(x) for x in k]
    sc = standardize2(Y, batch)
    stand_tY = sc["stand_Y"].T
    stand_sd = np.sqrt(sc["stand_var"])
    stand_mean = sc["stand_mean"]
    first = fastRUVIII(stand_tY, M, ctl, ks[0], svd_k=svd_k, exact=exact, fullalpha=fullalpha,
                       return_info=True, rng=rng)
    out = {ks[0]: first}
    for kk in ks[1:]:  # :70-74 re-use fullalpha
        out[kk] = fastRUVIII(stand_tY, M, ctl, kk, fullalpha=first["fullalpha"], return_info=True)
    if len(ks) == 1:
        f_score = np.ones(1)  # :80-82
    else:
        if cell_type is None:  # :89-92 replicate structure stands in for the cell types
            cell_type = group_ids(replicate_matrix(M))
        sil = np.array([calculate_sil(out[kk]["newY"], cell_type, batch) for kk in ks]).T  # :94-96, 2 x len(k)
        zo = (sil + 1.0) / 2.0  # zeroOneScale :143-146
        ct, bt = zo[0], 1.0 - zo[1]
        f_score = 2.0 * ct * bt / (ct + bt)  # f_measure :177-180
        out["sil"] = sil
    for kk in ks:  # :117-119
        r = dict(out[kk])
        r["newY"] = (r["newY"].T * stand_sd[:, None] + stand_mean[:, None]).T
        out[kk] = r
    out["f_score"] = f_score
    out["optimal_ruvK"] = ks[int(np.argmax(f_score))]
    out["stand_mean"], out["stand_sd"] = stand_mean, stand_sd
    return out


# --------------------------------------------------------------------------------------------------
# scMerge2: pseudobulk aggregation, pseudoRUVIII, getAdjustedMat
# --------------------------------------------------------------------------------------------------


def create_chunk(n: int, chunk_size: int = 1000):
    """R/scMerge2_utilis.R:153-170.  Returns 1-based inclusive (start, end) pairs like the reference."""
    starts = list(range(1, n + 1, chunk_size))
    ends = [s + chunk_size - 1 for s in starts]
    ends[-1] = n
    if len(ends) > 1 and ends[-1] - ends[-2] < chunk_size / 2:
        starts, ends = starts[:-1], ends[:-1]
        ends[-1] = n
    return list(zip(starts, ends))


def create_pseudobulk(exprs, cell_info, which, k_fold=None):
    """Arithmetic of create_pseudoBulk + aggregate.Matrix, R/scMerge2_utilis.R:461-486,325-346.

    exprs: genes x cells of ONE batch; cell_info: label per cell; `which`: fold id (1..K) per cell, i.e.
    cvTools::cvFolds' `which` re-ordered to cell order (the R RNG draw is an input, not restated).
    Returns (bulk: P x genes, names: 'type_fold'), rows fold-major, types in sorted level order."""
    exprs = np.asarray(exprs, dtype=np.float64)
    cell_info = np.asarray(cell_info)
    which = np.asarray(which)
    K = int(which.max()) if k_fold is None else int(k_fold)
    rows, names = [], []
    for i in range(1, K + 1):
        idx = np.nonzero(which == i)[0]
        types = np.unique(cell_info[idx])
        for t in types:
            sel = idx[cell_info[idx] == t]
            rows.append(exprs[:, sel].sum(axis=1) / sel.shape[0])
            names.append(f"{t}_{i}")
    return np.array(rows), names


def pseudoRUVIII(Y, Y_pseudo, M, ctl, k, exact=True, fullalpha=None, subset=None, normalised=True,
                 by_chunk=True, chunk_size=50000, rng=None, eta=None, include_intercept=True):
    """R/scMerge2_utilis.R:17-150.  Y: cells x genes, Y_pseudo: pseudobulks x genes, M over pseudobulks."""
    Y = np.asarray(Y, dtype=np.float64)
    Y_pseudo = np.asarray(Y_pseudo, dtype=np.float64)
    n = Y.shape[1]
    m_pseudo = Y_pseudo.shape[0]
    M = replicate_matrix(M)
    ctl = tological(ctl, n)
    Y_pseudo = RUV1(Y_pseudo, eta, ctl, include_intercept)  # :56 (the pseudobulks only)
    s = svd_k_bound(m_pseudo, M.shape[1], int(ctl.sum()), exact, k)  # :60-65 (cap = k)
    if M.shape[1] >= m_pseudo or k == 0:  # :72-76 returns the pseudobulk matrix itself
        return Y_pseudo
    if fullalpha is None:  # :83-85
        Y0 = my_residop(Y_pseudo, M)
        if exact:
            u, _d, _vt = np.linalg.svd(Y0, full_matrices=False)
            u = u[:, :s]
        else:
            u, _d, _v = rsvd(Y0, s, rng=rng)
        fullalpha = u.T @ Y_pseudo
    alpha = fullalpha[: min(k, fullalpha.shape[0])]
    if not normalised:
        return {"M": M, "fullalpha": fullalpha}
    ac = alpha[:, ctl]
    adjusted_means = Y.mean(axis=0)  # :97 over ALL cells
    inv = np.linalg.inv(ac @ ac.T)
    cols = slice(None) if subset is None else np.asarray(subset)

    def adjust(rows):
        Y_stand = Y[rows] - adjusted_means
        W = Y_stand[:, ctl] @ ac.T @ inv
        return Y[rows][:, cols] - W @ alpha[:, cols]  # subtraction from the UNcentred Y (:117-121)

    if Y.shape[0] >= 1.5 * chunk_size and by_chunk:  # :104-124
        parts = [adjust(slice(a - 1, b)) for a, b in create_chunk(Y.shape[0], chunk_size)]
        newY = np.concatenate(parts, axis=0)
    else:
        newY = adjust(slice(None))
    return {"newY": newY, "M": M, "fullalpha": fullalpha, "adjusted_means": adjusted_means}


def getAdjustedMat(exprsMat, fullalpha, ctl=None, adjusted_means=None, ruvK=20, return_subset_genes=None):
    """R/scMerge2.R:368-410.  exprsMat: genes x cells; ctl / return_subset_genes: boolean masks or 0-based
    indices over genes (the reference resolves names with intersect(colnames, .), keeping matrix order)."""
    Y = np.asarray(exprsMat, dtype=np.float64).T
    n = Y.shape[1]
    ctl = np.ones(n, dtype=bool) if ctl is None else tological(ctl, n)
    means = Y.mean(axis=0) if adjusted_means is None else np.asarray(adjusted_means)
    Y_stand = Y - means
    alpha = fullalpha[: min(ruvK, fullalpha.shape[0])]
    ac = alpha[:, ctl]
    W = Y_stand[:, ctl] @ ac.T @ np.linalg.inv(ac @ ac.T)
    if return_subset_genes is not None:
        sub = np.nonzero(tological(return_subset_genes, n))[0]
        newY = Y[:, sub] - W @ alpha[:, sub]
    else:
        newY = Y - W @ alpha
    return newY.T


# --------------------------------------------------------------------------------------------------
# synthetic data + comparison metrics
# --------------------------------------------------------------------------------------------------


def scRUVg(Y, ctl, k, eta=None, include_intercept=True, fullW=None):
    """R/scRUVg.R:30-85 with the default Z = 1 (literal: svd of Y0[, ctl] or of its m x m cross product when the aspect ratio
    is >= 5, :66-73).  `fullW` (m x >= k): the decomposition is skipped (:64) and W is its first k columns.  (`svdyc` alone
    leaves fullW NULL in the reference and fails at :77; it is not restated.)  Returns dict(newY, W, alpha)."""
    Y = np.asarray(Y, dtype=np.float64)
    m, n = Y.shape
    ctl = tological(ctl, n)
    if k > ctl.sum():  # :56-58
        raise ValueError("k must not be larger than the number of controls")
    Z = np.ones((m, 1))  # :37-41, design.matrix of an intercept
    q = 1
    Y0 = RUV1(Y, eta, ctl, include_intercept)  # :59
    Y0 = my_residop(Y0, Z)  # :61-63
    if fullW is None:  # :64-75
        Y0ctl = Y0[:, ctl]
        mat = Y0ctl
        if max(mat.shape) / min(mat.shape) >= 5:  # :68-71
            mat = Y0ctl @ Y0ctl.T
        u = np.linalg.svd(mat, full_matrices=False)[0]  # :72
        fullW = u[:, : min(m - q, int(ctl.sum()))]  # :73
    fullW = np.asarray(fullW, dtype=np.float64)
    W = fullW[:, :k]  # :77
    alpha = np.linalg.solve(W.T @ W, W.T) @ Y0  # :78
    return {"newY": Y - W @ alpha, "W": W, "alpha": alpha}  # :80


def ruv_simulate(m=100, n=5000, nc=None, n_celltypes=3, n_batch=2, k=20, lam=0.1, rng=None):
    """Distribution of R/ruvSimulate.R:41-84 with NumPy's RNG (R's RNG stream is not reproducible here)."""
    rng = np.random.default_rng(0) if rng is None else rng
    nc = n // 2 if nc is None else nc
    ctl = np.zeros(n, dtype=bool)
    ctl[:nc] = True
    celltypes = (np.arange(m) % n_celltypes) + 1
    X = celltypes[:, None].astype(np.float64)
    beta = rng.poisson(2 * lam, size=(1, n)).astype(np.float64)
    beta[:, ctl] = 0
    W = rng.poisson(lam, size=(m, k)).astype(np.float64)
    W[:, 0] = np.arange(m) % n_batch
    alpha = rng.poisson(lam, size=(k, n)).astype(np.float64)
    eps = rng.poisson(lam, size=(m, n)).astype(np.float64)
    Y = X @ beta + W @ alpha + eps
    cell_types = np.array([f"cellTypes{c}" for c in celltypes])
    batch = np.array([f"batch{(i % n_batch) + 1}" for i in range(m)])
    M = replicate_matrix(np.stack([cell_types, batch], axis=1))
    return {"Y": Y, "ctl": ctl, "M": M, "cellTypes": cell_types, "batch": batch, "X": X}


def planted_factors(m, n, c, k, n_types=20, n_batch=4, seed=20240, noise=0.5, dtype=np.float64):
    """Generator G1 of SURVEY.md section 8(d): gene offsets + cell-type means (zero on ctl) + k unwanted
    factors (one-hot batch + Gaussian columns, scaled to leave a spectral gap after factor k) + noise."""
    rng = np.random.default_rng(seed)
    ctl = np.zeros(n, dtype=bool)
    ctl[:c] = True
    types = rng.integers(0, n_types, size=m)
    batch = rng.integers(0, n_batch, size=m)
    theta = rng.standard_normal((n_types, n))
    theta[:, ctl] = 0
    Wb = np.zeros((m, k))
    nb = min(n_batch - 1, k)  # one-hot columns for all batches but the last: full column rank after centring
    hit = batch < nb
    Wb[np.nonzero(hit)[0], batch[hit]] = 1.0
    if k > nb:
        Wb[:, nb:] = rng.standard_normal((m, k - nb))
    scale = np.linspace(3.0, 1.5, k)
    A = rng.standard_normal((k, n)) * scale[:, None]
    Y = rng.standard_normal(n) * 2.0 + theta[types] + Wb @ A + noise * rng.standard_normal((m, n))
    return {"Y": Y.astype(dtype), "ctl": ctl, "types": types.astype(np.int32), "batch": batch.astype(np.int32)}


def rel_frobenius(a, b) -> float:
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def principal_angles(A_rows, B_rows) -> np.ndarray:
    """Principal angles (radians) between the row spaces of two k x n matrices."""
    Qa, _ = np.linalg.qr(np.asarray(A_rows).T)
    Qb, _ = np.linalg.qr(np.asarray(B_rows).T)
    sv = np.clip(np.linalg.svd(Qa.T @ Qb, compute_uv=False), -1.0, 1.0)
    # use sin for small angles: || (I - Qb Qb') Qa ||
    resid = Qa - Qb @ (Qb.T @ Qa)
    sin = np.linalg.svd(resid, compute_uv=False)
    return np.where(sv > 0.9, np.arcsin(np.clip(sin[::-1], 0, 1)), np.arccos(sv))


def sign_align(rows, ref_rows) -> np.ndarray:
    """Flip the sign of each row to match the reference rows (singular vectors are sign-ambiguous)."""
    rows = np.array(rows, copy=True)
    sg = np.sign(np.einsum("ij,ij->i", rows, np.asarray(ref_rows)))
    sg[sg == 0] = 1
    return rows * sg[:, None]


# --------------------------------------------------------------------------------------------------
# scMerge2 core (cosineNorm -> construct_pseudoBulk -> pseudoRUVIII), R/scMerge2.R:193-309
# --------------------------------------------------------------------------------------------------


def cosine_norm(X, min_norm=1e-8):
    """batchelor::cosineNorm(x) for genes x cells (un-vendored; call site R/scMerge2.R:197): each cell (column)
    divided by its L2 norm, norms floored at 1e-8."""
    X = np.asarray(X, dtype=np.float64)
    l2 = np.maximum(np.sqrt((X ** 2).sum(axis=0)), min_norm)
    return X / l2


def scmerge2_core(exprs_genes_x_cells, batch, clust, which, ctl=None, ruvK=20, k_fold=30, cosine=True, exact=False, rng=None):
    """The arithmetic of scMerge2 between clustering and output for the supervised replicate structure
    (pseudobulks of the same cluster label are replicates): cosineNorm (R/scMerge2.R:193-198), create_pseudoBulk per
    batch (R/scMerge2_utilis.R:259-263,461-486), pseudoRUVIII (R/scMerge2.R:298-309).  `which`: fold id per cell."""
    X = np.asarray(exprs_genes_x_cells, dtype=np.float64)
    if cosine:
        X = cosine_norm(X)
    batch, clust, which = np.asarray(batch), np.asarray(clust), np.asarray(which)
    n = X.shape[0]
    ctl = np.ones(n, dtype=bool) if ctl is None else tological(ctl, n)
    bulks, names, rep = [], [], []
    for i, b in enumerate(np.unique(batch), start=1):  # batch_list <- sort(unique(batch)) (R/scMerge2.R:139)
        sel = batch == b
        bulk, nm = create_pseudobulk(X[:, sel], clust[sel], which[sel], k_fold=k_fold)
        bulks.append(bulk)
        names += [f"{i}_{x}" for x in nm]  # rownames(res) <- paste(i, rownames(res), sep = "_"): the batch INDEX (R/scMerge2_utilis.R:266)
        rep += [x.rsplit("_", 1)[0] for x in nm]
    Yp = np.concatenate(bulks)
    res = pseudoRUVIII(X.T, Yp, np.array(rep), ctl, ruvK, exact=exact, rng=rng)
    res["bulkExprs"], res["names"] = Yp, names
    return res


# ---- replicate-finding numerics of scReplicate (the clustering / graph logic around them stays R) ------
def centroid_dist(exprs_mat):
    """centroidDist, R/scReplicate.R:458-464.  exprs_mat: genes x cells of ONE cluster.  Per-gene median (rowMedians), squared
    distance of every cell to that centroid, rank(point_dist) / length (rank(): ties get their average rank)."""
    from scipy.stats import rankdata
    X = np.asarray(exprs_mat, dtype=np.float64)
    centroid = np.median(X, axis=1)
    point_dist = ((X - centroid[:, None]) ** 2).sum(axis=0)
    point_rank = rankdata(point_dist, method="average")
    return point_rank / point_rank.shape[0]


def pair_dist_mat(mat1, mat2, dist="cor"):
    """compute_dist_mat, R/scReplicate.R:470-490.  mat1 / mat2: genes x cells.  proxyC (un-vendored) published definitions:
    simil(method = "cosine" | "correlation") between the cells (rows of t(mat)) -> 1 - similarity; dist() = Euclidean."""
    A, B = np.asarray(mat1, dtype=np.float64).T, np.asarray(mat2, dtype=np.float64).T
    if dist in ("cosine", "cor"):
        if dist == "cor":
            A, B = A - A.mean(axis=1, keepdims=True), B - B.mean(axis=1, keepdims=True)
        with np.errstate(invalid="ignore", divide="ignore"):
            A = A / np.sqrt((A * A).sum(axis=1, keepdims=True))
            B = B / np.sqrt((B * B).sum(axis=1, keepdims=True))
        return 1.0 - A @ B.T
    out = np.empty((A.shape[0], B.shape[0]))
    for r0 in range(0, A.shape[0], 256):
        diff = A[r0:r0 + 256, None, :] - B[None, :, :]
        out[r0:r0 + 256] = np.sqrt((diff * diff).sum(axis=2))
    return out


def compute_dist_res(exprs_mat, cells1, res1, cells2, res2, dist="cor"):
    """compute_dist_res, R/scReplicate.R:493-520.  cells1 / cells2: column indices of the two batches' cells, res1 / res2: their
    cluster ids 1..K.  dist_res[i, j] = median(dist_mat[res1 == i, res2 == j], na.rm = TRUE), a max(res1) x max(res2) matrix."""
    X = np.asarray(exprs_mat, dtype=np.float64)
    res1, res2 = np.asarray(res1), np.asarray(res2)
    D = pair_dist_mat(X[:, cells1], X[:, cells2], dist)
    K1, K2 = int(res1.max()), int(res2.max())
    out = np.full((K1, K2), np.nan)
    for i in range(1, K1 + 1):
        for j in range(1, K2 + 1):
            blk = D[np.ix_(res1 == i, res2 == j)]
            blk = blk[~np.isnan(blk)]
            if blk.size:
                out[i - 1, j - 1] = np.median(blk)
    return out


def run_pca_scores(x, rank=10):
    """BiocSingular::runPCA(x, rank, scale = TRUE, center = TRUE, BSPARAM = ExactParam())$x (R/scReplicate.R:813-815): scores U D
    of the column-centred, unit-variance (n - 1 denominator) cells x genes matrix; columns are defined up to sign."""
    X = np.asarray(x, dtype=np.float64)
    Xs = (X - X.mean(axis=0)) / X.std(axis=0, ddof=1)
    U, S, _ = np.linalg.svd(Xs, full_matrices=False)
    r = min(rank, S.shape[0])
    return U[:, :r] * S[:r]


def kmeans_hartigan_wong(x, centers, nstart=1, iter_max=10, rng=None, init=None):
    """stats::kmeans(x, centers = K, nstart, algorithm = "Hartigan-Wong") as identifyCluster calls it (R/scReplicate.R:394-395),
    restated from the published algorithm (Hartigan & Wong 1979, AS 136; base R's stats package is not part of the reference
    tree): every start takes K distinct rows of unique(x) as centres, assigns every point to the nearest one, then sweeps the
    points in order and moves point i from its cluster a (n_a > 1) to the cluster b minimising n_b |x - c_b|^2 / (n_b + 1)
    whenever that is below n_a |x - c_a|^2 / (n_a - 1), updating both centres at once, until a whole sweep moves nothing (or
    iter_max sweeps); the start with the smallest total within-cluster sum of squares is returned.  AS 136's live-set /
    quick-transfer bookkeeping changes the order of the moves, not the set of partitions that stop it.  The random stream is
    NumPy's, not R's.  `init`: list of index arrays (one per start) instead of random draws.  Pure Python loops: small cases.
    Returns dict(cluster (1-based), centers, withinss, tot_withinss, size, iter)."""
    x = np.asarray(x, dtype=np.float64)
    m, d = x.shape
    K = int(centers)
    rng = rng or np.random.default_rng(0)
    ux, first = np.unique(x, axis=0, return_index=True)
    if ux.shape[0] < K:
        raise ValueError("more cluster centers than distinct data points.")
    best = None
    for s in range(nstart):
        idx = np.asarray(init[s]) if init is not None else np.sort(first)[rng.choice(first.shape[0], K, replace=False)]
        cen = x[idx].copy()
        lab = np.argmin(((x[:, None, :] - cen[None]) ** 2).sum(-1), axis=1)
        cnt = np.bincount(lab, minlength=K).astype(np.float64)
        if (cnt == 0).any():
            continue
        cen = np.stack([x[lab == c].mean(axis=0) for c in range(K)])
        it = 0
        for it in range(1, iter_max + 1):
            moved = False
            for i in range(m):
                a = lab[i]
                if cnt[a] <= 1:
                    continue
                d2 = ((x[i] - cen) ** 2).sum(axis=1)
                r1 = cnt[a] * d2[a] / (cnt[a] - 1)
                r2 = cnt * d2 / (cnt + 1)
                r2[a] = np.inf
                b = int(np.argmin(r2))
                if r2[b] < r1:
                    cen[a] = (cen[a] * cnt[a] - x[i]) / (cnt[a] - 1)
                    cen[b] = (cen[b] * cnt[b] + x[i]) / (cnt[b] + 1)
                    cnt[a] -= 1
                    cnt[b] += 1
                    lab[i] = b
                    moved = True
            if not moved:
                break
        cen = np.stack([x[lab == c].mean(axis=0) for c in range(K)])
        wss = np.array([((x[lab == c] - cen[c]) ** 2).sum() for c in range(K)])
        if best is None or wss.sum() < best["tot_withinss"]:
            best = {"cluster": lab + 1, "centers": cen, "withinss": wss, "tot_withinss": float(wss.sum()),
                    "size": np.bincount(lab, minlength=K), "iter": it}
    return best


def same_partition(a, b) -> bool:
    """Two label vectors describe the same partition (cluster ids are arbitrary in kmeans)."""
    a, b = np.asarray(a), np.asarray(b)
    pairs = set(zip(a.tolist(), b.tolist()))
    return len(pairs) == len(set(a.tolist())) == len(set(b.tolist()))
