"""Replicate-finding numerics of scReplicate: oracle restatement (CPU) and device parity (GPU, through the C
ABI).  Reference lines: centroidDist R/scReplicate.R:458-464, compute_dist_mat / compute_dist_res :470-520,
computePCA_byHVGMarker :793-819."""
import numpy as np
import pytest

from oracle import ruv_oracle as orc


def _toy(seed=3, n=60, m=(70, 55), K=(3, 4)):
    rng = np.random.default_rng(seed)
    mats, labs = [], []
    for b, (mb, Kb) in enumerate(zip(m, K)):
        lab = rng.integers(1, Kb + 1, size=mb)
        lab[:Kb] = np.arange(1, Kb + 1)                      # every cluster id occurs
        centres = rng.normal(size=(Kb, n)) * 2.0
        X = centres[lab - 1] + rng.normal(size=(mb, n)) + 0.5 * b
        mats.append(np.log2(np.abs(X) + 1.0))
        labs.append(lab)
    exprs = np.concatenate(mats, axis=0).T                   # genes x cells
    cells = [np.arange(0, m[0]), np.arange(m[0], m[0] + m[1])]
    return exprs, cells, labs


# ---- oracle self-checks (CPU) ----------------------------------------------------------------------------------------------
def test_oracle_centroid_dist_definition():
    exprs, cells, labs = _toy()
    X = exprs[:, cells[0]][:, labs[0] == 1]
    got = orc.centroid_dist(X)
    med = np.sort(X, axis=1)
    c = X.shape[1]
    centroid = med[:, c // 2] if c % 2 else 0.5 * (med[:, c // 2 - 1] + med[:, c // 2])
    d = ((X - centroid[:, None]) ** 2).sum(axis=0)
    assert got.shape == (c,) and got.min() > 0 and np.isclose(got.max(), 1.0)
    assert np.array_equal(np.argsort(got, kind="stable"), np.argsort(d, kind="stable"))
    X2 = np.concatenate([X, X[:, :1]], axis=1)               # a tie: average ranks
    g2 = orc.centroid_dist(X2)
    assert g2[0] == g2[-1]


def test_oracle_dist_and_block_medians():
    exprs, cells, labs = _toy()
    for metric in ("cor", "cosine", "euclidean"):
        D = orc.pair_dist_mat(exprs[:, cells[0]], exprs[:, cells[1]], metric)
        assert D.shape == (70, 55)
        a, b = exprs[:, cells[0][3]], exprs[:, cells[1][7]]
        if metric == "cor":
            want = 1.0 - np.corrcoef(a, b)[0, 1]
        elif metric == "cosine":
            want = 1.0 - a @ b / np.linalg.norm(a) / np.linalg.norm(b)
        else:
            want = np.linalg.norm(a - b)
        assert np.isclose(D[3, 7], want, rtol=1e-12)
        R = orc.compute_dist_res(exprs, cells[0], labs[0], cells[1], labs[1], metric)
        assert R.shape == (3, 4)
        assert np.isclose(R[1, 2], np.median(D[np.ix_(labs[0] == 2, labs[1] == 3)]))
    ex2 = exprs.copy()
    ex2[:, cells[0][0]] = 1.0                                # a constant cell: correlation undefined -> NaN, ignored by the median
    D = orc.pair_dist_mat(ex2[:, cells[0]], ex2[:, cells[1]], "cor")
    assert np.isnan(D[0]).all() and not np.isnan(D[1:]).any()
    assert not np.isnan(orc.compute_dist_res(ex2, cells[0], labs[0], cells[1], labs[1], "cor")).any()


def test_oracle_run_pca_scores_is_prcomp():
    rng = np.random.default_rng(1)
    X = rng.normal(size=(40, 12)) @ rng.normal(size=(12, 12))
    S = orc.run_pca_scores(X, rank=5)
    Xs = (X - X.mean(0)) / X.std(0, ddof=1)
    w, V = np.linalg.eigh(Xs.T @ Xs)
    V = V[:, ::-1][:, :5]
    assert np.allclose(np.abs(S), np.abs(Xs @ V), atol=1e-10)
    assert np.allclose((S ** 2).sum(axis=0), w[::-1][:5], rtol=1e-10)


# ---- device parity (GPU) ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_centroid_dist_matches_oracle():
    import scmerge_b200 as sm
    exprs, cells, labs = _toy(n=257, m=(900, 333), K=(3, 5))
    Xb = exprs[:, cells[0]]
    rank, dist = sm.cluster_centroid_dist(Xb, labs[0])
    for g in range(1, 4):
        sel = labs[0] == g
        ref = orc.centroid_dist(Xb[:, sel])
        assert np.array_equal(rank[sel], ref)                                     # ranks: exact (an ordering)
        centroid = np.median(Xb[:, sel], axis=1)
        assert np.allclose(dist[sel], ((Xb[:, sel] - centroid[:, None]) ** 2).sum(axis=0), rtol=1e-13)
    one = sm.centroidDist(Xb[:, labs[0] == 2])
    assert np.array_equal(one, orc.centroid_dist(Xb[:, labs[0] == 2]))
    # an even and an odd cluster size, a singleton cluster, many gene blocks (nb < n)
    lab2 = np.array([1] * 10 + [2] * 7 + [3])
    r2, _ = sm.cluster_centroid_dist(Xb[:, :18], lab2)
    for g in (1, 2, 3):
        assert np.array_equal(r2[lab2 == g], orc.centroid_dist(Xb[:, :18][:, lab2 == g]))


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["cor", "cosine", "euclidean"])
def test_pair_distances_and_block_medians_match_oracle(metric):
    import scmerge_b200 as sm
    exprs, cells, labs = _toy(n=130, m=(301, 222), K=(3, 4))
    D = sm.compute_dist_mat(exprs[:, cells[0]], exprs[:, cells[1]], metric)
    Dref = orc.pair_dist_mat(exprs[:, cells[0]], exprs[:, cells[1]], metric)
    assert np.allclose(D, Dref, rtol=1e-10, atol=1e-12)
    R = sm.compute_dist_res(exprs, cells[0], labs[0], cells[1], labs[1], metric)
    Rref = orc.compute_dist_res(exprs, cells[0], labs[0], cells[1], labs[1], metric)
    assert R.shape == Rref.shape == (3, 4)
    assert np.allclose(R, Rref, rtol=1e-10, atol=1e-12)
    assert np.array_equal(R.argmin(axis=1), Rref.argmin(axis=1)) and np.array_equal(R.argmin(axis=0), Rref.argmin(axis=0))  # what findMNC uses
    if metric == "cor":
        ex2 = exprs.copy()
        ex2[:, cells[0][5]] = 2.0                                                  # constant cell -> NaN row, ignored (na.rm = TRUE)
        R2 = sm.compute_dist_res(ex2, cells[0], labs[0], cells[1], labs[1], metric)
        assert np.allclose(R2, orc.compute_dist_res(ex2, cells[0], labs[0], cells[1], labs[1], metric), rtol=1e-10)
        lab_gap = labs[1].copy()
        lab_gap[lab_gap == 2] = 1                                                  # a cluster id without cells: NaN block
        R3 = sm.compute_dist_res(exprs, cells[0], labs[0], cells[1], lab_gap, metric)
        R3ref = orc.compute_dist_res(exprs, cells[0], labs[0], cells[1], lab_gap, metric)
        assert np.isnan(R3[:, 1]).all() and np.allclose(np.nan_to_num(R3), np.nan_to_num(R3ref), rtol=1e-10)


@pytest.mark.gpu
def test_per_batch_pca_matches_oracle():
    import scmerge_b200 as sm
    exprs, cells, labs = _toy(n=200, m=(500, 400), K=(3, 3), seed=9)
    batch = np.array([0] * 500 + [1] * 400)
    hvg = [np.arange(0, 200, 2), np.arange(10, 150)]
    for b in (0, 1):
        S = sm.computePCA_byHVGMarker(b, batch, [], None, exprs, hvg)
        ref = orc.run_pca_scores(exprs[np.ix_(hvg[b], np.nonzero(batch == b)[0])].T, rank=10)
        assert S.shape == ref.shape == ((500, 400)[b], 10)
        sg = np.sign((S * ref).sum(axis=0))
        assert np.allclose(S * sg, ref, rtol=0, atol=1e-7 * np.abs(ref).max())
    assert sm.computePCA_byHVGMarker(1, batch, [1], None, exprs, hvg) is None      # kmeansK == 1: NULL
    S = sm.computePCA_byHVGMarker(0, batch, [], np.arange(30), exprs, hvg)         # user-supplied markers win over the HVGs
    ref = orc.run_pca_scores(exprs[:30, :500].T, rank=10)
    sg = np.sign((S * ref).sum(axis=0))
    assert np.allclose(S * sg, ref, rtol=0, atol=1e-7 * np.abs(ref).max())
