"""Replicate-finding numerics of scReplicate: oracle restatement (CPU) and device parity (GPU, through the C
ABI).  Reference lines: centroidDist R/scReplicate.R:458-464, compute_dist_mat / compute_dist_res :470-520,
computePCA_byHVGMarker :793-819, kmeans(nstart = 1000) of identifyCluster :394-395."""
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


# ---- kmeans with many starts (R/scReplicate.R:394-395) ---------------------------------------------------------------------
def _blobs(seed, K, d, per, spread=3.0, noise=1.0):
    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(K, d)) * spread
    x = np.concatenate([cent[c] + noise * rng.normal(size=(per, d)) for c in range(K)])
    return x[rng.permutation(x.shape[0])]


def test_oracle_kmeans_reaches_a_hartigan_wong_stable_optimum():
    """The restated stats::kmeans: a stable partition under the transfer rule of AS 136, objective not above the best Lloyd
    solution scikit-learn finds, and exactly the planted partition on separated blobs."""
    from sklearn.cluster import KMeans
    x = _blobs(1, 4, 10, 60)
    r = orc.kmeans_hartigan_wong(x, 4, nstart=8, rng=np.random.default_rng(2))
    km = KMeans(4, n_init=20, random_state=0).fit(x)
    assert r["tot_withinss"] <= km.inertia_ * (1 + 1e-12) and orc.same_partition(r["cluster"], km.labels_)
    lab, cen, cnt = r["cluster"] - 1, r["centers"], r["size"].astype(float)
    assert np.allclose(cen, np.stack([x[lab == c].mean(axis=0) for c in range(4)]))
    for i in range(x.shape[0]):                                  # no single transfer lowers the objective
        a = lab[i]
        d2 = ((x[i] - cen) ** 2).sum(axis=1)
        r2 = cnt * d2 / (cnt + 1)
        r2[a] = np.inf
        assert r2.min() >= cnt[a] * d2[a] / (cnt[a] - 1)
    assert np.isclose(r["withinss"].sum(), r["tot_withinss"]) and r["size"].sum() == x.shape[0]
    with pytest.raises(ValueError):
        orc.kmeans_hartigan_wong(np.zeros((5, 2)), 2)            # more centres than distinct points (stats::kmeans' error)


@pytest.mark.gpu
@pytest.mark.parametrize("K,d,per,spread", [(3, 10, 80, 3.0), (5, 10, 50, 1.2), (2, 3, 40, 0.8), (8, 16, 30, 2.0)])
def test_kmeans_matches_oracle_optimum(K, d, per, spread):
    """scm_kmeans against the restated stats::kmeans: the same optimum (total within-cluster sum of squares to 1e-10, the same
    partition up to the arbitrary cluster ids), exact centres / sizes / withinss of that partition, Hartigan-Wong stability,
    and bit-identical results for a given seed."""
    import scmerge_b200 as sm
    from scmerge_b200 import replicate as rp
    x = _blobs(10 + K, K, d, per, spread=spread)
    ref = orc.kmeans_hartigan_wong(x, K, nstart=30, rng=np.random.default_rng(K))
    got = rp.kmeans(x, K, nstart=300, seed=1)
    assert got["tot_withinss"] <= ref["tot_withinss"] * (1 + 1e-10)
    if abs(got["tot_withinss"] - ref["tot_withinss"]) <= 1e-10 * ref["tot_withinss"]:
        assert orc.same_partition(got["cluster"], ref["cluster"])
    lab = got["cluster"] - 1
    assert lab.min() == 0 and lab.max() == K - 1 and np.array_equal(np.bincount(lab, minlength=K), got["size"])
    cen = np.stack([x[lab == c].mean(axis=0) for c in range(K)])
    assert np.allclose(got["centers"], cen, rtol=1e-12, atol=1e-12)
    wss = np.array([((x[lab == c] - cen[c]) ** 2).sum() for c in range(K)])
    assert np.allclose(got["withinss"], wss, rtol=1e-11) and np.isclose(got["tot_withinss"], wss.sum(), rtol=1e-12)
    cnt = got["size"].astype(float)
    for i in range(x.shape[0]):                                  # stable under the reference's transfer rule
        a = lab[i]
        d2 = ((x[i] - cen) ** 2).sum(axis=1)
        r2 = cnt * d2 / (cnt + 1)
        r2[a] = np.inf
        assert r2.min() >= cnt[a] * d2[a] / (cnt[a] - 1) * (1 - 1e-12)
    again = rp.kmeans(sm.DeviceMatrix.from_host(sm.default_context(), x), K, nstart=300, seed=1)
    assert np.array_equal(again["cluster"], got["cluster"]) and again["tot_withinss"] == got["tot_withinss"]
    one = rp.kmeans(x, 1, nstart=3)
    assert np.all(one["cluster"] == 1) and np.isclose(one["tot_withinss"], ((x - x.mean(axis=0)) ** 2).sum(), rtol=1e-12)


@pytest.mark.gpu
def test_identify_cluster_pipeline_on_device():
    """identifyCluster (R/scReplicate.R:358-447) end to end: per-batch PCA -> kmeans -> centroidDist, against the oracle pieces
    chained the same way (planted clusters: the kmeans optimum is the planted partition)."""
    from scmerge_b200 import replicate as rp
    exprs, cells, labs = _toy(seed=11, n=80, m=(90, 70), K=(3, 2))
    batch = np.concatenate([np.full(c.shape[0], b) for b, c in enumerate(cells)])
    batch = np.where(batch == 0, "b2", "b1")                     # unique(batch) order is first appearance: b2, b1
    genes = np.arange(exprs.shape[0])
    cl, dp = rp.identifyCluster(exprs, batch, None, {"b2": genes, "b1": genes}, [3, 2], nstart=200)
    for j, c in enumerate(cells):
        X = exprs[:, c]
        scores = orc.run_pca_scores(X.T, rank=10)
        ref = orc.kmeans_hartigan_wong(scores[:, :10], (3, 2)[j], nstart=20, rng=np.random.default_rng(j))
        assert orc.same_partition(cl[j], ref["cluster"])
        want = np.empty(c.shape[0])
        for g in np.unique(cl[j]):
            want[cl[j] == g] = orc.centroid_dist(X[:, cl[j] == g])
        assert np.allclose(dp[j], want)
    cl1, dp1 = rp.identifyCluster(exprs, batch, genes, None, [1, 2], nstart=50)   # a batch with kmeansK == 1 is one cluster
    assert np.all(cl1[0] == 1) and np.allclose(dp1[0], orc.centroid_dist(exprs[:, cells[0]]))
