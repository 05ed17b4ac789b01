"""CPU tests: pin the oracle on the relations the reference's own tests/docs assert
and on the committed golden fixtures."""
import os

import numpy as np
import pytest

from oracle import ruv_oracle as orc


def mean_rel_diff(a, b):
    """R's all.equal() numeric criterion: mean relative difference."""
    return np.abs(a - b).sum() / np.abs(a).sum()


@pytest.fixture(scope="module")
def sim():
    return orc.ruv_simulate(m=400, n=500, nc=400, n_celltypes=3, n_batch=2, lam=0.1,
                            rng=np.random.default_rng(12345))


def test_literal_equals_reduced(sim):
    a = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 20, return_info=True)
    b = orc.fastRUVIII_reduced(sim["Y"], sim["M"], sim["ctl"], 20)
    assert orc.rel_frobenius(b["newY"], a["newY"]) < 1e-10
    assert np.max(orc.principal_angles(b["fullalpha"][:20], a["fullalpha"][:20])) < 1e-8
    fa = orc.sign_align(b["fullalpha"][:20], a["fullalpha"][:20])
    assert orc.rel_frobenius(fa, a["fullalpha"][:20]) < 1e-8


def test_equals_classic_ruvIII(sim):
    # documented equality, R/fastRUVIII.R:30-38 (all.equal(improved1, old))
    a = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 20)
    b = orc.ruvIII_classic(sim["Y"], sim["M"], sim["ctl"], 20)
    assert mean_rel_diff(a, b) < 1.5e-8


def test_exact_vs_randomized(sim):
    # tests/testthat/test-fastRUVIII.R:11-16: expect_equal(improved1, improved2, tol = 0.02)
    a = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 20, exact=True)
    b = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 20, exact=False, svd_k=100,
                       rng=np.random.default_rng(1))
    assert mean_rel_diff(a, b) < 0.02


def test_cell_permutation_equivariance(sim):
    # tests/testthat/test-resampling-columns.R:26 (default tolerance 1.5e-8)
    perm = np.random.default_rng(3).permutation(sim["Y"].shape[0])
    a = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 20)
    b = orc.fastRUVIII(sim["Y"][perm], sim["M"][perm], sim["ctl"], 20)
    assert mean_rel_diff(a[perm], b) < 1.5e-8


def test_early_out_and_truncation(sim):
    Y, ctl = sim["Y"][:6], sim["ctl"]
    out = orc.fastRUVIII(Y, np.arange(6), ctl, 3, return_info=True)  # ncol(M) >= m
    assert out["fullalpha"] is None and np.array_equal(out["newY"], Y)
    out = orc.fastRUVIII(sim["Y"], sim["M"], ctl, 0, return_info=True)  # k == 0
    assert out["fullalpha"] is None
    out = orc.fastRUVIII(sim["Y"], sim["M"], ctl, 80, return_info=True)  # k > s silently truncates
    assert out["fullalpha"].shape[0] == 50


def test_input_checks(sim):
    Y = sim["Y"].copy()
    Y[3, 4] = np.nan
    with pytest.raises(ValueError, match="missing"):
        orc.fastRUVIII(Y, sim["M"], sim["ctl"], 5)
    Y[3, 4] = np.inf
    with pytest.raises(ValueError, match="infinities"):
        orc.fastRUVIII(Y, sim["M"], sim["ctl"], 5)


def test_residop_is_group_mean_subtraction(sim):
    g = orc.group_ids(sim["M"])
    Y0 = orc.my_residop(sim["Y"], sim["M"])
    for j in range(int(g.max()) + 1):
        assert np.allclose(Y0[g == j], sim["Y"][g == j] - sim["Y"][g == j].mean(axis=0), atol=1e-12)


def test_gram_identity_holds_for_any_shift_and_centred_gram_is_recoverable(sim):
    """The algebra the device fit relies on.  For ANY fixed vector a
        Y0'Y0 = sum_i (y_i - a)(y_i - a)' - sum_g n_g (mu_g - a)(mu_g - a)'
    (the fit uses the mean of a strided sample of cells as `a`, so that the Gram kernel can read a pre-shifted copy), and the
    centred Gram of all cells, needed by the ruvK selection, is the same correction with ONE group:
        sum_i (y_i - ybar)(y_i - ybar)' = sum_i (y_i - a)(y_i - a)' - m (ybar - a)(ybar - a)'."""
    Y, M = sim["Y"], orc.replicate_matrix(sim["M"])
    m, n = Y.shape
    g = orc.group_ids(M)
    Y0 = orc.my_residop(Y, M)                       # literal dense projector (R/fastRUVIII.R:121-129)
    G0 = Y0.T @ Y0
    ybar = Y.mean(axis=0)
    Gc = (Y - ybar).T @ (Y - ybar)
    rng = np.random.default_rng(5)
    stride = max(1, m // 64)
    for a in (np.zeros(n), ybar, Y[::stride].mean(axis=0), ybar + rng.normal(size=n)):
        Ga = (Y - a).T @ (Y - a)
        corr = np.zeros((n, n))
        for grp in range(M.shape[1]):
            rows = g == grp
            d = Y[rows].mean(axis=0) - a
            corr += rows.sum() * np.outer(d, d)
        scale = np.abs(G0).max()
        assert np.abs(Ga - corr - G0).max() / scale < 1e-12
        assert np.abs(Ga - m * np.outer(ybar - a, ybar - a) - Gc).max() / np.abs(Gc).max() < 1e-12


def test_standardize2_plain_formulas():
    rng = np.random.default_rng(0)
    Y = rng.normal(size=(30, 50)) + 3
    batch = np.array(["b%d" % (i % 3) for i in range(50)])
    sc = orc.standardize2(Y, batch)
    _lv, b = np.unique(batch, return_inverse=True)
    bm = np.stack([Y[:, b == j].mean(axis=1) for j in range(3)], axis=1)
    var = ((Y - bm[:, b]) ** 2).sum(axis=1) / (50 - 3)
    assert np.allclose(sc["stand_var"], var, rtol=1e-12)
    assert np.allclose(sc["stand_Y"], (Y - Y.mean(axis=1)[:, None]) / np.sqrt(var)[:, None], rtol=1e-12)


def test_scruviii_multik_reuses_fullalpha():
    L = orc.ruv_simulate(m=200, n=300, nc=100, rng=np.random.default_rng(1234))
    Y = np.log2(L["Y"] + 1).T
    Y = Y[Y.var(axis=1) > 0]
    ctl = L["ctl"][: Y.shape[0]]
    res = orc.scRUVIII(Y, L["M"], ctl, [5, 10], L["batch"])
    single = orc.scRUVIII(Y, L["M"], ctl, 10, L["batch"])
    assert orc.rel_frobenius(res[10]["newY"], single[10]["newY"]) < 1e-12


def test_create_chunk():
    assert orc.create_chunk(9, 4) == [(1, 4), (5, 9)]  # short tail (< chunk/2) merged
    assert orc.create_chunk(10, 4) == [(1, 4), (5, 8), (9, 10)]  # tail == chunk/2 is kept
    assert orc.create_chunk(11, 4) == [(1, 4), (5, 8), (9, 11)]
    assert orc.create_chunk(3, 4) == [(1, 3)]
    assert orc.create_chunk(8, 4) == [(1, 4), (5, 8)]


def test_pseudoruviii_chunked_equals_whole_and_getadjusted():
    P = orc.planted_factors(900, 120, 60, 5, n_types=3, n_batch=2, seed=11)
    which = (np.random.default_rng(5).permutation(900) % 4) + 1
    bulks, keys = [], []
    for b in range(2):
        sel = P["batch"] == b
        bulk, nm = orc.create_pseudobulk(P["Y"][sel].T, P["types"][sel], which[sel], k_fold=4)
        bulks.append(bulk)
        keys += [int(x.split("_")[0]) for x in nm]
    Yp = np.concatenate(bulks)
    whole = orc.pseudoRUVIII(P["Y"], Yp, np.array(keys), P["ctl"], 5, by_chunk=False)
    chunk = orc.pseudoRUVIII(P["Y"], Yp, np.array(keys), P["ctl"], 5, chunk_size=200)
    assert orc.rel_frobenius(chunk["newY"], whole["newY"]) < 1e-13
    adj = orc.getAdjustedMat(P["Y"].T, whole["fullalpha"], ctl=P["ctl"], ruvK=5)
    assert orc.rel_frobenius(adj.T, whole["newY"]) < 1e-13
    sub = np.arange(10, 40)
    adj = orc.getAdjustedMat(P["Y"].T, whole["fullalpha"], ctl=P["ctl"], ruvK=5, return_subset_genes=sub)
    assert orc.rel_frobenius(adj.T, whole["newY"][:, sub]) < 1e-13


def test_pseudobulk_means():
    rng = np.random.default_rng(2)
    X = rng.normal(size=(7, 40))
    types = np.array(["a", "b"])[rng.integers(0, 2, 40)]
    which = (np.arange(40) % 3) + 1
    bulk, names = orc.create_pseudobulk(X, types, which)
    assert names[0] == "a_1" and len(names) == 6
    sel = (which == 2) & (types == "b")
    assert np.allclose(bulk[names.index("b_2")], X[:, sel].mean(axis=1))


def test_golden_example(golden_dir):
    d = np.load(os.path.join(golden_dir, "example_sce.npz"))
    g = np.load(os.path.join(golden_dir, "golden_example.npz"))
    assert d["logcounts"].shape == (1047, 200) and int(d["ctl"].sum()) == 752
    res = orc.scRUVIII(d["logcounts"], d["cellTypes"], d["ctl"], 20, d["batch"])
    assert orc.rel_frobenius(res[20]["newY"][:16], g["newY_rows"]) < 1e-10
    assert abs(np.linalg.norm(res[20]["newY"]) / g["newY_fro"] - 1) < 1e-12
    fa = orc.sign_align(res[20]["fullalpha"][:20], g["fullalpha"][:20])
    assert orc.rel_frobenius(fa, g["fullalpha"][:20]) < 1e-8


def test_golden_sim(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_sim.npz"))
    out = orc.fastRUVIII(g["sim_Y"].astype(np.float64), g["sim_M"].astype(np.float64), g["sim_ctl"], 20)
    assert orc.rel_frobenius(out[:16], g["sim_newY_rows"]) < 1e-10
    assert abs(np.linalg.norm(out) / g["sim_newY_fro"] - 1) < 1e-12


def test_scruvg_literal_vs_reduced_and_ruv2():
    """scRUVg (R/scRUVg.R:30-85): the reference's own test pins abs(W) to ruv::RUV2's W, i.e. the first k left singular
    vectors of the centred control block (tests/testthat/test-scRUVg.R:7); the Gram-space form used on the device agrees."""
    L = orc.ruv_simulate(m=80, n=1000, nc=50, n_celltypes=10, rng=np.random.default_rng(1234))
    Y, ctl = L["Y"], orc.tological(L["ctl"], 1000)
    r = orc.scRUVg(Y, ctl, 20)
    Y0 = Y - Y.mean(axis=0)
    u = np.linalg.svd(Y0[:, ctl], full_matrices=False)[0][:, :20]       # RUV2's W (Z = 1)
    assert np.allclose(np.abs(r["W"]), np.abs(u), atol=1e-10)
    G = Y0.T @ Y0
    lam, V = np.linalg.eigh(G[np.ix_(ctl, ctl)])
    lam, V = lam[::-1][:20], V[:, ::-1][:, :20]
    B = np.zeros((1000, 20))
    B[ctl] = V / np.sqrt(lam)
    A = (V / np.sqrt(lam)).T @ G[ctl]
    assert orc.rel_frobenius(Y - (Y0 @ B) @ A, r["newY"]) < 1e-12
    # tall case takes the m x m cross-product branch (aspect ratio >= 5): same answer
    L2 = orc.ruv_simulate(m=400, n=300, nc=40, n_celltypes=4, rng=np.random.default_rng(7))
    r2 = orc.scRUVg(L2["Y"], L2["ctl"], 8)
    Y02 = L2["Y"] - L2["Y"].mean(axis=0)
    u2 = np.linalg.svd(Y02[:, orc.tological(L2["ctl"], 300)], full_matrices=False)[0][:, :8]
    assert np.allclose(np.abs(r2["W"]), np.abs(u2), atol=1e-8)
    with pytest.raises(ValueError, match="number of controls"):
        orc.scRUVg(Y, ctl, 51)


def test_silhouette_and_k_selection():
    """kBET_batch_sil / calculateSil / f_measure (R/scRUVIII.R:83-105,177-199) restated: brute-force check of the
    silhouette definition, and the selection picks the k with the largest F-score."""
    rng = np.random.default_rng(0)
    X = rng.normal(size=(60, 4))
    lab = rng.integers(0, 3, 60)
    lab[0] = 7  # singleton cluster -> s_i = 0
    s = []
    for i in range(60):
        d = np.sqrt(((X - X[i]) ** 2).sum(axis=1))
        own = lab == lab[i]
        if own.sum() == 1:
            s.append(0.0)
            continue
        a = d[own].sum() / (own.sum() - 1)
        b = min(d[lab == c].mean() for c in set(lab) if c != lab[i])
        s.append((b - a) / max(a, b))
    assert abs(np.mean(s) - orc.silhouette_avg_width(X, lab)) < 1e-14
    assert np.isnan(orc.silhouette_avg_width(X, np.zeros(60)))
    L = orc.ruv_simulate(m=120, n=200, nc=60, rng=np.random.default_rng(4))
    Y = np.log2(L["Y"] + 1).T
    keep = Y.var(axis=1) > 0
    res = orc.scRUVIII(Y[keep], L["M"], L["ctl"][keep], [2, 6], L["batch"], cell_type=L["cellTypes"])
    assert res["sil"].shape == (2, 2) and res["optimal_ruvK"] == [2, 6][int(np.argmax(res["f_score"]))]
    zo = (res["sil"] + 1) / 2
    assert np.allclose(res["f_score"], 2 * zo[0] * (1 - zo[1]) / (zo[0] + 1 - zo[1]))
    # runPCA(center, scale) is invariant to the per-gene de-standardisation: same silhouettes on the final matrices
    for i, kk in enumerate((2, 6)):
        assert np.allclose(orc.calculate_sil(res[kk]["newY"], L["cellTypes"], L["batch"]), res["sil"][:, i], atol=1e-10)


def test_golden_select(golden_dir):
    """The oracle reproduces the committed fixture for the ruvK selection and scRUVg on the reference's bundled example_sce."""
    d = np.load(os.path.join(golden_dir, "example_sce.npz"))
    g = np.load(os.path.join(golden_dir, "golden_select.npz"))
    res = orc.scRUVIII(d["logcounts"], d["cellTypes"], d["ctl"], [int(k) for k in g["ks"]], d["batch"], cell_type=d["cellTypes"])
    assert np.allclose(res["sil"], g["sil"], atol=1e-9) and np.allclose(res["f_score"], g["f_score"], atol=1e-9)
    assert res["optimal_ruvK"] == int(g["optimal"])
    r = orc.scRUVg(d["logcounts"].T, d["ctl"], 10)
    assert orc.rel_frobenius(r["newY"][:16], g["ruvg_newY_rows"]) < 1e-9
    assert np.allclose(np.abs(r["W"]).sum(axis=0), g["ruvg_absW_colsum"], rtol=1e-8)


def test_ruv1_gene_covariates(sim):
    """ruv::RUV1 restatement (call sites R/fastRUVIII.R:64, R/scRUVg.R:59, R/scMerge2_utilis.R:56): eta = NULL is the identity;
    the removed part lies in the row space of eta, and anything already in that row space on ALL genes is removed entirely;
    an intercept is added exactly when the constant is not in the span; fastRUVIII(eta) == fastRUVIII on RUV1(Y)."""
    Y, M, ctl = sim["Y"], sim["M"], sim["ctl"]
    n = Y.shape[1]
    rng = np.random.default_rng(3)
    assert orc.RUV1(Y, None, ctl) is Y
    eta = rng.normal(size=(n, 2))
    E = orc.design_matrix(eta, n, True)
    assert E.shape == (n, 3) and np.all(E[:, 0] == 1)                      # intercept put first
    assert orc.design_matrix(np.c_[np.ones(n), eta], n, True).shape == (n, 3)    # already there: not added twice
    assert orc.design_matrix(eta.T, n, False).shape == (n, 2)                    # n columns -> transposed
    assert orc.design_matrix(1, n).shape == (n, 1)
    Y1 = orc.RUV1(Y, eta, ctl)
    removed = Y - Y1
    coef = np.linalg.lstsq(E, removed.T, rcond=None)[0]
    assert np.allclose(E @ coef, removed.T, atol=1e-9)                      # removed part is in span(eta rows)
    planted = rng.normal(size=(Y.shape[0], 3)) @ E.T                        # a signal living in the row space of eta
    assert np.allclose(orc.RUV1(Y + planted, eta, ctl), Y1, atol=1e-8)       # ... is removed completely
    a = orc.fastRUVIII(Y, M, ctl, 5, eta=eta)
    b = orc.fastRUVIII(Y1, M, ctl, 5)
    assert np.array_equal(a, b)
    g = orc.scRUVg(Y, ctl, 4, eta=eta)
    assert np.allclose(g["newY"], Y - g["W"] @ g["alpha"])                  # the subtraction is from the ORIGINAL Y (R/scRUVg.R:80)
    g2 = orc.scRUVg(Y, ctl, 3, fullW=g["W"])                                # re-use of a decomposition (:64, :77)
    assert np.allclose(g2["W"], g["W"][:, :3])
