"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every device result comes through the C ABI
(ctypes -> libscmerge_b200.so) and is compared with the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): adjusted matrix relative Frobenius error <= 1e-8 (sign-invariant by
construction), principal angles of the alpha / W subspaces ~ 0, group / pseudobulk indices and counts bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import ruv_oracle as orc

pytestmark = pytest.mark.gpu

TOL_NEWY = 1e-8


@pytest.fixture(scope="module")
def sm():
    import scmerge_b200 as s
    return s


@pytest.fixture(scope="module")
def ctx(sm):
    return sm.default_context()


def mean_rel_diff(a, b):
    return np.abs(a - b).sum() / np.abs(a).sum()


# ---- stage level ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,G", [(1000, 257, 7), (5000, 512, 1), (3000, 130, 700), (64, 33, 64), (20000, 96, 3)])
def test_seg_colsum(sm, ctx, m, n, G):
    from scmerge_b200.ruv import segmented_sums
    rng = np.random.default_rng(m + n)
    Y = rng.normal(size=(m, n)) * 3 + 1
    keys = rng.integers(0, G, size=m).astype(np.int32)
    keys[rng.random(m) < 0.05] = -1  # skipped rows
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    sums, cnt, flag = segmented_sums(dY, keys, G, check_finite=True)
    ref = np.zeros((G, n))
    valid = keys >= 0
    np.add.at(ref, keys[valid], Y[valid])
    assert np.array_equal(cnt, np.bincount(keys[valid], minlength=G))  # bit-exact counts
    assert flag == 0
    assert np.allclose(sums, ref, rtol=1e-12, atol=1e-11)
    # squared deviations about per-group centres + determinism
    centre = ref / np.maximum(cnt, 1)[:, None]
    ssq, _ = segmented_sums(dY, keys, G, center=centre)
    ref2 = np.zeros((G, n))
    np.add.at(ref2, keys[valid], (Y[valid] - centre[keys[valid]]) ** 2)
    assert np.allclose(ssq, ref2, rtol=1e-11, atol=1e-11)
    sums2, _ = segmented_sums(dY, keys, G)
    assert np.array_equal(sums, sums2)  # bit-identical run to run
    dY.free()


def test_seg_colsum_nonfinite_flag(sm, ctx):
    from scmerge_b200.ruv import segmented_sums
    Y = np.ones((300, 40))
    Y[17, 5] = np.nan
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    assert segmented_sums(dY, np.zeros(300, dtype=np.int32), 1, check_finite=True)[2] == 1
    Y[17, 5] = -np.inf
    dY2 = sm.DeviceMatrix.from_host(ctx, Y)
    assert segmented_sums(dY2, np.zeros(300, dtype=np.int32), 1, check_finite=True)[2] == 1


@pytest.mark.parametrize("m,n,shift", [(1000, 128, False), (777, 300, True), (5000, 257, True), (40, 513, False),
                                       (70000, 260, True), (3000, 2000, True), (1500, 3000, True), (900, 240, True),
                                       (500, 64, False), (333, 1000, True), (2000, 56, True), (129, 208, True)])
def test_gram(sm, ctx, m, n, shift):
    """K3 against NumPy for every unit type of the TMA kernel: full tiles, the triangular diagonal units, the thin last tile
    row (n mod 128 <= 80: 300, 257, 513, 260, 2000, 3000, 208) and a padded last row that stays full (240, 1000)."""
    from scmerge_b200 import _lib as L
    rng = np.random.default_rng(m * 7 + n)
    Y = rng.normal(size=(m, n)) + 2.0
    a = Y.mean(axis=0) if shift else None
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    dG = ctx.empty((n, n))
    da = ctx.to_device(a) if shift else None
    L.check(L.lib().scm_gram(ctx.handle, dY.ptr, m, n, dY.ld, da.ptr if da else None, dG.ptr))
    G = dG.to_host()
    Z = Y - a if shift else Y
    ref = Z.T @ Z
    assert np.allclose(G, G.T, rtol=0, atol=0)  # exactly symmetric (mirrored tiles)
    assert np.linalg.norm(G - ref) / np.linalg.norm(ref) < 1e-13


def test_gram_residual_matches_centred_gram(sm, ctx):
    from scmerge_b200 import _lib as L
    from scmerge_b200.ruv import segmented_sums
    rng = np.random.default_rng(5)
    m, n, G = 4000, 200, 9
    Y = rng.normal(size=(m, n)) + rng.normal(size=n) * 4
    g = rng.integers(0, G, size=m).astype(np.int32)
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    sums, cnt = segmented_sums(dY, g, G)
    a = Y.mean(axis=0)
    scale = 1.0 / (1.0 + rng.random(n))
    dG, da, ds = ctx.empty((n, n)), ctx.to_device(a), ctx.to_device(scale)
    dsums, dcnt = ctx.to_device(sums), ctx.to_device(cnt)
    lib = L.lib()
    L.check(lib.scm_gram(ctx.handle, dY.ptr, m, n, dY.ld, da.ptr, dG.ptr))
    L.check(lib.scm_gram_residual(ctx.handle, dG.ptr, n, dsums.ptr, dcnt.ptr, G, da.ptr, ds.ptr))
    Y0 = (Y - (sums / cnt[:, None])[g]) * scale
    ref = Y0.T @ Y0
    assert np.linalg.norm(dG.to_host() - ref) / np.linalg.norm(ref) < 1e-12


@pytest.mark.parametrize("n,s,mode", [(300, 20, 0), (1047, 50, 0), (500, 30, 1)])
def test_eig_topk(sm, ctx, n, s, mode):
    from scmerge_b200 import _lib as L
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    lam = np.concatenate([np.linspace(100, 20, s), np.linspace(5, 0.01, n - s)])
    Gm = (Q * lam) @ Q.T
    Gm = 0.5 * (Gm + Gm.T)
    dG = ctx.to_device(Gm)
    dvals, dvecs = ctx.empty((s,)), ctx.empty((s, n))
    it, res = C.c_int32(), C.c_double()
    rc = L.lib().scm_eig_topk(ctx.handle, dG.ptr, n, s, s, mode, 1e-11, 500, 7, dvals.ptr, dvecs.ptr, C.byref(it), C.byref(res))
    L.check(rc)
    vals, vecs = dvals.to_host(), dvecs.to_host()
    if mode == 0:
        assert np.allclose(vals, lam[:s], rtol=1e-10)
        ang = orc.principal_angles(vecs, Q[:, :s].T)
        assert ang.max() < 1e-7
        assert np.allclose(vecs @ vecs.T, np.eye(s), atol=1e-10)
    else:  # randomized: 3 products, loose
        assert np.allclose(vals[:5], lam[:5], rtol=1e-2)


# ---- end to end through the reference-facing functions ---------------------------------------------------
@pytest.fixture(scope="module")
def sim():
    return orc.ruv_simulate(m=400, n=500, nc=400, n_celltypes=3, n_batch=2, lam=0.1, rng=np.random.default_rng(12345))


def test_fastruviii_exact_vs_oracle(sm, sim):
    ref = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 20, return_info=True)
    out = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=20, return_info=True)
    assert out["fullalpha"].shape == ref["fullalpha"].shape == (50, 500)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(out["fullalpha"][:20], ref["fullalpha"][:20]).max() < 1e-6
    # singular values (row norms of fullalpha) of the well separated part of the spectrum
    rn, rr = np.linalg.norm(out["fullalpha"], axis=1), np.linalg.norm(ref["fullalpha"], axis=1)
    assert np.allclose(rn[:20], rr[:20], rtol=1e-8)


def test_fastruviii_label_vector_and_fortran_input(sm, sim):
    keys = orc.group_ids(sim["M"])
    a = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=10)
    b = sm.fastRUVIII(np.asfortranarray(sim["Y"]), keys, np.nonzero(sim["ctl"])[0], k=10)  # R layout, labels, int ctl
    assert orc.rel_frobenius(b, a) < 1e-12


def test_fastruviii_randomized_close_to_exact(sm, sim):
    # tests/testthat/test-fastRUVIII.R:11-16, tol = 0.02 (mean relative difference)
    a = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=20)
    # the non-exact branch ignores svd_k and asks for min(m - G, n_ctl) rows (R/fastRUVIII.R:69): all of them are returned
    b = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=20, BSPARAM=sm.RandomParam(), svd_k=100, return_info=True)
    assert b["fullalpha"].shape[0] == min(sim["Y"].shape[0] - sim["M"].shape[1], int(np.sum(sim["ctl"])))
    assert mean_rel_diff(a, b["newY"]) < 0.02


def test_cell_permutation_equivariance(sm, sim):
    # tests/testthat/test-resampling-columns.R:26 (tolerance 1.5e-8)
    perm = np.random.default_rng(3).permutation(sim["Y"].shape[0])
    a = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=20)
    b = sm.fastRUVIII(sim["Y"][perm], sim["M"][perm], sim["ctl"], k=20)
    assert mean_rel_diff(a[perm], b) < 1.5e-8


def test_fastruviii_reuses_fullalpha_and_truncates_k(sm, sim):
    first = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=20, return_info=True)
    again = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=5, fullalpha=first["fullalpha"])
    ref = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 5)
    assert orc.rel_frobenius(again, ref) < TOL_NEWY
    big = sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=60, fullalpha=first["fullalpha"])  # k > s -> s rows
    ref60 = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 60)
    assert mean_rel_diff(ref60, big) < 0.05  # rows 21..50 live in the noise bulk: only loosely pinned


def test_early_out_and_errors(sm, sim):
    Y = sim["Y"][:6]
    out = sm.fastRUVIII(Y, np.arange(6), sim["ctl"], k=3, return_info=True)  # ncol(M) >= m
    assert out["fullalpha"] is None and out["newY"] is Y
    assert sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], k=0) is sim["Y"]
    bad = sim["Y"].copy()
    bad[3, 4] = np.nan
    with pytest.raises(ValueError, match="missing values or infinities"):
        sm.fastRUVIII(bad, sim["M"], sim["ctl"], k=5)
    with pytest.raises(ValueError, match="one-hot"):
        sm.fastRUVIII(sim["Y"], np.ones((400, 2)), sim["ctl"], k=5)
    with pytest.raises(TypeError):
        sm.fastRUVIII(sim["Y"], sim["M"], sim["ctl"])


def test_golden_example_scruviii(sm, golden_dir):
    """BASELINE.json configs[0]: bundled example data, mouse scSEG controls, ruvK = 20, via scRUVIII."""
    d = np.load(os.path.join(golden_dir, "example_sce.npz"))
    g = np.load(os.path.join(golden_dir, "golden_example.npz"))
    res = sm.scRUVIII(d["logcounts"], d["cellTypes"], d["ctl"], k=20, batch=d["batch"])
    ref = orc.scRUVIII(d["logcounts"], d["cellTypes"], d["ctl"], 20, d["batch"])
    assert orc.rel_frobenius(res[20]["newY"], ref[20]["newY"]) < TOL_NEWY
    assert orc.rel_frobenius(res[20]["newY"][:16], g["newY_rows"]) < TOL_NEWY
    assert np.allclose(res["stand_mean"], g["stand_mean"], rtol=1e-12)
    assert np.allclose(res["stand_sd"], g["stand_sd"], rtol=1e-11)
    assert orc.principal_angles(res[20]["fullalpha"][:20], g["fullalpha"][:20]).max() < 1e-6


def test_scruviii_multi_k(sm):
    """scRUVIII with several k (R/scRUVIII.R:64-114): every adjusted matrix, the two silhouette coefficients per k, the
    F-scores and the chosen k against the oracle (exact PCA + all-pairs silhouette on the materialised matrices)."""
    L_ = orc.ruv_simulate(m=200, n=300, nc=100, rng=np.random.default_rng(1234))
    Y = np.log2(L_["Y"] + 1).T
    keep = Y.var(axis=1) > 0
    Y, ctl = Y[keep], L_["ctl"][keep]
    ks = [2, 5, 10]
    res = sm.scRUVIII(Y, L_["M"], ctl, k=ks, batch=L_["batch"], cell_type=L_["cellTypes"])
    ref = orc.scRUVIII(Y, L_["M"], ctl, ks, L_["batch"], cell_type=L_["cellTypes"])
    for kk in ks:
        assert orc.rel_frobenius(res[kk]["newY"], ref[kk]["newY"]) < TOL_NEWY
    assert np.allclose(res["sil"], ref["sil"], rtol=0, atol=1e-7)
    assert np.allclose(res["f_score"], ref["f_score"], rtol=0, atol=1e-7)
    assert res["optimal_ruvK"] == ref["optimal_ruvK"]
    # cell_type = NULL: the replicate structure stands in (R/scRUVIII.R:89-92); only the optimal matrix is produced
    one = sm.scRUVIII(Y, L_["M"], ctl, k=ks, batch=L_["batch"], return_all_RUV=False)
    ref2 = orc.scRUVIII(Y, L_["M"], ctl, ks, L_["batch"])
    assert one["optimal_ruvK"] == ref2["optimal_ruvK"]
    assert orc.rel_frobenius(one["newY"], ref2[ref2["optimal_ruvK"]]["newY"]) < TOL_NEWY


@pytest.mark.parametrize("m,d,Ca,Cb", [(500, 10, 4, 3), (1300, 3, 7, 1), (257, 16, 2, 2), (900, 10, 12, 5)])
def test_silhouette_vs_oracle(sm, ctx, m, d, Ca, Cb):
    """scm_silhouette == summary(cluster::silhouette(labels, dist(x)))$avg.width for both labelings from one distance pass;
    singleton clusters (s_i = 0), unused label ids and a split over cell ranges (the multi-rank decomposition) included."""
    rng = np.random.default_rng(m + d)
    la = rng.integers(0, Ca - 1, m).astype(np.int32)       # id Ca - 1 is used exactly once -> singleton cluster
    la[5] = Ca - 1
    lb = rng.integers(0, Cb, m).astype(np.int32)
    if Cb > 2:
        lb[lb == 1] = 0                                     # label id 1 unused
    X = rng.normal(size=(m, d)) + 1.5 * rng.normal(size=(Ca, d))[la]
    dX, dla, dlb = ctx.to_device(X), ctx.to_device(la), ctx.to_device(lb)
    sa, sb = C.c_double(), C.c_double()
    lib = sm.lib()

    def run(i0, i1, with_b=True):
        rc = lib.scm_silhouette(ctx.handle, dX.ptr, m, d, d, dla.ptr, Ca, dlb.ptr if with_b else None, Cb, i0, i1,
                                C.byref(sa), C.byref(sb))
        assert rc == 0, lib.scm_last_error()
        return sa.value, sb.value
    ta, tb = run(0, m)
    assert abs(ta / m - orc.silhouette_avg_width(X, la)) < 1e-12
    if len(np.unique(lb)) > 1:
        assert abs(tb / m - orc.silhouette_avg_width(X, lb)) < 1e-12
    else:
        assert tb == 0.0
    pa, pb = run(0, m // 3)
    qa, qb = run(m // 3, m)
    assert abs(pa + qa - ta) < 1e-9 and abs(pb + qb - tb) < 1e-9
    assert abs(run(0, m, with_b=False)[0] - ta) < 1e-9
    assert run(0, m) == (ta, tb)                            # deterministic
    bad = ctx.to_device(np.full(m, Ca, dtype=np.int32))
    assert lib.scm_silhouette(ctx.handle, dX.ptr, m, d, d, bad.ptr, Ca, None, 1, 0, m, C.byref(sa), C.byref(sb)) == -1


def test_pca_scores_of_adjusted_matrix(sm, ctx):
    """scm_pca_scores: runPCA(newY, rank = 10, center, scale)$x of the adjusted matrix from the centred Gram of Y and the
    rank-k factors, without materialising newY -- against the SVD of the materialised, centred, scaled oracle matrix."""
    from scmerge_b200.ruv import _Coef, _fit
    P = orc.planted_factors(4000, 200, 60, 6, n_types=4, n_batch=3, seed=11)
    Y, ctlm = P["Y"], orc.tological(P["ctl"], 200)
    ref = orc.fastRUVIII(Y, P["types"], ctlm, 6, return_info=True)
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    fit = _fit(ctx, dY, P["types"].astype(np.int32), 4, 50, 6, True, keep_gram=True)
    Yc = Y - Y.mean(axis=0)
    assert np.allclose(fit.d_gram.to_host(), Yc.T @ Yc, rtol=1e-10, atol=1e-8 * np.abs(Yc.T @ Yc).max())
    assert np.allclose(fit.d_colmean.to_host(), Y.mean(axis=0), rtol=1e-13)
    lib = sm.lib()
    for coef_k in (6, None):
        target = ref["newY"] if coef_k else Y
        Xs = (target - target.mean(axis=0)) / target.std(axis=0, ddof=1)
        U, S, Vt = np.linalg.svd(Xs, full_matrices=False)
        want = U[:, :10] * S[:10]
        d_scores, d_sing = ctx.empty((4000, 10)), ctx.empty((10,))
        it, rs = C.c_int32(), C.c_double()

        def call(h):
            rc = lib.scm_pca_scores(ctx.handle, fit.d_gram.ptr, 200, 4000, h, fit.d_colmean.ptr, 10, dY.ptr, 4000, dY.ld,
                                    d_scores.ptr, d_sing.ptr, 1e-10, 500, C.byref(it), C.byref(rs))
            assert rc == 0, lib.scm_last_error()
        if coef_k:
            with _Coef(ctx, 200, fit.fullalpha, coef_k, ctlm) as coef:
                call(coef.handle)
        else:
            call(None)
        got = d_scores.to_host()
        assert np.allclose(d_sing.to_host(), S[:10], rtol=1e-9)
        sgn = np.sign(np.sum(got * want, axis=0))
        gaps = np.minimum(np.abs(np.diff(S[:11])), np.r_[np.inf, np.abs(np.diff(S[:10]))]) / S[0]
        ok = gaps > 1e-3                                     # individual PCs are defined where the spectrum is separated
        assert ok.sum() >= 3
        assert np.allclose((got * sgn)[:, ok], want[:, ok], rtol=0, atol=1e-7 * np.abs(want).max())
        # pairwise distances (all the silhouette sees) are invariant to rotations inside near-degenerate PC groups
        idx = np.random.default_rng(0).integers(0, 4000, (200, 2))
        dg = np.linalg.norm(got[idx[:, 0]] - got[idx[:, 1]], axis=1)
        dw = np.linalg.norm(want[idx[:, 0]] - want[idx[:, 1]], axis=1)
        assert np.allclose(dg, dw, rtol=1e-7, atol=1e-9)


def _pseudo_case(seed=777, m=1200, n=300, c=120, k=8):
    P = orc.planted_factors(m, n, c, k, n_types=4, n_batch=3, seed=seed)
    which = (np.random.default_rng(5).permutation(m) % 6) + 1
    return P, which


def test_pseudobulk_and_pseudoruviii(sm, golden_dir):
    P, which = _pseudo_case()
    bulks, keys, bulks_ref = [], [], []
    for b in range(3):
        sel = P["batch"] == b
        bulk, names = sm.create_pseudoBulk(P["Y"][sel].T, P["types"][sel], k_fold=6, which=which[sel])
        ref, names_ref = orc.create_pseudobulk(P["Y"][sel].T, P["types"][sel], which[sel], k_fold=6)
        assert names == names_ref  # pseudobulk indices bit-exact
        assert np.allclose(bulk, ref, rtol=1e-13, atol=1e-13)
        bulks.append(bulk)
        bulks_ref.append(ref)
        keys += [int(x.split("_")[0]) for x in names]
    Yp = np.concatenate(bulks)
    out = sm.pseudoRUVIII(P["Y"], Yp, np.array(keys), P["ctl"], k=8, return_info=True)
    ref = orc.pseudoRUVIII(P["Y"], np.concatenate(bulks_ref), np.array(keys), P["ctl"], 8)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < TOL_NEWY
    assert np.allclose(out["adjusted_means"], ref["adjusted_means"], rtol=1e-13)
    g = np.load(os.path.join(golden_dir, "golden_sim.npz"))
    assert orc.rel_frobenius(out["newY"][:16], g["ps_newY_rows"]) < TOL_NEWY
    # subset of genes + getAdjustedMat from the stored fullalpha
    sub = np.arange(10, 90, 3)
    o2 = sm.pseudoRUVIII(P["Y"], Yp, np.array(keys), P["ctl"], k=8, subset=sub, fullalpha=out["fullalpha"])
    assert orc.rel_frobenius(o2, ref["newY"][:, sub]) < TOL_NEWY
    adj = sm.getAdjustedMat(P["Y"].T, out["fullalpha"], ctl=P["ctl"], ruvK=8)
    assert orc.rel_frobenius(adj.T, ref["newY"]) < TOL_NEWY
    adj = sm.getAdjustedMat(P["Y"].T, out["fullalpha"], ctl=P["ctl"], ruvK=8, return_subset_genes=sub)
    assert orc.rel_frobenius(adj.T, ref["newY"][:, sub]) < TOL_NEWY


@pytest.mark.parametrize("k", [5, 20, 33, 50])
def test_k_sweep_planted(sm, k):
    """ruvK sweep (BASELINE.json configs[4] shape family) on planted factors with a spectral gap after k."""
    P = orc.planted_factors(6000, 400, 150, k, n_types=5, n_batch=3, seed=100 + k)
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], k, return_info=True)
    out = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=k, return_info=True)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(out["fullalpha"][:k], ref["fullalpha"][:k]).max() < 1e-6


def test_host_pipeline_multichunk(sm, monkeypatch):
    """scm_ruv3_host with the upload / download cut into several chunks (fit accumulates behind the H2D copies,
    shift = mean of the first chunk only) must equal the oracle and the single-chunk run."""
    P = orc.planted_factors(5000, 260, 100, 7, n_types=5, n_batch=3, seed=5)
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 7, return_info=True)
    one = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=7)
    monkeypatch.setenv("SCM_HOST_CHUNK_ROWS", "1024")
    Yp, out = sm.pinned_empty(P["Y"].shape), sm.pinned_empty(P["Y"].shape)
    Yp[:] = P["Y"]
    res = sm.fastRUVIII(Yp, P["types"], P["ctl"], k=7, return_info=True, out=out)
    assert res["newY"] is out
    assert orc.rel_frobenius(out, ref["newY"]) < TOL_NEWY
    assert orc.rel_frobenius(out, one) < 1e-11
    # fullalpha re-use (adjust only) and the scRUVIII / pseudoRUVIII host paths, still chunked
    again = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=4, fullalpha=res["fullalpha"])
    assert orc.rel_frobenius(again, orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 4)) < TOL_NEWY
    batch = np.array([f"b{b}" for b in P["batch"]])
    Yg = np.asfortranarray(P["Y"].T)  # genes x cells, column-major: R's layout
    sc = sm.scRUVIII(Yg, P["types"], P["ctl"], k=7, batch=batch)
    ref_sc = orc.scRUVIII(P["Y"].T, P["types"], P["ctl"], 7, batch)
    assert orc.rel_frobenius(sc[7]["newY"], ref_sc[7]["newY"]) < TOL_NEWY
    assert np.allclose(sc["stand_sd"], ref_sc["stand_sd"], rtol=1e-11)
    adj = sm.pseudoRUVIII(P["Y"], P["Y"][:300], P["types"][:300], P["ctl"], k=5, fullalpha=res["fullalpha"], return_info=True)
    ref_adj = orc.pseudoRUVIII(P["Y"], P["Y"][:300], P["types"][:300], P["ctl"], 5, fullalpha=ref["fullalpha"])
    assert orc.rel_frobenius(adj["newY"], ref_adj["newY"]) < TOL_NEWY
    assert np.allclose(adj["adjusted_means"], P["Y"].mean(axis=0), rtol=1e-12)
    # getAdjustedMat on R's layout: with the stored adjusted_means the call streams in ONE pass (upload / adjust / download of
    # different chunks at the same time), without them the means accumulate during the upload first; both equal the oracle,
    # the resident path, and each other's bits
    ctlm = orc.tological(P["ctl"], P["Y"].shape[1])
    ref_g = orc.getAdjustedMat(P["Y"].T, ref["fullalpha"], ctl=ctlm, ruvK=5)
    g1 = sm.getAdjustedMat(Yg, res["fullalpha"], ctl=ctlm, ruvK=5)
    g2 = sm.getAdjustedMat(Yg, res["fullalpha"], ctl=ctlm, ruvK=5, adjusted_means=adj["adjusted_means"])
    g3 = sm.getAdjustedMat(np.ascontiguousarray(Yg), res["fullalpha"], ctl=ctlm, ruvK=5)      # C-ordered: resident path
    assert g1.shape == ref_g.shape and orc.rel_frobenius(g1, ref_g) < TOL_NEWY and orc.rel_frobenius(g3, ref_g) < TOL_NEWY
    assert np.array_equal(g1, g2)
    bad = P["Y"].copy()
    bad[4321, 17] = np.inf  # NA/Inf in a late chunk is still caught (R/fastRUVIII.R:52-60)
    with pytest.raises(ValueError, match="missing values or infinities"):
        sm.fastRUVIII(bad, P["types"], P["ctl"], k=7)


def test_kernels_are_race_free_and_bit_reproducible(sm):
    """Hundreds of launches of the pipelined (TMA + mbarrier) kernels on the same input must give bit-identical
    results: a rare producer/consumer race shows up as one wrong 16-gene block in ~1 % of the launches
    (this is how the cp.async.bulk fragment staging of the adjust kernel was caught)."""
    import torch
    from scmerge_b200 import _lib as L
    from scmerge_b200 import ruv
    torch.cuda.set_device(0)
    ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    m, n, c, k = 32768, 1000, 300, 10
    P = orc.planted_factors(m, n, c, k, n_types=12, n_batch=4, seed=321)
    dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
    keys, G = ruv.replicate_keys(P["types"])
    fit = ruv._fit(ctx, dY, keys, G, 50, k, True)
    lib = L.lib()
    d_alpha = ctx.to_device(np.ascontiguousarray(fit.fullalpha[:k]))
    d_ctl = ctx.to_device(P["ctl"].astype(np.uint8))
    coef = C.c_void_p()
    L.check(lib.scm_coef_create(ctx.handle, d_alpha.ptr, n, k, d_ctl.ptr, None, None, None, C.byref(coef)))
    dO, dW, dG = sm.DeviceMatrix(ctx, m, n), sm.DeviceMatrix(ctx, m, n), ctx.empty((n, n))
    tY, tO, tW, tG = (torch.as_tensor(x, device="cuda") for x in (dY.buf, dO.buf, dW.buf, dG))
    L.check(lib.scm_adjust(ctx.handle, dY.ptr, m, n, dY.ld, coef, None, 0, dO.ptr, dO.ld, None))
    L.check(lib.scm_gram(ctx.handle, dY.ptr, m, n, dY.ld, None, dG.ptr))
    ctx.sync()
    ref_adj, ref_gram = tO.clone(), tG.clone()
    assert orc.rel_frobenius(dO.to_host(), orc.fastRUVIII_reduced(P["Y"], P["types"], P["ctl"], k)["newY"]) < TOL_NEWY
    bad = 0
    for it in range(400):
        tO.zero_()
        L.check(lib.scm_adjust(ctx.handle, dY.ptr, m, n, dY.ld, coef, None, 0, dO.ptr, dO.ld, None))  # out of place
        tW.copy_(tY)
        L.check(lib.scm_adjust(ctx.handle, dW.ptr, m, n, dW.ld, coef, None, 0, dW.ptr, dW.ld, None))  # in place
        if it % 8 == 0:
            tG.zero_()
            L.check(lib.scm_gram(ctx.handle, dY.ptr, m, n, dY.ld, None, dG.ptr))
            bad += int(not torch.equal(tG, ref_gram))
        ctx.sync()
        bad += int(not torch.equal(tO, ref_adj)) + int(not torch.equal(tW, ref_adj))
    lib.scm_coef_destroy(ctx.handle, coef)
    assert bad == 0


def test_w_output_and_device_resident(sm, ctx, sim):
    from scmerge_b200.ruv import _adjust
    ref = orc.fastRUVIII(sim["Y"], sim["M"], sim["ctl"], 20, return_info=True)
    alpha = ref["fullalpha"][:20]
    ac = alpha[:, sim["ctl"]]
    W_ref = sim["Y"][:, sim["ctl"]] @ ac.T @ np.linalg.inv(ac @ ac.T)
    dY = sm.DeviceMatrix.from_host(ctx, sim["Y"])
    dout, W = _adjust(ctx, dY, ref["fullalpha"], 20, sim["ctl"], want_W=True)
    # W is defined up to the basis of alpha: compare W alpha (basis-free) and the column spaces
    assert orc.rel_frobenius(dout.to_host(), ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(W.T, W_ref.T).max() < 1e-6
    res = sm.fastRUVIII(dY, sim["M"], sim["ctl"], k=20)  # device in -> device out
    assert isinstance(res, sm.DeviceMatrix)
    assert orc.rel_frobenius(res.to_host(), ref["newY"]) < TOL_NEWY


def test_scmerge2_core_cosine_pseudobulk_ruv(sm):
    """BASELINE.json configs[2] in miniature: cosine-normalised cells, pseudobulks over (batch, fold, cluster),
    RUV-III fit on the pseudobulks, every cell adjusted -- against the oracle's scMerge2 core."""
    from scmerge_b200.scmerge2 import cosineNorm, pseudobulk_keys, scMerge2_core
    P = orc.planted_factors(6000, 320, 320, 6, n_types=5, n_batch=4, seed=2024)
    Y = np.abs(P["Y"]) + 0.1  # expression-like, non-negative
    batch = np.array([f"batch{b}" for b in P["batch"]])
    clust = np.array([f"ct{t}" for t in P["types"]])
    keys, npb, names, pb_t, pb_b = pseudobulk_keys(batch, clust, 10, seed=3)
    # fold ids implied by the keys (what cvFolds' `which` would be): recover them from the names for the oracle
    fold_of_key = np.array([int(nm.rsplit("_", 1)[1]) for nm in names])
    which = fold_of_key[keys]
    ctx = sm.default_context()
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    cosineNorm(dY)
    assert np.allclose(dY.to_host(), orc.cosine_norm(Y.T).T, rtol=1e-14, atol=0)
    out = scMerge2_core(Y, batch, clust, ruvK=6, k_pseudoBulk=10, which=which, BSPARAM=sm.ExactParam())
    ref = orc.scmerge2_core(Y.T, batch, clust, which, ruvK=6, k_fold=10, exact=True)
    assert out["pseudobulk_names"] == ref["names"]                      # pseudobulk indices bit-exact
    assert np.allclose(out["bulkExprs"], ref["bulkExprs"], rtol=1e-12, atol=1e-15)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(out["fullalpha"][:6], ref["fullalpha"][:6]).max() < 1e-6
    # scMerge2's default (RandomParam): the reference rule asks for min(P - G, n) rows of fullalpha -- all of them come back
    # (whole spectrum: the P x P side when there are fewer pseudobulks than genes), the adjusted matrix is the exact one
    rnd = scMerge2_core(Y, batch, clust, ruvK=6, k_pseudoBulk=10, which=which)
    Pn, Gn = len(ref["names"]), len(set(clust))
    assert rnd["fullalpha"].shape == (min(Pn - Gn, 320), 320)
    assert orc.rel_frobenius(rnd["newY"], ref["newY"]) < TOL_NEWY
    # opt-out: max(k, 50) converged rows from the subspace iteration (what round 1 always did), with a warning
    import scmerge_b200.ruv as ruv_mod
    ruv_mod.FULLALPHA_ROWS = "truncate"
    try:
        with pytest.warns(UserWarning, match="truncated"):
            trn = scMerge2_core(Y, batch, clust, ruvK=6, k_pseudoBulk=10, which=which)
    finally:
        ruv_mod.FULLALPHA_ROWS = "reference"
    assert trn["fullalpha"].shape[0] == 50 and orc.rel_frobenius(trn["newY"], ref["newY"]) < TOL_NEWY
    # device-resident, in place
    dY2 = sm.DeviceMatrix.from_host(ctx, Y)
    res = scMerge2_core(dY2, batch, clust, ruvK=6, k_pseudoBulk=10, which=which, BSPARAM=sm.ExactParam())
    assert res["newY"] is dY2 and orc.rel_frobenius(dY2.to_host(), ref["newY"]) < TOL_NEWY


# ---- K8: sparse (dgCMatrix) ingest ---------------------------------------------------------------------
def _sparse_case(m, n, density, seed, empty_cells=()):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    X = sp.random(m, n, density=density, format="csr", random_state=rng, data_rvs=lambda k: rng.gamma(2.0, 1.0, k) + 0.05)
    X = X.tolil()
    for r in empty_cells:
        X.rows[r], X.data[r] = [], []
    return X.tocsr()


@pytest.mark.parametrize("m,n,density,cosine", [(700, 257, 0.1, True), (300, 5000, 0.02, True), (64, 33, 0.5, False),
                                                (2000, 4097, 0.05, True)])
def test_csc_to_dense_bit_exact(sm, ctx, m, n, density, cosine):
    """Densification is a scatter: bit-exact against the host; the fused cosine scale matches batchelor's formula
    (x / max(||x||, 1e-8)) to the last few ulps (the squared norm is summed in a different order)."""
    X = _sparse_case(m, n, density, seed=m + n, empty_cells=(0, m // 2))
    D = X.toarray()
    d = sm.csc_to_device(ctx, X, cosine=cosine)
    got = d.to_host()
    d.free()
    if not cosine:
        assert np.array_equal(got, D)
    else:
        ref = orc.cosine_norm(D.T).T
        assert np.array_equal(got != 0, D != 0)                      # sparsity pattern (gene indices) bit-exact
        assert np.allclose(got, ref, rtol=1e-14, atol=0)
        assert np.all(got[0] == 0) and np.all(got[m // 2] == 0)      # empty cells stay all-zero (norm floored, no NaN)
    # the CSC view of genes x cells is the same three arrays
    d2 = sm.csc_to_device(ctx, X.T.tocsc(), cosine=cosine)
    assert np.array_equal(d2.to_host(), got)
    d2.free()


def test_csc_ingest_rejects_corrupt_and_nonfinite(sm, ctx):
    from types import SimpleNamespace
    X = _sparse_case(200, 64, 0.2, seed=5)

    def raw(indptr=None, indices=None, data=None):  # the three dgCMatrix slots without scipy's own validation
        return SimpleNamespace(format="csr", shape=X.shape, indptr=X.indptr.copy() if indptr is None else indptr,
                               indices=X.indices.copy() if indices is None else indices,
                               data=X.data.copy() if data is None else data)

    idx = X.indices.copy()
    idx[3] = 64                                                     # gene index out of range
    with pytest.raises(sm.ScmError, match="gene index"):
        sm.csc_to_device(ctx, raw(indices=idx))
    ptr = X.indptr.copy()
    ptr[5] = ptr[6] + 1                                             # decreasing offsets
    with pytest.raises(sm.ScmError, match="column pointers"):
        sm.csc_to_device(ctx, raw(indptr=ptr))
    val = X.data.copy()
    val[7] = np.inf
    with pytest.raises(sm.NonFiniteError):
        sm.csc_to_device(ctx, raw(data=val))
    with pytest.raises(sm.NonFiniteError):                          # also without the norm pass
        sm.csc_to_device(ctx, raw(data=val), cosine=False)
    # un-canonical input (adjacent duplicates) is summed on the device
    p0 = int(X.indptr[10])
    assert X.indptr[11] - p0 >= 2
    idx = X.indices.copy()
    idx[p0 + 1] = idx[p0]
    want = np.zeros(X.shape)
    rows = np.repeat(np.arange(X.shape[0]), np.diff(X.indptr))
    np.add.at(want, (rows, idx), X.data)
    d = sm.csc_to_device(ctx, raw(indices=idx), cosine=False)
    assert np.allclose(d.to_host(), want, rtol=1e-15, atol=0)
    d.free()


def test_scmerge2_core_sparse_input_matches_dense(sm):
    """scMerge2 on the dgCMatrix it actually holds (R/scMerge2.R:164-166): same result as the dense path / the oracle,
    with the adjusted matrix streamed into a page-locked host buffer."""
    from scmerge_b200.scmerge2 import pseudobulk_keys, scMerge2_core
    m, n = 5000, 300
    rng = np.random.default_rng(99)
    batch = rng.integers(0, 3, m)
    clust = rng.integers(0, 4, m)
    X = _sparse_case(m, n, 0.15, seed=17).tolil()
    # cell-type and batch x cell-type signal on top of the sparse noise so that the fit has factors to find
    T = rng.gamma(2.0, 1.0, (4, n)) * (rng.random((4, n)) < 0.2)
    Bf = rng.gamma(2.0, 0.5, (3, 4, n)) * (rng.random((3, 4, n)) < 0.1)
    D = X.toarray() + T[clust] + Bf[batch, clust]
    import scipy.sparse as sp
    Xs = sp.csr_matrix(D)
    # 3 batches x 4 types -> 8 batch-by-type factors: ruvK = 8 sits at the spectral gap (sigma_8 / sigma_9 ~ 4)
    keys, npb, names, pb_t, pb_b = pseudobulk_keys(batch, clust, 8, seed=3)
    which = np.array([int(nm.rsplit("_", 1)[1]) for nm in names])[keys]
    ref = orc.scmerge2_core(D.T, batch, clust, which, ruvK=8, k_fold=8, exact=True)
    host_out = sm.pinned_empty((m, n))
    out = scMerge2_core(Xs, batch, clust, ruvK=8, k_pseudoBulk=8, which=which, BSPARAM=sm.ExactParam(), out=host_out)
    assert out["newY"] is host_out
    assert out["pseudobulk_names"] == ref["names"]
    assert np.allclose(out["bulkExprs"], ref["bulkExprs"], rtol=1e-12, atol=1e-15)
    assert orc.rel_frobenius(host_out, ref["newY"]) < TOL_NEWY
    dense = scMerge2_core(D, batch, clust, ruvK=8, k_pseudoBulk=8, which=which, BSPARAM=sm.ExactParam())
    assert orc.rel_frobenius(dense["newY"], host_out) < 1e-10
    # scMerge2's default BSPARAM (RandomParam): every row of fullalpha the reference's rule asks for (here P - G = 92 > 88: the
    # whole-spectrum solve).  With a host result that solve runs on a second context beside the download, the cells being
    # adjusted with the top rows of a subspace iteration: same cells, same rows as the sequential order (SCM_SCM2_OVERLAP=0)
    import os
    r_over = scMerge2_core(Xs, batch, clust, ruvK=8, k_pseudoBulk=8, which=which, out=host_out)
    y_over = np.array(host_out)
    os.environ["SCM_SCM2_OVERLAP"] = "0"
    try:
        r_seq = scMerge2_core(Xs, batch, clust, ruvK=8, k_pseudoBulk=8, which=which, out=host_out)
    finally:
        del os.environ["SCM_SCM2_OVERLAP"]
    assert r_over["fullalpha"].shape == r_seq["fullalpha"].shape and r_over["fullalpha"].shape[0] > 88
    assert orc.rel_frobenius(y_over, host_out) < 1e-9 and orc.rel_frobenius(y_over, ref["newY"]) < TOL_NEWY
    assert np.array_equal(np.asarray(r_over["fullalpha"]), np.asarray(r_seq["fullalpha"]))     # the same solver on the same matrix
    assert orc.principal_angles(r_over["fullalpha"][:8], ref["fullalpha"][:8]).max() < 1e-6
    scMerge2_core(Xs, batch, clust, ruvK=8, k_pseudoBulk=8, which=which, BSPARAM=sm.ExactParam(), out=host_out)   # back to the exact run
    # the densify-first path (what un-canonical input takes) agrees with the sparse-native kernels
    dens = scMerge2_core(Xs, batch, clust, ruvK=8, k_pseudoBulk=8, which=which, BSPARAM=sm.ExactParam(), sparse_native=False)
    assert orc.rel_frobenius(dens["newY"], host_out) < 1e-10
    assert np.allclose(dens["bulkExprs"], out["bulkExprs"], rtol=1e-13, atol=1e-16)
    # genes x cells CSC (the reference's orientation) gives the same answer, bit for bit (deterministic kernels)
    out2 = scMerge2_core(Xs.T.tocsc(), batch, clust, ruvK=8, k_pseudoBulk=8, which=which, BSPARAM=sm.ExactParam())
    assert np.array_equal(out2["newY"], host_out)
    # device-resident sparse input -> device-resident dense output
    ctx = sm.default_context()
    dcsc = sm.DeviceCSC.from_scipy(ctx, Xs)
    res = scMerge2_core(dcsc, batch, clust, ruvK=8, k_pseudoBulk=8, which=which, BSPARAM=sm.ExactParam())
    assert isinstance(res["newY"], sm.DeviceMatrix) and np.array_equal(res["newY"].to_host(), host_out)
    dcsc.free()


@pytest.mark.parametrize("m,n,density,k,G", [(3000, 257, 0.1, 5, 40), (5000, 2000, 0.05, 20, 700), (700, 4500, 0.02, 32, 3),
                                             (64, 96, 0.5, 8, 64)])
def test_sparse_native_kernels_vs_dense(sm, ctx, m, n, density, k, G):
    """scm_csc_rownorm / scm_csc_seg_colmean / scm_csc_adjust against the dense kernels on the densified matrix: norms and
    group counts bit-exact, means and the adjusted matrix to rounding (different summation order), empty cells / groups,
    ragged tiles, odd n and a control subset included."""
    from scmerge_b200.ruv import _Coef
    X = _sparse_case(m, n, density, seed=m + k, empty_cells=(1, m - 1))
    rng = np.random.default_rng(k)
    keys = rng.integers(0, G, m).astype(np.int32)
    keys[keys == 1] = 0                                             # group 1 is empty
    dcsc = sm.DeviceCSC.from_scipy(ctx, X)
    assert dcsc.is_canonical()
    d_inv = dcsc.row_scales(cosine=True)
    D = X.toarray()
    nrm = np.sqrt((D * D).sum(axis=1))
    assert np.allclose(d_inv.to_host(), 1.0 / np.maximum(nrm, 1e-8), rtol=1e-14)
    Dn = orc.cosine_norm(D.T).T
    lib = sm.lib()
    dP = sm.DeviceMatrix(ctx, G, n)
    d_keys, d_cnt = ctx.to_device(keys), ctx.empty((G,), np.int64)
    assert lib.scm_csc_seg_colmean(ctx.handle, dcsc.indptr.ptr, dcsc.indices.ptr, dcsc.data.ptr, d_inv.ptr, m, n, d_keys.ptr, G,
                                   dP.ptr, dP.ld, d_cnt.ptr) == 0, lib.scm_last_error()
    cnt = np.bincount(keys, minlength=G)
    assert np.array_equal(d_cnt.to_host(), cnt)
    want = np.zeros((G, n))
    np.add.at(want, keys, Dn)
    want[cnt > 0] /= cnt[cnt > 0, None]
    got = dP.to_host()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-16) and np.all(got[1] == 0)
    again = sm.DeviceMatrix(ctx, G, n)
    assert lib.scm_csc_seg_colmean(ctx.handle, dcsc.indptr.ptr, dcsc.indices.ptr, dcsc.data.ptr, d_inv.ptr, m, n, d_keys.ptr, G,
                                   again.ptr, again.ld, d_cnt.ptr) == 0
    assert np.array_equal(again.to_host(), got)                     # deterministic
    # rank-k update: random factor rows, controls = every third gene, centre = column means
    alpha = rng.normal(size=(k, n))
    ctl = np.zeros(n, bool)
    ctl[::3] = True
    center = Dn.mean(axis=0)
    ac = alpha[:, ctl]
    W = (Dn - center)[:, ctl] @ ac.T @ np.linalg.inv(ac @ ac.T)
    ref = Dn - W @ alpha
    dout = sm.DeviceMatrix(ctx, m, n)
    with _Coef(ctx, n, alpha, k, ctl, center=center) as coef:
        assert lib.scm_csc_adjust(ctx.handle, dcsc.indptr.ptr, dcsc.indices.ptr, dcsc.data.ptr, d_inv.ptr, m, n, coef.handle,
                                  dout.ptr, dout.ld) == 0, lib.scm_last_error()
        ctx.sync()
        assert orc.rel_frobenius(dout.to_host(), ref) < 1e-11
        host = sm.pinned_empty((m, n))
        os.environ["SCM_HOST_CHUNK_ROWS"] = "100"                   # several staging chunks
        try:
            assert lib.scm_csc_adjust_to_host(ctx.handle, dcsc.indptr.ptr, dcsc.indices.ptr, dcsc.data.ptr, d_inv.ptr, m, n,
                                              coef.handle, host.ctypes.data_as(C.c_void_p), n) == 0, lib.scm_last_error()
        finally:
            del os.environ["SCM_HOST_CHUNK_ROWS"]
        assert np.array_equal(host, dout.to_host())
    for a in (d_inv, d_keys, d_cnt):
        a.free()
    dcsc.free()


def test_scruvg_eta_and_fullw_reuse_vs_oracle(sm):
    """scRUVg's remaining arguments (R/scRUVg.R:59,64,77-80): eta (the decomposition sees RUV1(Y), the subtraction the original
    Y) and a re-used fullW (alpha by least squares on the given W: scm_ls_coef + scm_rankk_subtract)."""
    L = orc.ruv_simulate(m=400, n=700, nc=90, n_celltypes=4, rng=np.random.default_rng(321))
    Y, ctl, k = L["Y"], L["ctl"], 7
    eta = np.random.default_rng(2).normal(size=(Y.shape[1], 2))
    ref = orc.scRUVg(Y, ctl, k, eta=eta)
    got = sm.scRUVg(Y, ctl, k, eta=eta)
    assert orc.rel_frobenius(got["newY"], ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(got["W"].T, ref["W"].T).max() < 1e-6
    assert orc.principal_angles(got["alpha"], ref["alpha"]).max() < 1e-6
    base = orc.scRUVg(Y, ctl, 12)                                      # a previous decomposition with more columns
    for e in (None, eta):
        ref2 = orc.scRUVg(Y, ctl, k, eta=e, fullW=base["W"])
        got2 = sm.scRUVg(Y, ctl, k, eta=e, fullW=base["W"])
        assert orc.rel_frobenius(got2["newY"], ref2["newY"]) < TOL_NEWY
        assert np.allclose(got2["W"], base["W"][:, :k]) and np.allclose(got2["alpha"], ref2["alpha"], rtol=1e-8, atol=1e-10)
    with pytest.raises(ValueError):
        sm.scRUVg(Y, ctl, k, svdyc=object())                           # svdyc alone fails in the reference too (:77)
    with pytest.raises(ValueError):
        sm.scRUVg(Y, ctl, k, fullW=base["W"][:, :3])


def test_group_means_with_thousands_of_groups(sm, ctx):
    """scm_seg_colmean / scm_group_colmean with more groups than any launch dimension they use (2-D grid clamped at 2048 group
    rows, 32 thread rows per gene): the 4800 pseudobulks of scMerge2 and beyond; empty groups stay 0; colMeans of ALL cells
    recovered from the group means and counts."""
    rng = np.random.default_rng(11)
    m, n, G = 30000, 130, 5000
    Y = rng.normal(size=(m, n)) + 3.0
    keys = rng.integers(0, G, m).astype(np.int32)
    keys[keys == 7] = 8                                                  # group 7 is empty
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    d_keys, d_cnt = ctx.to_device(keys), ctx.empty((G,), np.int64)
    dP = sm.DeviceMatrix(ctx, G, n)
    lib = sm.lib()
    assert lib.scm_seg_colmean(ctx.handle, dY.ptr, m, n, dY.ld, d_keys.ptr, G, dP.ptr, d_cnt.ptr, None) == 0, lib.scm_last_error()
    cnt = np.bincount(keys, minlength=G)
    assert np.array_equal(d_cnt.to_host(), cnt)
    want = np.zeros((G, n))
    np.add.at(want, keys, Y)
    want[cnt > 0] /= cnt[cnt > 0, None]
    got = dP.to_host()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-14) and np.all(got[7] == 0)
    d_mean = ctx.empty((n,))
    assert lib.scm_group_colmean(ctx.handle, dP.ptr, d_cnt.ptr, G, n, dP.ld, d_mean.ptr) == 0
    assert np.allclose(d_mean.to_host(), Y.mean(axis=0), rtol=1e-12)
    for a in (d_keys, d_cnt, d_mean):
        a.free()
    dP.free()
    dY.free()


# ---- eta: gene-wise covariates (ruv::RUV1, R/fastRUVIII.R:63-64, R/scMerge2_utilis.R:56) ---------------------------------
@pytest.mark.parametrize("intercept", [True, False])
def test_fastruviii_and_pseudoruviii_with_eta_vs_oracle(sm, ctx, intercept):
    """Y <- RUV1(Y, eta, ctl) runs through the rank-k update kernels with the rows of eta in the place of alpha; host matrix,
    device-resident matrix (the caller's matrix must stay untouched), the early-out and the pseudobulk fit."""
    L = orc.ruv_simulate(m=300, n=600, nc=150, n_celltypes=4, rng=np.random.default_rng(77))
    Y, M, ctl = L["Y"], L["M"], L["ctl"]
    n = Y.shape[1]
    eta = np.random.default_rng(5).normal(size=(n, 2))
    ref = orc.fastRUVIII(Y, M, ctl, 6, eta=eta, include_intercept=intercept, return_info=True)
    got = sm.fastRUVIII(Y, M, ctl, k=6, eta=eta, include_intercept=intercept, return_info=True)
    assert orc.rel_frobenius(got["newY"], ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(got["fullalpha"][:6], ref["fullalpha"][:6]).max() < 1e-6
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    dres = sm.fastRUVIII(dY, M, ctl, k=6, eta=eta.T, include_intercept=intercept)       # q x n orientation, resident input
    assert isinstance(dres, sm.DeviceMatrix) and np.array_equal(dres.to_host(), got["newY"])
    assert np.array_equal(dY.to_host(), Y)                                               # the caller's matrix is not overwritten
    dres.free()
    # early-out (k = 0): the reference returns RUV1(Y), not Y
    r0 = sm.fastRUVIII(Y, M, ctl, k=0, eta=eta, include_intercept=intercept)
    assert orc.rel_frobenius(r0, orc.RUV1(Y, eta, ctl, intercept)) < 1e-12
    # pseudoRUVIII: RUV1 applies to the pseudobulk matrix only
    grp = np.arange(Y.shape[0]) % 60
    Yp = np.stack([Y[grp == g].mean(axis=0) for g in range(60)])
    Mp = np.arange(60) % 12
    pref = orc.pseudoRUVIII(Y, Yp, Mp, ctl, 5, eta=eta, include_intercept=intercept)
    pgot = sm.pseudoRUVIII(Y, Yp, Mp, ctl, k=5, eta=eta, include_intercept=intercept, BSPARAM=sm.ExactParam(), return_info=True)
    assert orc.rel_frobenius(pgot["newY"], pref["newY"]) < TOL_NEWY
    dY.free()


# ---- scRUVg ---------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,nc,k", [(80, 1000, 50, 20), (3000, 400, 120, 10)])
def test_scruvg_vs_oracle(sm, m, n, nc, k):
    """R/scRUVg.R:30-85 on the device (Gram of the centred cells -> eigenpairs of its control block -> fused rank-k
    update) against the literal oracle; the reference's own assertion is abs(W) == abs(RUV2 W) (test-scRUVg.R:7)."""
    L = orc.ruv_simulate(m=m, n=n, nc=nc, n_celltypes=5, rng=np.random.default_rng(1234 + m))
    ref = orc.scRUVg(L["Y"], L["ctl"], k)
    got = sm.scRUVg(L["Y"], L["ctl"], k)
    assert orc.rel_frobenius(got["newY"], ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(got["W"].T, ref["W"].T).max() < 1e-6
    assert orc.principal_angles(got["alpha"], ref["alpha"]).max() < 1e-6
    assert np.allclose(got["W"].T @ got["W"], np.eye(k), atol=1e-9)            # W orthonormal like svd's u
    sgn = np.sign(np.sum(got["W"] * ref["W"], axis=0))
    # individual columns match up to sign where the singular values are separated
    s = np.linalg.svd((L["Y"] - L["Y"].mean(0))[:, orc.tological(L["ctl"], n)], compute_uv=False)
    sep = np.minimum(np.abs(np.diff(s[: k + 1])), np.r_[np.inf, np.abs(np.diff(s[:k]))]) / s[0] > 1e-3
    assert np.allclose((got["W"] * sgn)[:, sep], ref["W"][:, sep], atol=1e-6)
    with pytest.raises(ValueError, match="number of controls"):
        sm.scRUVg(L["Y"], L["ctl"], nc + 1)
    # device-resident in -> device-resident out
    dY = sm.DeviceMatrix.from_host(sm.default_context(), L["Y"])
    res = sm.scRUVg(dY, L["ctl"], k)
    assert isinstance(res["newY"], sm.DeviceMatrix) and orc.rel_frobenius(res["newY"].to_host(), ref["newY"]) < TOL_NEWY


def test_golden_select_and_ruvg_on_example_sce(sm, golden_dir):
    """BASELINE.json configs[0] (the reference's bundled example_sce) through scMerge(ruvK = c(5, 10, 20))'s arithmetic:
    the device selection (PC scores from the centred Gram + exact silhouette) against the committed oracle fixture, and
    scRUVg on the same data."""
    d = np.load(os.path.join(golden_dir, "example_sce.npz"))
    g = np.load(os.path.join(golden_dir, "golden_select.npz"))
    ks = [int(k) for k in g["ks"]]
    res = sm.scRUVIII(d["logcounts"], d["cellTypes"], d["ctl"], k=ks, batch=d["batch"], cell_type=d["cellTypes"])
    assert np.allclose(res["sil"], g["sil"], rtol=0, atol=1e-7)
    assert np.allclose(res["f_score"], g["f_score"], rtol=0, atol=1e-7)
    assert res["optimal_ruvK"] == int(g["optimal"])
    assert abs(np.linalg.norm(res[5]["newY"]) - float(g["newY5_fro"])) / float(g["newY5_fro"]) < 1e-9
    assert np.allclose(res[5]["newY"].sum(axis=0), g["newY5_colsum"], rtol=1e-8, atol=1e-8)
    r = sm.scRUVg(d["logcounts"].T, d["ctl"], 10)
    assert orc.rel_frobenius(r["newY"][:16], g["ruvg_newY_rows"]) < TOL_NEWY
    assert abs(np.linalg.norm(r["newY"]) - float(g["ruvg_newY_fro"])) / float(g["ruvg_newY_fro"]) < 1e-9
    assert np.allclose(np.linalg.norm(r["alpha"], axis=1), g["ruvg_alpha_rownorm"], rtol=1e-7)


def test_scmerge2_core_is_deterministic(sm):
    """The reference pins scMerge2 on seed-determinism (tests/testthat/test-scMerge2.R:68-77,131-136): same seed, same
    result.  Here the pseudobulk folds come from a seeded generator and every kernel is order-deterministic, so two runs are
    bit-identical, on the dense and on the sparse path."""
    import scipy.sparse as sp
    from scmerge_b200.scmerge2 import scMerge2_core
    rng = np.random.default_rng(7)
    m, n = 3000, 200
    batch, clust = rng.integers(0, 3, m), rng.integers(0, 4, m)
    D = rng.gamma(2.0, 1.0, (m, n)) * (rng.random((m, n)) < 0.3) + rng.gamma(2.0, 1.0, (4, n))[clust] * (rng.random((4, n)) < 0.3)[clust]
    D += (rng.gamma(2.0, 0.5, (3, 4, n)) * (rng.random((3, 4, n)) < 0.2))[batch, clust]
    runs = [scMerge2_core(D, batch, clust, ruvK=6, k_pseudoBulk=6, seed=11, BSPARAM=sm.ExactParam()) for _ in range(2)]
    assert np.array_equal(runs[0]["newY"], runs[1]["newY"]) and runs[0]["pseudobulk_names"] == runs[1]["pseudobulk_names"]
    Xs = sp.csr_matrix(D)
    sruns = [scMerge2_core(Xs, batch, clust, ruvK=6, k_pseudoBulk=6, seed=11, BSPARAM=sm.ExactParam()) for _ in range(2)]
    assert np.array_equal(sruns[0]["newY"], sruns[1]["newY"])
    assert orc.rel_frobenius(sruns[0]["newY"], runs[0]["newY"]) < 1e-9
    other = scMerge2_core(D, batch, clust, ruvK=6, k_pseudoBulk=6, seed=12, BSPARAM=sm.ExactParam())
    assert not np.array_equal(other["newY"], runs[0]["newY"])        # a different seed draws different folds


def test_scruviii_reference_test_shape(sm):
    """The shape of the reference's own scRUVIII test (tests/testthat/test-scRUVIII.R:3-7): ruvSimulate(m = 200, n = 1000,
    nc = 100, nCelltypes = 3, nBatch = 2, lambda = 0.1), Y = t(log2(Y + 1)), k = c(5, 10, 15, 20), M a one-hot matrix and no
    cell_type (the replicate structure stands in, R/scRUVIII.R:89-92).  The reference only checks that it runs; here every
    output is compared with the oracle."""
    L_ = orc.ruv_simulate(m=200, n=1000, nc=100, n_celltypes=3, n_batch=2, lam=0.1, rng=np.random.default_rng(1234))
    Y = np.log2(L_["Y"] + 1).T
    keep = Y.var(axis=1) > 0                       # standardize2 has no zero-variance guard (R/scRUVIII.R:171)
    Y, ctl = Y[keep], L_["ctl"][keep]
    ks = [5, 10, 15, 20]
    res = sm.scRUVIII(Y, L_["M"], ctl, k=ks, batch=L_["batch"])
    ref = orc.scRUVIII(Y, L_["M"], ctl, ks, L_["batch"])
    for kk in ks:
        assert orc.rel_frobenius(res[kk]["newY"], ref[kk]["newY"]) < TOL_NEWY
    assert np.allclose(res["sil"], ref["sil"], rtol=0, atol=1e-7)
    assert np.allclose(res["f_score"], ref["f_score"], rtol=0, atol=1e-7)
    assert res["optimal_ruvK"] == ref["optimal_ruvK"]
    assert set(res.keys()) >= set(ks) | {"optimal_ruvK"}       # list named by k + $optimal_ruvK (R/scRUVIII.R:77,126-136)


# ---- round 2: one summing pass writes the shifted copy, device-side eigensolver control, converged rank -----------------
def test_k1_fused_shifted_copy_is_bit_identical(sm, monkeypatch):
    """The pre-shifted copy of the cells written from K1's summing pass (SCM_K1_FUSED_COPY=1) must give the same bits as the
    separate copy kernel (SCM_K1_FUSED_COPY=0): same shift vector, same subtraction, only the kernel that stores it differs.
    Odd gene count (scalar tail column, padded leading dimension)."""
    P = orc.planted_factors(5000, 257, 96, 6, n_types=4, n_batch=3, seed=7)
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("SCM_K1_FUSED_COPY", flag)
        ctx = sm.Context(0)
        dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
        res = sm.fastRUVIII(dY, P["types"], P["ctl"], k=6, return_info=True, ctx=ctx)
        outs.append((res["newY"].to_host(), np.asarray(res["fullalpha"])))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 6)
    assert orc.rel_frobenius(outs[0][0], ref) < TOL_NEWY


def test_gram_shift_paths_agree(sm, monkeypatch):
    """The centred Gram does not depend on HOW the local shift is applied: pre-shifted copy + SHIFT=false kernel (default),
    shift subtracted at fragment load (SCM_GRAM_PRESHIFT=0) and the cp.async kernel (SCM_GRAM_TMA=0) must all reproduce the
    oracle; means far from zero (offset 50) make a wrong re-centring visible."""
    P = orc.planted_factors(6000, 320, 100, 8, n_types=5, n_batch=3, seed=17)
    Y = P["Y"] + 50.0
    ref = orc.fastRUVIII(Y, P["types"], P["ctl"], 8, return_info=True)
    for env in ({}, {"SCM_GRAM_PRESHIFT": "0"}, {"SCM_GRAM_TMA": "0"}):
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        ctx = sm.Context(0)
        out = sm.fastRUVIII(Y, P["types"], P["ctl"], k=8, return_info=True, ctx=ctx)
        for k_ in env:
            monkeypatch.delenv(k_)
        assert orc.rel_frobenius(out["newY"], ref["newY"]) < TOL_NEWY, env
        assert orc.principal_angles(out["fullalpha"][:8], ref["fullalpha"][:8]).max() < 1e-6, env


def test_every_cell_needs_a_replicate_key(sm, ctx):
    """M assigns every cell to a group (ruv::replicate.matrix): a key outside [0, G) in a FIT is rejected (the Gram is over all
    cells, the group sums would silently miss the row)."""
    P = orc.planted_factors(800, 64, 32, 3, n_types=3, n_batch=2, seed=3)
    from scmerge_b200.ruv import _fit
    dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
    keys = P["types"].astype(np.int32).copy()
    keys[5] = -1
    with pytest.raises(sm.ScmError, match="replicate key"):
        _fit(ctx, dY, keys, 3, 10, 3, True)
    dY.free()


def test_converged_rank_is_reported_and_guards_reuse(sm, ctx):
    """fullalpha is the reference's checkpoint (R/fastRUVIII.R:83, R/scMerge2.R:399).  The device fit converges the ranks it
    was asked for; the returned array carries `conv_rank`, re-use within it matches a direct fit, re-use beyond it warns."""
    import warnings
    P = orc.planted_factors(4000, 200, 80, 12, n_types=4, n_batch=3, seed=23)
    dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
    r12 = sm.fastRUVIII(dY, P["types"], P["ctl"], k=12, return_info=True, ctx=ctx)
    fa = r12["fullalpha"]
    assert isinstance(fa, sm.FullAlpha) and fa.conv_rank is not None and fa.conv_rank >= 12
    ref5 = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 5)
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # no warning: 5 <= conv_rank
        got5 = sm.fastRUVIII(dY, P["types"], P["ctl"], k=5, fullalpha=fa, ctx=ctx).to_host()
    assert orc.rel_frobenius(got5, ref5) < TOL_NEWY
    r3 = sm.fastRUVIII(dY, P["types"], P["ctl"], k=3, return_info=True, ctx=ctx)
    fa3 = r3["fullalpha"]
    if fa3.conv_rank < 40:
        with pytest.warns(sm.NotConvergedWarning, match="exceeds the rank"):
            sm.fastRUVIII(dY, P["types"], P["ctl"], k=40, fullalpha=fa3, ctx=ctx)
    info = ctx.fit_info()
    assert info["iters"] >= 1 and info["conv_rank"] == fa3.conv_rank
    dY.free()


def test_chol_inv_and_eig_small_blocks(sm, ctx):
    """The single-CTA Cholesky / Jacobi kernels behind K4 and K5 at block sizes from 1 to the limit: eigenvalues against LAPACK
    (relative 1e-12), eigenvectors as invariant subspaces."""
    rng = np.random.default_rng(5)
    lib = sm.lib()
    for n, s in [(40, 1), (64, 7), (150, 33), (300, 64), (513, 96)]:
        A = rng.normal(size=(n + 20, n)) * np.linspace(1, 6, n)[None, :] ** 2
        Gm = A.T @ A
        d_G = ctx.to_device(Gm)
        d_vals, d_vecs = ctx.empty((s,)), ctx.empty((s, n))
        it, rs = C.c_int32(0), C.c_double(0.0)
        rc = lib.scm_eig_topk(ctx.handle, d_G.ptr, n, s, s, 0, 1e-10, 500, 1, d_vals.ptr, d_vecs.ptr, C.byref(it), C.byref(rs))
        assert rc in (0, -6), lib.scm_last_error()
        w, V = np.linalg.eigh(Gm)
        w, V = w[::-1], V[:, ::-1]
        vals, vecs = d_vals.to_host(), d_vecs.to_host()
        assert np.allclose(vals, w[:s], rtol=1e-9), (n, s)
        if rc == 0:
            assert orc.principal_angles(vecs, V[:, :s].T).max() < 1e-6, (n, s)
        for a in (d_G, d_vals, d_vecs):
            a.free()


def test_single_process_multi_context_same_device(sm):
    """scm_multi_create with ONE device: the single-process driver (what an R session holds) must give exactly the per-context
    result; with more devices it is covered by tests/test_gpu_multi.py."""
    P = orc.planted_factors(3000, 128, 64, 5, n_types=4, n_batch=2, seed=2)
    mg = sm.MultiContext(devices=[0])
    from scmerge_b200.ruv import multi_fastRUVIII
    out = multi_fastRUVIII(mg, P["Y"], P["types"], P["ctl"], k=5, return_info=True)
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 5, return_info=True)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < TOL_NEWY
    assert orc.principal_angles(out["fullalpha"][:5], ref["fullalpha"][:5]).max() < 1e-6


def test_pageable_and_pinned_host_matrices_give_identical_bits(sm, monkeypatch):
    """scm_ruv3_host stages ordinary (pageable) host memory -- what an R session's REAL(Y) is -- through page-locked bounce
    buffers with helper threads; page-locked callers take the direct copies.  Same chunks, same kernels: the results must agree
    bit for bit.  SCM_HOST_CHUNK_ROWS forces 11 chunks (ragged last one), i.e. several trips around the 4-buffer ring."""
    monkeypatch.setenv("SCM_HOST_CHUNK_ROWS", "1000")
    P = orc.planted_factors(10_500, 130, 60, 5, n_types=4, n_batch=3, seed=12)
    ctx = sm.Context(0)
    Yp = sm.pinned_empty((10_500, 130))
    Yp[:] = P["Y"]
    outp = sm.pinned_empty((10_500, 130))
    r_pin = sm.fastRUVIII(Yp, P["types"], P["ctl"], k=5, return_info=True, ctx=ctx, out=outp)
    r_pag = sm.fastRUVIII(np.array(P["Y"]), P["types"], P["ctl"], k=5, return_info=True, ctx=ctx)
    np.testing.assert_array_equal(np.asarray(r_pin["newY"]), r_pag["newY"])
    np.testing.assert_array_equal(np.asarray(r_pin["fullalpha"]), np.asarray(r_pag["fullalpha"]))
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 5)
    assert orc.rel_frobenius(r_pag["newY"], ref) < TOL_NEWY
    # scRUVIII (batch statistics) and the adjust-only entry (pseudoRUVIII with column means) through the staged path
    batch = np.array([f"b{b}" for b in P["batch"]])
    s = sm.scRUVIII(np.asfortranarray(P["Y"].T), P["types"], P["ctl"], k=5, batch=batch, ctx=ctx)
    refs = orc.scRUVIII(P["Y"].T, P["types"], P["ctl"], 5, batch)
    assert orc.rel_frobenius(s[5]["newY"], refs[5]["newY"]) < TOL_NEWY


@pytest.mark.parametrize("pinned", [False, True])
def test_out_of_core_host_path_matches_in_core(sm, monkeypatch, pinned):
    """A host matrix that does not fit the device is streamed twice through three chunk buffers (scm_ruv3_host, out-of-core form:
    the chunked adjust of R/scMerge2_utilis.R:104-124 for a matrix that lives in host memory).  Forced here with
    SCM_HOST_STREAM=1 and 13 ragged chunks.  fastRUVIII runs the same kernels on the same chunks in both forms: bit-identical.
    scRUVIII sums its squared deviations about the shift and re-centres them (no second pass over the cells): oracle parity.
    Pageable and page-locked buffers (staged / direct copies), in-place host result, adjust-only entry with column means."""
    monkeypatch.setenv("SCM_HOST_CHUNK_ROWS", "800")
    P = orc.planted_factors(10_100, 120, 50, 5, n_types=4, n_batch=3, seed=21)
    ctx = sm.Context(0)

    def buf(a):
        if not pinned:
            return np.array(a)
        b = sm.pinned_empty(a.shape)
        b[:] = a
        return b
    Y = buf(P["Y"])
    monkeypatch.setenv("SCM_HOST_STREAM", "0")
    r_in = sm.fastRUVIII(Y, P["types"], P["ctl"], k=5, return_info=True, ctx=ctx, out=buf(np.zeros_like(P["Y"])))
    monkeypatch.setenv("SCM_HOST_STREAM", "1")
    r_oc = sm.fastRUVIII(Y, P["types"], P["ctl"], k=5, return_info=True, ctx=ctx, out=buf(np.zeros_like(P["Y"])))
    np.testing.assert_array_equal(np.asarray(r_in["newY"]), np.asarray(r_oc["newY"]))
    np.testing.assert_array_equal(np.asarray(r_in["fullalpha"]), np.asarray(r_oc["fullalpha"]))
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 5)
    assert orc.rel_frobenius(np.asarray(r_oc["newY"]), ref) < TOL_NEWY
    # in place on the host: the result overwrites the input buffer
    Yio = buf(P["Y"])
    sm.fastRUVIII(Yio, P["types"], P["ctl"], k=5, ctx=ctx, out=Yio)
    np.testing.assert_array_equal(np.asarray(Yio), np.asarray(r_oc["newY"]))
    # scRUVIII (genes x cells, column-major == cells x genes row-major) with batch statistics taken in the single pass
    batch = np.array([f"b{b}" for b in P["batch"]])
    Yt = np.asfortranarray(P["Y"].T) if not pinned else buf(P["Y"]).T
    s_oc = sm.scRUVIII(Yt, P["types"], P["ctl"], k=5, batch=batch, ctx=ctx)
    refs = orc.scRUVIII(P["Y"].T, P["types"], P["ctl"], 5, batch)
    assert orc.rel_frobenius(s_oc[5]["newY"], refs[5]["newY"]) < TOL_NEWY
    st = orc.standardize2(P["Y"].T, batch)
    assert np.allclose(s_oc["stand_mean"], st["stand_mean"], rtol=1e-12, atol=1e-13)
    assert np.allclose(s_oc["stand_sd"], np.sqrt(st["stand_var"]), rtol=1e-10)
    # adjust-only with column means accumulated during pass 1 (pseudoRUVIII / getAdjustedMat) and without any sums (fullalpha re-use)
    fa = np.asarray(r_oc["fullalpha"])
    a1 = sm.pseudoRUVIII(Y, P["Y"][:40], P["types"][:40], P["ctl"], k=5, fullalpha=fa, ctx=ctx)
    ref1 = orc.pseudoRUVIII(P["Y"], P["Y"][:40], P["types"][:40], orc.tological(P["ctl"], 120), 5, fullalpha=fa)
    assert orc.rel_frobenius(a1, ref1["newY"]) < TOL_NEWY
    a0 = sm.fastRUVIII(Y, P["types"], P["ctl"], k=5, fullalpha=fa, ctx=ctx)
    np.testing.assert_array_equal(np.asarray(a0), np.asarray(r_oc["newY"]))


# ---- whole-spectrum solver (eig_full.cuh): fits that return more rows of fullalpha than a subspace block carries ----------
@pytest.mark.parametrize("n,s", [(200, 150), (600, 300), (1100, 512), (2100, 120)])
def test_eig_full_spectrum_vs_lapack(sm, ctx, n, s):
    """One-sided block Jacobi on the Gram, every kernel geometry (n <= 512 / 1024 / 2048 / 4096): eigenvalues against LAPACK to
    1e-10 relative (of the largest), eigenvectors through the residual ||G v - lambda v|| and orthonormality -- bulk eigenvalues
    are nearly degenerate, so individual vectors are not comparable, invariant residuals are."""
    lib = sm.lib()
    rng = np.random.default_rng(n)
    A = rng.normal(size=(3 * n, n)) * np.linspace(1, 4, n)[None, :]
    A[:, :7] += rng.normal(size=(3 * n, 1)) * 6.0  # a few dominant directions
    Gm = A.T @ A
    d_G = ctx.to_device(Gm)
    d_vals, d_vecs = ctx.empty((s,)), ctx.empty((s, n))
    it, rs = C.c_int32(0), C.c_double(0.0)
    L_ = sm._lib
    L_.check(lib.scm_eig_topk(ctx.handle, d_G.ptr, n, s, s, 0, 1e-10, 500, 1, d_vals.ptr, d_vecs.ptr, C.byref(it), C.byref(rs)))
    w = np.linalg.eigvalsh(Gm)[::-1]
    vals, vecs = d_vals.to_host(), d_vecs.to_host()
    assert np.all(np.diff(vals) <= 0)
    assert np.abs(vals - w[:s]).max() <= 1e-10 * w[0], (n, s, it.value)
    assert np.abs(vecs @ vecs.T - np.eye(s)).max() < 1e-10
    assert np.abs(vecs @ Gm - vals[:, None] * vecs).max() <= 1e-9 * w[0]
    assert ctx.fit_info()["conv_rank"] == s
    for a in (d_G, d_vals, d_vecs):
        a.free()


def test_randomized_rule_returns_all_rows(sm):
    """R/fastRUVIII.R:66-70: with a non-exact BSPARAM the reference asks for min(m - ncol(M), sum(ctl)) singular vectors and
    `fullalpha` has that many rows.  The device returns all of them from the whole spectrum of the Gram; checked against the
    oracle's exact SVD with the same row count through the rotation-invariant fullalpha' fullalpha (bulk singular values are
    close, individual rows are not comparable), the singular values and the adjusted matrix.  A later call that re-uses the
    array with another k must equal a direct fit."""
    P = orc.planted_factors(3000, 400, 300, 10, n_types=6, n_batch=3, seed=77)
    out = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10, BSPARAM="random", return_info=True)
    fa = np.asarray(out["fullalpha"])
    assert fa.shape == (300, 400) and out["fullalpha"].conv_rank == 300
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 10, svd_k=300, exact=True, return_info=True)
    rfa = ref["fullalpha"]
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < TOL_NEWY
    assert np.allclose(np.linalg.norm(fa, axis=1), np.linalg.norm(rfa, axis=1), rtol=1e-9)
    assert orc.rel_frobenius(fa.T @ fa, rfa.T @ rfa) < 1e-9
    again = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=40, fullalpha=out["fullalpha"])
    direct = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 40, svd_k=300, exact=True)
    assert orc.rel_frobenius(again, direct) < TOL_NEWY


@pytest.mark.parametrize("n", [320, 4500])
def test_pseudobulk_fit_returns_the_full_row_space(sm, ctx, n):
    """scMerge2's default (RandomParam, ctl = all genes): svd_k = P - ncol(M), the whole row space of the residual pseudobulk
    matrix, whose Gram (n x n, n > P) is rank deficient -- the null space must be left alone and exactly P - G rows returned."""
    rng = np.random.default_rng(3)
    Pn, n, G = 150, 320, 6
    types = np.repeat(np.arange(G), Pn // G)
    B = rng.normal(size=(G, n)) * 2
    Yp = B[types] + rng.normal(size=(Pn, 4)) @ rng.normal(size=(4, n)) + 0.3 * rng.normal(size=(Pn, n))
    Y = Yp[rng.integers(0, Pn, size=2000)] + 0.1 * rng.normal(size=(2000, n))
    ctl = np.ones(n, dtype=bool)
    out = sm.pseudoRUVIII(Y, Yp, types, ctl, k=4, BSPARAM="random", return_info=True)
    fa = np.asarray(out["fullalpha"])
    assert fa.shape == (Pn - G, n)
    # oracle: exact SVD of the residual pseudobulk matrix with every non-null singular vector
    Y0 = orc.my_residop(Yp, orc.replicate_matrix(types))
    U, S, Vt = np.linalg.svd(Y0, full_matrices=False)
    rfa = U[:, :Pn - G].T @ Yp
    assert np.allclose(np.linalg.norm(fa, axis=1), S[:Pn - G], rtol=1e-8)
    assert orc.rel_frobenius(fa.T @ fa, (Vt[:Pn - G].T * S[:Pn - G] ** 2) @ Vt[:Pn - G]) < 1e-9
    assert orc.principal_angles(fa[:4], rfa[:4]).max() < 1e-6


def test_short_matrix_routes_agree(sm, monkeypatch):
    """A pseudobulk matrix with fewer rows than genes and n <= 4096 can take either whole-spectrum route: the m x m side
    (fit_dual.cuh, default) or the rank-deficient n x n Gram (eig_full.cuh, SCM_FIT_DUAL=0).  Same singular values, same
    fullalpha' fullalpha, same adjusted cells."""
    rng = np.random.default_rng(11)
    Pn, n, G = 260, 700, 5
    types = rng.integers(0, G, size=Pn)
    Yp = rng.normal(size=(G, n))[types] * 2 + rng.normal(size=(Pn, 6)) @ rng.normal(size=(6, n)) + 0.4 * rng.normal(size=(Pn, n))
    Y = Yp[rng.integers(0, Pn, size=1500)] + 0.1 * rng.normal(size=(1500, n))
    ctl = np.ones(n, dtype=bool)
    a = sm.pseudoRUVIII(Y, Yp, types, ctl, k=6, BSPARAM="random", return_info=True)
    monkeypatch.setenv("SCM_FIT_DUAL", "0")
    b = sm.pseudoRUVIII(Y, Yp, types, ctl, k=6, BSPARAM="random", return_info=True)
    fa, fb = np.asarray(a["fullalpha"]), np.asarray(b["fullalpha"])
    assert fa.shape == fb.shape == (Pn - G, n)
    assert np.allclose(np.linalg.norm(fa, axis=1), np.linalg.norm(fb, axis=1), rtol=1e-7)
    assert orc.rel_frobenius(fa.T @ fa, fb.T @ fb) < 1e-9
    assert orc.rel_frobenius(a["newY"], b["newY"]) < 1e-9
