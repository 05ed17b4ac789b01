"""GPU tests at BASELINE.json-like sizes, through size-independent properties and the reduced-form oracle
(the literal oracle needs a thin SVD of the whole matrix, which takes minutes at these sizes).

  configs[1]: fastRUVIII 50k cells x 2000 genes, 500 ctl genes, k = 20, exact vs randomized
  properties: idempotence of the adjustment (alpha_c B = I  =>  adjust(newY) == newY), diag of the shifted Gram,
              bit-exact group counts, cell-permutation equivariance, linearity of the adjust step.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import ruv_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    import scmerge_b200 as s
    return s


@pytest.fixture(scope="module")
def c2():
    return orc.planted_factors(50_000, 2000, 500, 20, n_types=20, n_batch=4, seed=20242)


def test_config2_exact_vs_reduced_oracle(sm, c2):
    ref = orc.fastRUVIII_reduced(c2["Y"], c2["types"], c2["ctl"], 20)  # Gram + eigh on the CPU (== literal form, test_oracle.py)
    out = sm.fastRUVIII(c2["Y"], c2["types"], c2["ctl"], k=20, return_info=True)
    assert out["fullalpha"].shape == (50, 2000)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < 1e-8
    assert orc.principal_angles(out["fullalpha"][:20], ref["fullalpha"][:20]).max() < 1e-6
    # randomized mode: the reference's rule asks for min(m - G, n_ctl) = 500 rows of fullalpha: all returned (whole spectrum of the
    # 2000 x 2000 Gram), the leading 20 equal to the exact fit's converged ones, the adjusted matrix the exact one
    rnd = sm.fastRUVIII(c2["Y"], c2["types"], c2["ctl"], k=20, BSPARAM=sm.RandomParam(), return_info=True)
    assert rnd["fullalpha"].shape == (500, 2000) and rnd["fullalpha"].conv_rank == 500
    assert np.abs(rnd["newY"] - out["newY"]).sum() / np.abs(out["newY"]).sum() < 0.02  # tests/testthat/test-fastRUVIII.R:11-16
    assert orc.rel_frobenius(rnd["newY"], ref["newY"]) < 1e-8
    assert np.allclose(np.linalg.norm(rnd["fullalpha"][:20], axis=1), np.linalg.norm(out["fullalpha"][:20], axis=1), rtol=1e-9)
    ref500 = orc.fastRUVIII_reduced(c2["Y"], c2["types"], c2["ctl"], 20, svd_k=500)["fullalpha"]
    fa = np.asarray(rnd["fullalpha"])
    assert np.allclose(np.linalg.norm(fa, axis=1), np.linalg.norm(ref500, axis=1), rtol=1e-9)  # all 500 singular values
    assert orc.rel_frobenius(fa.T @ fa, ref500.T @ ref500) < 1e-9  # rotation-invariant: the bulk is nearly degenerate
    # idempotence: the adjusted matrix is a fixed point of the adjustment with the same alpha
    again = sm.fastRUVIII(out["newY"], c2["types"], c2["ctl"], k=20, fullalpha=out["fullalpha"])
    assert orc.rel_frobenius(again, out["newY"]) < 1e-10


def test_randomized_mode_with_small_rank(sm):
    """Genuine rsvd semantics (block = s + 10, three products): few control genes keep s below the block limit."""
    P = orc.planted_factors(20_000, 600, 40, 12, n_types=8, n_batch=3, seed=7)
    ref = orc.fastRUVIII_reduced(P["Y"], P["types"], P["ctl"], 12)
    out = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=12, BSPARAM=sm.RandomParam(), return_info=True)
    assert out["fullalpha"].shape[0] == 40  # min(m - G, n_ctl): the reference's rule, no truncation
    assert np.abs(out["newY"] - ref["newY"]).sum() / np.abs(ref["newY"]).sum() < 0.02


def test_large_stage_properties(sm):
    from scmerge_b200 import _lib as L
    from scmerge_b200.ruv import segmented_sums
    ctx = sm.default_context()
    rng = np.random.default_rng(99)
    m, n, G = 400_000, 512, 37
    Y = rng.standard_normal((m, n)) * 1.5 + rng.standard_normal(n) * 3
    keys = rng.integers(0, G, size=m).astype(np.int32)
    dY = sm.DeviceMatrix.from_host(ctx, Y)
    sums, cnt = segmented_sums(dY, keys, G)
    assert np.array_equal(cnt, np.bincount(keys, minlength=G))  # bit-exact
    ref = np.zeros((G, n))
    np.add.at(ref, keys, Y)
    assert np.allclose(sums, ref, rtol=1e-11, atol=1e-9)
    a = Y.mean(axis=0)
    dG, da = ctx.empty((n, n)), ctx.to_device(a)
    L.check(L.lib().scm_gram(ctx.handle, dY.ptr, m, n, dY.ld, da.ptr, dG.ptr))
    Gm = dG.to_host()
    assert np.allclose(np.diag(Gm), ((Y - a) ** 2).sum(axis=0), rtol=1e-12)  # checksum of the Gram: its diagonal
    sub = rng.choice(n, 40, replace=False)
    Z = Y[:, sub] - a[sub]
    assert np.allclose(Gm[np.ix_(sub, sub)], Z.T @ Z, rtol=1e-11, atol=1e-6)
    assert np.array_equal(Gm, Gm.T)
    dY.free()


def test_permutation_and_linearity_at_scale(sm):
    P = orc.planted_factors(120_000, 1000, 300, 10, n_types=12, n_batch=4, seed=123)
    out = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10, return_info=True)
    perm = np.random.default_rng(1).permutation(P["Y"].shape[0])
    outp = sm.fastRUVIII(P["Y"][perm], P["types"][perm], P["ctl"], k=10)
    assert np.abs(out["newY"][perm] - outp).sum() / np.abs(outp).sum() < 1.5e-8  # test-resampling-columns.R:26
    # with alpha fixed the adjustment is linear in Y: adjust(2Y) = 2 adjust(Y)
    twice = sm.fastRUVIII(2.0 * P["Y"], P["types"], P["ctl"], k=10, fullalpha=out["fullalpha"])
    assert orc.rel_frobenius(twice, 2.0 * out["newY"]) < 1e-12


def test_full_size_1m_x_2000_properties_and_k_sweep(sm):
    """BASELINE.json's metric size (1M cells x 2000 genes, 500 controls) and its ruvK sweep (configs[4]: k = 5/10/20/50),
    device-resident, through size-independent properties checked with torch on the device (no 16 GB host copies):
    bit-exact group counts, trace / symmetry of the centred Gram, the column means, idempotence of the adjustment for
    every k (alpha_c B = I  =>  adjust(newY) == newY) and removal of the planted unwanted factors."""
    import torch
    from scmerge_b200.ruv import _adjust, _fit
    m, n, c, T, kp = 1_000_000, 2000, 500, 20, 20
    dev = torch.device("cuda", 0)
    ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(2024)
    theta = torch.randn(T, n, generator=g, device=dev, dtype=torch.float64)
    theta[:, :c] = 0                                               # controls carry no biology
    A = torch.randn(kp, n, generator=g, device=dev, dtype=torch.float64) * torch.linspace(3.0, 1.5, kp, device=dev, dtype=torch.float64)[:, None]
    types = torch.randint(0, T, (m,), generator=g, device=dev)
    dY = sm.DeviceMatrix(ctx, m, n)
    Yt = torch.as_tensor(dY.buf, device=dev).view(m, dY.ld)
    Wall = torch.empty(m, kp, device=dev, dtype=torch.float64)
    for r0 in range(0, m, 100_000):
        r1 = min(m, r0 + 100_000)
        Wb = torch.randn(r1 - r0, kp, generator=g, device=dev, dtype=torch.float64)
        Wall[r0:r1] = Wb
        Yt[r0:r1] = theta[types[r0:r1]] + Wb @ A + 0.5 * torch.randn(r1 - r0, n, generator=g, device=dev, dtype=torch.float64)
    torch.cuda.synchronize()
    keys = types.to(torch.int32).cpu().numpy()
    ctl = np.zeros(n, dtype=bool)
    ctl[:c] = True
    fit = _fit(ctx, dY, keys, T, 50, kp, True, keep_gram=True)
    assert fit.fullalpha.shape == (50, n) and np.all(np.isfinite(fit.fullalpha))
    # column means and centred Gram against torch reductions on the device
    mean_t = Yt.mean(dim=0)
    assert torch.allclose(torch.as_tensor(fit.d_colmean, device=dev), mean_t, rtol=1e-12, atol=1e-13)
    Gc = torch.as_tensor(fit.d_gram, device=dev).view(n, n)
    assert torch.equal(Gc, Gc.T)
    ssq = torch.zeros(n, device=dev, dtype=torch.float64)
    for r0 in range(0, m, 100_000):
        ssq += ((Yt[r0:r0 + 100_000] - mean_t) ** 2).sum(dim=0)
    assert torch.allclose(torch.diagonal(Gc), ssq, rtol=1e-11)
    sub = torch.arange(0, n, 97, device=dev)
    Zs = Yt[:200_000][:, sub] - mean_t[sub]                         # a 21 x 21 block of the Gram of the first 200k cells ...
    Gfull = torch.zeros(len(sub), len(sub), device=dev, dtype=torch.float64)
    for r0 in range(0, m, 200_000):
        Zs = Yt[r0:r0 + 200_000][:, sub] - mean_t[sub]
        Gfull += Zs.T @ Zs
    assert torch.allclose(Gc[sub][:, sub], Gfull, rtol=1e-10, atol=1e-6)
    fit.d_gram.free()
    fit.d_colmean.free()
    # the planted factor space is recovered: rows of A lie in the row space of alpha_k
    ang = orc.principal_angles(A.cpu().numpy(), fit.fullalpha[:kp])
    assert ang.max() < 0.05
    dOut, dAgain = sm.DeviceMatrix(ctx, m, n), sm.DeviceMatrix(ctx, m, n)
    Ot = torch.as_tensor(dOut.buf, device=dev).view(m, dOut.ld)
    At = torch.as_tensor(dAgain.buf, device=dev).view(m, dAgain.ld)
    for k in (5, 10, 20, 50):
        _adjust(ctx, dY, fit.fullalpha, k, ctl, out=dOut)
        _adjust(ctx, dOut, fit.fullalpha, k, ctl, out=dAgain)
        torch.cuda.synchronize()
        num = den = 0.0
        for r0 in range(0, m, 250_000):
            num += float(((At[r0:r0 + 250_000] - Ot[r0:r0 + 250_000]) ** 2).sum())
            den += float((Ot[r0:r0 + 250_000] ** 2).sum())
        assert (num / den) ** 0.5 < 1e-10, (k, (num / den) ** 0.5)   # idempotent
        if k == 20:
            # the unwanted factors are gone: the adjusted cells no longer correlate with the planted W
            blk = Ot[:200_000, c:] - theta[types[:200_000]][:, c:]
            before = Yt[:200_000, c:] - theta[types[:200_000]][:, c:]
            cw = Wall[:200_000]
            r_after = float((cw.T @ blk).norm() / (cw.norm() * blk.norm()))
            r_before = float((cw.T @ before).norm() / (cw.norm() * before.norm()))
            assert r_before > 0.1 and r_after < 0.05 * r_before, (r_before, r_after)   # measured 0.223 -> 0.002
    for d_ in (dY, dOut, dAgain):
        d_.free()


def test_wide_matrix_20000_genes(sm):
    """Whole-transcriptome width (n = 20000 genes, what scMerge2 is usually given): 12403 Gram tiles, 1250 gene blocks in the
    rank-k update, shared-memory row images beyond one slab in the sparse kernels.  Small m keeps the oracle cheap."""
    import scipy.sparse as sp
    from scmerge_b200.scmerge2 import pseudobulk_keys, scMerge2_core
    P = orc.planted_factors(600, 20000, 2000, 5, n_types=4, n_batch=3, seed=5)
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 5, return_info=True)
    out = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=5, return_info=True)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < 1e-8
    assert orc.principal_angles(out["fullalpha"][:5], ref["fullalpha"][:5]).max() < 1e-6
    # scMerge2 core on a sparse 20000-gene matrix: sparse-native kernels and the densify-first path agree with the oracle
    rng = np.random.default_rng(3)
    m, n = 1500, 20000
    batch, clust = rng.integers(0, 3, m), rng.integers(0, 3, m)
    D = rng.gamma(2.0, 1.0, (m, n)) * (rng.random((m, n)) < 0.03)
    D += (rng.gamma(2.0, 1.0, (3, n)) * (rng.random((3, n)) < 0.05))[clust] + (rng.gamma(2.0, 0.7, (3, 3, n)) * (rng.random((3, 3, n)) < 0.03))[batch, clust]
    keys, npb, names, pb_t, pb_b = pseudobulk_keys(batch, clust, 6, seed=1)
    which = np.array([int(nm.rsplit("_", 1)[1]) for nm in names])[keys]
    refs = orc.scmerge2_core(D.T, batch, clust, which, ruvK=6, k_fold=6, exact=True)
    Xs = sp.csr_matrix(D)
    for native in (True, False):
        got = scMerge2_core(Xs, batch, clust, ruvK=6, k_pseudoBulk=6, which=which, BSPARAM=sm.ExactParam(), sparse_native=native)
        assert got["pseudobulk_names"] == refs["names"]
        assert orc.rel_frobenius(got["newY"], refs["newY"]) < 1e-8, native
