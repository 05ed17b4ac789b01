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
    # randomized mode: the reference's rule asks for min(m - G, n_ctl) = 500 rows -> truncated + converged on the device
    with pytest.warns(UserWarning, match="truncated"):
        rnd = sm.fastRUVIII(c2["Y"], c2["types"], c2["ctl"], k=20, BSPARAM=sm.RandomParam())
    assert np.abs(rnd - out["newY"]).sum() / np.abs(out["newY"]).sum() < 0.02  # tests/testthat/test-fastRUVIII.R:11-16
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
