"""Times scm_csc_adjust (projection + pass B) on a synthetic sparse matrix and checks the RED-based pass B against the tile
kernel.  The pass-B variant is chosen by SCM_CSC_RED, which the library reads ONCE per process: run this script once per
variant;  `--save F` writes a strided sample of the result rows, `--check F` compares against it.

    python tools/probe/csc_adjust_probe.py [cells] [--k 20] [--density 0.1] [--save F | --check F]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import scmerge_b200 as sm  # noqa: E402
from scmerge_b200 import _lib as L  # noqa: E402

if os.environ.get("SCM_LIB_PATH"):  # another build of the library (A/B runs)
    L.LIB_PATH = os.path.abspath(os.environ["SCM_LIB_PATH"])
from scmerge_b200.ruv import _Coef  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("cells", nargs="?", type=int, default=1_000_000)
p.add_argument("--k", type=int, default=20)
p.add_argument("--genes", type=int, default=2000)
p.add_argument("--density", type=float, default=0.1)
p.add_argument("--reps", type=int, default=5)
p.add_argument("--save")
p.add_argument("--check")
o = p.parse_args()


class A:
    pass


a = A()
a.cells, a.genes, a.ctl, a.k, a.types, a.batches, a.workload, a.density = o.cells, o.genes, 500, o.k, 20, 4, "scmerge2", o.density
torch.cuda.set_device(0)
ctx = sm.Context(0)
dY, keys, ctl, _b = bench.make_data(a, ctx, 0, torch, o.cells, scm2=True)
Xh = bench.sparsify(a, ctx, dY, 0, torch)
dY.free()
m, n = o.cells, o.genes
src = sm.DeviceCSC(ctx, Xh.indptr, Xh.indices, Xh.data, m, n)
d_inv = src.row_scales(cosine=True)
rng = np.random.default_rng(7)
alpha = rng.normal(size=(o.k, n))
ctlm = np.ones(n, bool)
center = rng.normal(size=n) * 1e-3
dout = sm.DeviceMatrix(ctx, m, n)
lib = sm.lib()
by = 12.0 * Xh.nnz + 8.0 * m * n
with _Coef(ctx, n, alpha, o.k, ctlm, center=center) as coef:
    def run():
        L.check(lib.scm_csc_adjust(ctx.handle, src.indptr.ptr, src.indices.ptr, src.data.ptr, d_inv.ptr, m, n, coef.handle, dout.ptr, dout.ld))
    for _ in range(2):
        run()
    ctx.sync()
    ts = []
    for _ in range(o.reps):
        ctx.sync()
        t0 = time.perf_counter()
        run()
        ctx.sync()
        ts.append(1e3 * (time.perf_counter() - t0))
    best = min(ts)
    print(f"{os.path.basename(L.LIB_PATH)} SCM_CSC_RED={os.environ.get('SCM_CSC_RED', '(default)')} SCM_CSC_DEBUG={os.environ.get('SCM_CSC_DEBUG', '0')} m={m} n={n} k={o.k} nnz={Xh.nnz}: "
          f"best {best:.2f} ms (all {['%.2f' % t for t in ts]}), {by / best * 1e-6:.0f} GB/s HBM-equivalent", flush=True)
    rows = np.arange(0, m, max(1, m // 4000))
    t = torch.as_tensor(dout.buf, device="cuda:0").view(m, dout.ld)[torch.as_tensor(rows, device="cuda:0"), :n].cpu().numpy()
    if o.save:
        np.save(o.save, t)
    if o.check:
        ref = np.load(o.check)
        err = np.linalg.norm(t - ref) / np.linalg.norm(ref)
        print(f"relative Frobenius difference to {o.check}: {err:.3e}  max abs {np.abs(t - ref).max():.3e}", flush=True)
        assert err < 1e-13, err
