"""Times scm_csc_seg_colmean (pseudobulk means on sparse rows) at 1M cells and checks it against a saved result.
    python tools/probe/csc_seg_probe.py [cells] [--save F | --check F]"""
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

p = argparse.ArgumentParser()
p.add_argument("cells", nargs="?", type=int, default=1_000_000)
p.add_argument("--save")
p.add_argument("--check")
o = p.parse_args()


class A:
    pass


a = A()
a.cells, a.genes, a.ctl, a.k, a.types, a.batches, a.workload, a.density = o.cells, 2000, 500, 20, 20, 4, "scmerge2", 0.1
torch.cuda.set_device(0)
ctx = sm.Context(0)
dY, keys, ctl, b = bench.make_data(a, ctx, 0, torch, o.cells, scm2=True)
Xh = bench.sparsify(a, ctx, dY, 0, torch)
dY.free()
m, n = o.cells, 2000
src = sm.DeviceCSC(ctx, Xh.indptr, Xh.indices, Xh.data, m, n)
d_inv = src.row_scales(cosine=True)
kk, P, names, pb_type, _ = sm.pseudobulk_keys(b, keys, 30, seed=1)
d_keys, d_cnt = ctx.to_device(np.ascontiguousarray(kk, dtype=np.int32)), ctx.empty((P,), np.int64)
dP = sm.DeviceMatrix(ctx, P, n)
lib = sm.lib()


def run():
    L.check(lib.scm_csc_seg_colmean(ctx.handle, src.indptr.ptr, src.indices.ptr, src.data.ptr, d_inv.ptr, m, n, d_keys.ptr, P, dP.ptr, dP.ld, d_cnt.ptr))


for _ in range(2):
    run()
ctx.sync()
ts = []
for _ in range(5):
    ctx.sync()
    t0 = time.perf_counter()
    run()
    ctx.sync()
    ts.append(1e3 * (time.perf_counter() - t0))
print(f"scm_csc_seg_colmean m={m} n={n} P={P} nnz={Xh.nnz}: best {min(ts):.2f} ms (all {['%.2f' % t for t in ts]}), {12.0 * Xh.nnz / min(ts) * 1e-6:.0f} GB/s", flush=True)
res = dP.to_host()
if o.save:
    np.save(o.save, res)
if o.check:
    ref = np.load(o.check)
    print("bit-identical to", o.check, ":", np.array_equal(res, ref), " max abs diff", np.abs(res - ref).max(), flush=True)
    assert np.array_equal(res, ref)
