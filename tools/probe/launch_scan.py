"""One call of every closure at a realistic size, for an ncu launch list (finds small kernels that are unexpectedly slow):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/probe/launch_scan.py
    python tools/probe/launch_scan.py --show out.csv
"""
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

if len(sys.argv) > 2 and sys.argv[1] == "--show":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[k].split("(")[0][-70:]
        if "at::" in name or "cutlass" in name or "at_cuda" in name:
            continue
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += float(r[v])
        a[2] = max(a[2], float(r[v]))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print(f"{n:72s} {a[0]:5d} launches {a[1] / 1e6:9.3f} ms total {a[2] / 1e3:10.1f} us longest")
    sys.exit(0)

import warnings  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import scmerge_b200 as sm  # noqa: E402

warnings.simplefilter("ignore")


class A:
    pass


a = A()
m, n = 200_000, 2000
a.cells, a.genes, a.ctl, a.k, a.types, a.batches, a.workload, a.density = m, n, 500, 20, 20, 4, "fastruviii", 0.1
torch.cuda.set_device(0)
ctx = sm.Context(0)
dY, types, ctl, batch = bench.make_data(a, ctx, 0, torch, m, scm2=True)
for rep in range(2):  # the second round is the one to read
    s1 = sm.scRUVIII(dY, types, ctl, k=20, batch=batch, cell_type=types, ctx=ctx)
    s1[20]["newY"].free()
    sub = slice(0, 20_000)
    dsub = sm.DeviceMatrix.from_host(ctx, dY.to_host()[sub])
    s3 = sm.scRUVIII(dsub, types[sub], ctl, k=[5, 10, 20], batch=batch[sub], cell_type=types[sub], return_all_RUV=False, ctx=ctx)
    s3["newY"].free()
    dsub.free()
    g = sm.scRUVg(dY, ctl, 10, ctx=ctx)
    g["newY"].free()
    pb = sm.pseudobulk_keys(batch, types, 30, seed=1)
    dOut = sm.DeviceMatrix(ctx, m, n)
    r = sm.scMerge2_core(dY, None, None, ctl=None, ruvK=20, pb=pb, return_bulk=False, ctx=ctx, out=dOut, BSPARAM=sm.ExactParam())
    ad = sm.getAdjustedMat(dOut, r["fullalpha"], ruvK=10, ctx=ctx)
    ad.free()
    dOut.free()
    ctx.sync()
print("done")
