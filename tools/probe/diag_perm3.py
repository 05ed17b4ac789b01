import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scmerge_b200 as sm
from scmerge_b200 import ruv
from oracle import ruv_oracle as orc
P = orc.planted_factors(120_000, 1000, 300, 10, n_types=12, n_batch=4, seed=123)
ref = orc.fastRUVIII_reduced(P["Y"], P["types"], P["ctl"], 10)
ctx = sm.default_context()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = sys.argv[2] if len(sys.argv) > 2 else "host"
dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
dO = sm.DeviceMatrix(ctx, dY.m, dY.n)
nbad = 0
for it in range(N):
    if mode == "host":
        a = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10, fullalpha=ref["fullalpha"])   # adjust only, host pipeline (in place, chunked)
    else:
        out, _ = ruv._adjust(ctx, dY, ref["fullalpha"], 10, P["ctl"], out=dO)              # device resident, out of place
        a = out.to_host()
    d = np.abs(a - ref["newY"])
    rows = np.nonzero(d.max(axis=1) > 1e-9)[0]
    if rows.shape[0]:
        nbad += 1
        cols = np.nonzero(d.max(axis=0) > 1e-9)[0]
        print(it, mode, "bad rows", rows.shape[0], rows[:3], rows[-1], "bad cols", cols.shape[0], cols[:12], "max", d.max(), flush=True)
print("mode", mode, "iterations", N, "bad", nbad)
