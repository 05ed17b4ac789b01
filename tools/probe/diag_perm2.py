import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scmerge_b200 as sm
from oracle import ruv_oracle as orc
P = orc.planted_factors(120_000, 1000, 300, 10, n_types=12, n_batch=4, seed=123)
ref = orc.fastRUVIII_reduced(P["Y"], P["types"], P["ctl"], 10)["newY"]
perm = np.random.default_rng(1).permutation(P["Y"].shape[0])
Yp = P["Y"][perm]; tp = P["types"][perm]; refp = ref[perm]
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    a = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10)
    b = sm.fastRUVIII(Yp, tp, P["ctl"], k=10)
    da = np.abs(a - ref); db = np.abs(b - refp)
    ra = np.nonzero(da.max(axis=1) > 1e-9)[0]; rb = np.nonzero(db.max(axis=1) > 1e-9)[0]
    print(it, "L1 a", da.sum()/np.abs(ref).sum(), "L1 b", db.sum()/np.abs(ref).sum(), "bad rows a", ra.shape[0], ra[:6], "bad rows b", rb.shape[0], rb[:6],
          "maxerr", da.max(), db.max())
    if rb.shape[0]:
        r = rb[0]; cols = np.nonzero(db[r] > 1e-9)[0]; print("   row", r, "bad cols", cols.shape[0], cols[:8], cols[-3:], "vals", b[r, cols[:3]], refp[r, cols[:3]], Yp[r, cols[:3]])
    if ra.shape[0]:
        r = ra[0]; cols = np.nonzero(da[r] > 1e-9)[0]; print("   row", r, "bad cols", cols.shape[0], cols[:8], cols[-3:], "vals", a[r, cols[:3]], ref[r, cols[:3]], P["Y"][r, cols[:3]])
