"""Times scm_kmeans at the size identifyCluster meets (one batch: cells x 10 PCs, nstart = 1000) against scikit-learn's Lloyd
(n_init reduced, scaled) on the host.   python tools/probe/kmeans_probe.py [cells] [K]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

import scmerge_b200 as sm  # noqa: E402
from scmerge_b200 import replicate as rp  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
cent = rng.normal(size=(K, 10)) * 2.0
x = cent[rng.integers(0, K, m)] + rng.normal(size=(m, 10))
ctx = sm.default_context()
dX = sm.DeviceMatrix.from_host(ctx, x)
for nstart in (10, 1000):
    rp.kmeans(dX, K, nstart=nstart)
    ctx.sync()
    t0 = time.perf_counter()
    r = rp.kmeans(dX, K, nstart=nstart)
    ctx.sync()
    t = time.perf_counter() - t0
    print(f"m={m} K={K} nstart={nstart}: {1e3 * t:.1f} ms  tot.withinss {r['tot_withinss']:.6f}  Lloyd iterations of the winner {r['iter']}  "
          f"transfers {r['transfers']}  sizes {r['size'].tolist()}", flush=True)
try:
    from sklearn.cluster import KMeans
    t0 = time.perf_counter()
    km = KMeans(K, n_init=20, algorithm="lloyd", random_state=0).fit(x)
    t = time.perf_counter() - t0
    print(f"scikit-learn Lloyd, n_init=20, {os.cpu_count()} cores: {1e3 * t:.0f} ms (x50 for 1000 starts: {50 * t:.1f} s)  inertia {km.inertia_:.6f}")
except Exception as e:  # noqa: BLE001
    print("sklearn:", e)
