"""Phase trace of the end-to-end host driver (SCM_HOST_TRACE=1) at the bench shape.  Usage: python tools/probe/e2e_trace.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["SCM_HOST_TRACE"] = "1"
import numpy as np
import scmerge_b200 as sm

m, n = int(os.environ.get("CELLS", 1_000_000)), 2000
ctx = sm.default_context(0)
rng = np.random.default_rng(0)
Yh, Oh = sm.pinned_empty((m, n)), sm.pinned_empty((m, n))
blk = rng.normal(size=(10000, n))
for r0 in range(0, m, 10000):
    Yh[r0:r0 + 10000] = blk[: min(10000, m - r0)] + 0.01 * (r0 // 10000)
keys = (np.arange(m) % 20).astype(np.int32)
ctl = np.zeros(n, bool); ctl[:500] = True
for it in range(3):
    t0 = time.perf_counter()
    sm.fastRUVIII(Yh, keys, ctl, k=20, out=Oh, ctx=ctx)
    print(f"call {it}: {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr)
