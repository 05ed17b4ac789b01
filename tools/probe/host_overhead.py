"""Where the host time of a device-resident fastRUVIII call goes at a strong-scaling shard size (125k cells x 2000 genes)."""
import cProfile, pstats, time, io
import numpy as np
import scmerge_b200 as sm
from oracle import ruv_oracle as orc
ctx = sm.default_context()
P = orc.planted_factors(125_000, 2000, 500, 20, n_types=20, n_batch=4, seed=1)
dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
dOut = sm.DeviceMatrix(ctx, dY.m, dY.n)
keys = P["types"].astype(np.int32)
f = lambda: sm.fastRUVIII(dY, keys, P["ctl"], k=20, ctx=ctx, out=dOut)
for _ in range(5): f()
ctx.sync(); t0 = time.perf_counter()
for _ in range(50): f()
ctx.sync(); dt = (time.perf_counter() - t0) / 50
tm = ctx.timers() if hasattr(ctx, "timers") else {}
print(f"wall per call {dt*1e3:.3f} ms; stages {tm}")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): f()
ctx.sync(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
