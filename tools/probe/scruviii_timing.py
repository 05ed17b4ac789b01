"""Device-resident scRUVIII (standardise -> fit -> adjust -> de-standardise, R/scRUVIII.R:41-137) at the bench shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import scmerge_b200 as sm, bench
class A: pass
a = A(); a.cells = int(os.environ.get("CELLS", 1_000_000)); a.genes = 2000; a.ctl = 500; a.k = 20; a.types = 20; a.batches = 4; a.workload = "fastruviii"
torch.cuda.set_device(0)
ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
dY, keys, ctl = bench.make_data(a, ctx, 0, torch)
batch = a._batch
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = sm.scRUVIII(dY, keys, ctl, k=20, batch=batch, ctx=ctx)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    tm = ctx.timers()
    res[20]["newY"].free()
    print(f"scRUVIII it{it}: {1e3 * (t1 - t0):.1f} ms  ({a.cells / (t1 - t0) * 1e-6:.2f} M cells/s)  stages", {k: round(v, 2) for k, v in tm.items()})
