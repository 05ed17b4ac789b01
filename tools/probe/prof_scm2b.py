import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import scmerge_b200 as sm, bench
from scmerge_b200 import ruv
warnings.simplefilter("ignore")
class A: pass
a = A(); a.cells=int(sys.argv[2]) if len(sys.argv)>2 else 1_000_000; a.genes=2000; a.ctl=500; a.k=20; a.types=20; a.batches=4; a.workload="scmerge2"
torch.cuda.set_device(0)
own = sys.argv[1] == "own"
ctx = sm.Context(0, stream=None if own else torch.cuda.current_stream().cuda_stream)
dY, keys, ctl = bench.make_data(a, ctx, 0, torch)
torch.cuda.synchronize()
pb = sm.pseudobulk_keys(a._batch, keys, 30, seed=1)
def T(): ctx.sync(); torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0=T(); sm.cosineNorm(dY); t1=T()
    kk, P, names, pb_type, _ = pb
    dP = sm.DeviceMatrix(ctx, P, 2000); d_keys = ctx.to_device(kk); d_cnt = ctx.empty((P,), np.int64)
    from scmerge_b200 import _lib as L
    L.check(L.lib().scm_seg_colmean(ctx.handle, dY.ptr, dY.m, 2000, dY.ld, d_keys.ptr, P, dP.ptr, d_cnt.ptr, None)); t2=T()
    rk, G = ruv.replicate_keys(pb_type)
    fit = ruv._fit(ctx, dP, rk, G, 50, 20, True); t3=T()
    print(f"{'own' if own else 'legacy'} it{it}: cosine {1e3*(t1-t0):.1f} ms, pseudobulk {1e3*(t2-t1):.1f} ms, fit {1e3*(t3-t2):.1f} ms iters {fit.iters}  timers {ctx.timers()}", flush=True)
