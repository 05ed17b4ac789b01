"""Wall-clock breakdown of one device-resident fastRUVIII step (host overheads vs kernels)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import scmerge_b200 as sm
from scmerge_b200 import ruv
import bench
class A: pass
a = A(); a.cells=int(sys.argv[1]) if len(sys.argv)>1 else 1_000_000; a.genes=2000; a.ctl=500; a.k=20; a.types=20; a.batches=4; a.workload="fastruviii"
torch.cuda.set_device(0)
ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
dY, keys, ctl = bench.make_data(a, ctx, 0, torch)
dOut = sm.DeviceMatrix(ctx, a.cells, a.genes)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    t0 = T()
    k2, G = ruv.replicate_keys(keys); t1 = T()
    fit = ruv._fit(ctx, dY, k2, G, 50, 20, True); t2 = T()
    out, _ = ruv._adjust(ctx, dY, fit.fullalpha, 20, ctl, out=dOut); t3 = T()
    print(f"it{it}: keys {1e3*(t1-t0):.1f} ms, fit {1e3*(t2-t1):.1f} ms (iters {fit.iters}), adjust {1e3*(t3-t2):.1f} ms, timers {ctx.timers()}")
# finer: inside fit
import ctypes as C
from scmerge_b200 import _lib as L
lib = L.lib()
for it in range(3):
    t0=T(); d_keys = ctx.to_device(keys.astype(np.int32)); t1=T()
    d_fa = ctx.empty((50, 2000)); d_sig = ctx.empty((50,)); t2=T()
    iters, resid = C.c_int32(0), C.c_double(0.0)
    rc = lib.scm_ruv3_fit(ctx.handle, dY.ptr, dY.m, 2000, dY.ld, d_keys.ptr, 20, None, 0, 50, 20, 0, 1e-10, 300, 1, d_fa.ptr, d_sig.ptr, None, None, C.byref(iters), C.byref(resid)); t3=T()
    fa = d_fa.to_host(); t4=T()
    d_keys.free(); d_fa.free(); d_sig.free(); t5=T()
    print(f"  upload keys {1e3*(t1-t0):.2f}, alloc {1e3*(t2-t1):.2f}, scm_ruv3_fit {1e3*(t3-t2):.2f}, d2h {1e3*(t4-t3):.2f}, free {1e3*(t5-t4):.2f}  rc={rc}")
