import sys, os, time, cProfile, pstats, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import scmerge_b200 as sm, bench
warnings.simplefilter("ignore")
class A: pass
a = A(); a.cells=1_000_000; a.genes=2000; a.ctl=500; a.k=20; a.types=20; a.batches=4; a.workload="scmerge2"
torch.cuda.set_device(0)
ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
dY, keys, ctl = bench.make_data(a, ctx, 0, torch)
pb = sm.pseudobulk_keys(a._batch, keys, 30, seed=1)
step = lambda: sm.scMerge2_core(dY, None, None, ctl=None, ruvK=20, pb=pb, return_bulk=False, ctx=ctx)
step(); step()
pr = cProfile.Profile(); pr.enable(); t0=time.perf_counter(); step(); ctx.sync(); t1=time.perf_counter(); pr.disable()
print("step wall ms", 1e3*(t1-t0))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
