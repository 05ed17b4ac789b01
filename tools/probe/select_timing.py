"""Timing of the ruvK selection stages (scm_pca_scores + scm_silhouette) at a given size.  CELLS=200000 python tools/probe/select_timing.py"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import scmerge_b200 as sm
from scmerge_b200.ruv import _Coef, _fit

m, n, k = int(os.environ.get("CELLS", 200_000)), 2000, 20
torch.cuda.set_device(0)
ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda"); g.manual_seed(1)
T, Bt = 20, 4
types = torch.randint(0, T, (m,), generator=g, device="cuda")
batch = torch.randint(0, Bt, (m,), generator=g, device="cuda")
theta = torch.randn(T, n, generator=g, device="cuda", dtype=torch.float64); theta[:, :500] = 0
A = torch.randn(k, n, generator=g, device="cuda", dtype=torch.float64) * torch.linspace(3, 1.5, k, device="cuda", dtype=torch.float64)[:, None]
dY = sm.DeviceMatrix(ctx, m, n)
Yt = torch.as_tensor(dY.buf, device="cuda").view(m, dY.ld)
for r0 in range(0, m, 50_000):
    r1 = min(m, r0 + 50_000)
    Wb = torch.randn(r1 - r0, k, generator=g, device="cuda", dtype=torch.float64)
    Wb[:, 0] = (batch[r0:r1] == 0).double()
    Yt[r0:r1] = theta[types[r0:r1]] + Wb @ A + 0.5 * torch.randn(r1 - r0, n, generator=g, device="cuda", dtype=torch.float64)
torch.cuda.synchronize()
tk, bk = types.to(torch.int32).cpu().numpy(), batch.to(torch.int32).cpu().numpy()
ctl = np.zeros(n, bool); ctl[:500] = True
t0 = time.perf_counter()
fit = _fit(ctx, dY, tk, T, 50, k, True, keep_gram=True)
ctx.sync(); t1 = time.perf_counter()
print(f"fit (keep_gram) {1e3 * (t1 - t0):.1f} ms")
lib = sm.lib()
d_scores = ctx.empty((m, 10))
d_ct, d_b = ctx.to_device(tk), ctx.to_device(bk)
for kk in (5, 20):
    with _Coef(ctx, n, fit.fullalpha, kk, ctl) as coef:
        it, rs = C.c_int32(), C.c_double()
        ctx.sync(); t0 = time.perf_counter()
        rc = lib.scm_pca_scores(ctx.handle, fit.d_gram.ptr, n, m, coef.handle, fit.d_colmean.ptr, 10, dY.ptr, m, dY.ld, d_scores.ptr, None,
                                1e-9, 500, C.byref(it), C.byref(rs))
        ctx.sync(); t1 = time.perf_counter()
        sa, sb = C.c_double(), C.c_double()
        rc2 = lib.scm_silhouette(ctx.handle, d_scores.ptr, m, 10, 10, d_ct.ptr, T, d_b.ptr, Bt, 0, m, C.byref(sa), C.byref(sb))
        ctx.sync(); t2 = time.perf_counter()
        pairs = float(m) * m
        print(f"k={kk}: pca_scores rc={rc} iters={it.value} {1e3 * (t1 - t0):.1f} ms; silhouette rc={rc2} {1e3 * (t2 - t1):.1f} ms "
              f"({pairs / (t2 - t1) * 1e-9:.1f} Gpair/s, ~{pairs * 42 / (t2 - t1) * 1e-12:.1f} TFLOP/s fp64); sil ct {sa.value / m:.4f} batch {sb.value / m:.4f}")
