"""Stress the adjust kernel for rare races: many launches, out-of-place and in-place, compared on the device."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import numpy as np, torch
import scmerge_b200 as sm
from scmerge_b200 import ruv, _lib as L
from oracle import ruv_oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
m, n, c, k = 65536, 1000, 300, 10
P = orc.planted_factors(m, n, c, k, n_types=12, n_batch=4, seed=123)
torch.cuda.set_device(0)
ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
keys, G = ruv.replicate_keys(P["types"])
fit = ruv._fit(ctx, dY, keys, G, 50, k, True)
lib = L.lib()
d_alpha = ctx.to_device(np.ascontiguousarray(fit.fullalpha[:k])); d_ctl = ctx.to_device(P["ctl"].astype(np.uint8))
coef = C.c_void_p(); L.check(lib.scm_coef_create(ctx.handle, d_alpha.ptr, n, k, d_ctl.ptr, None, None, None, C.byref(coef)))
dO = sm.DeviceMatrix(ctx, m, n); dW = sm.DeviceMatrix(ctx, m, n)
tY = torch.as_tensor(dY.buf, device="cuda"); tO = torch.as_tensor(dO.buf, device="cuda"); tW = torch.as_tensor(dW.buf, device="cuda")
L.check(lib.scm_adjust(ctx.handle, dY.ptr, m, n, dY.ld, coef, None, 0, dO.ptr, dO.ld, None)); ctx.sync()
ref = tO.clone()
err = orc.rel_frobenius(dO.to_host(), orc.fastRUVIII_reduced(P["Y"], P["types"], P["ctl"], k)["newY"]); print("reference run err vs oracle", err)
for mode in ("out-of-place", "in-place"):
    bad = 0
    for it in range(N):
        if mode == "out-of-place":
            tO.zero_()
            L.check(lib.scm_adjust(ctx.handle, dY.ptr, m, n, dY.ld, coef, None, 0, dO.ptr, dO.ld, None)); res = tO
        else:
            tW.copy_(tY)
            L.check(lib.scm_adjust(ctx.handle, dW.ptr, m, n, dW.ld, coef, None, 0, dW.ptr, dW.ld, None)); res = tW
        ctx.sync()
        d = (res - ref).abs().view(m, n)
        mx = float(d.max())
        if mx > 1e-9:
            bad += 1
            rows = torch.nonzero(d.max(dim=1).values > 1e-9).flatten(); cols = torch.nonzero(d.max(dim=0).values > 1e-9).flatten()
            print(mode, it, "bad rows", rows.numel(), rows[:3].tolist(), "cols", cols.numel(), cols[:16].tolist(), "max", mx, flush=True)
    print(mode, "launches", N, "bad", bad, flush=True)
