"""Timing of the whole-spectrum solver at the sizes a fit meets (SCM_DEBUG_EIG=1 prints the sweeps)."""
import ctypes as C, time, sys
import numpy as np
import scmerge_b200 as sm
from scmerge_b200 import _lib as L
ctx = sm.default_context()
lib = sm.lib()
for n, s, kind in [(500, 490, "bulk"), (1000, 500, "bulk"), (2000, 500, "bulk"), (2000, 500, "lowrank")]:
    rng = np.random.default_rng(n)
    if kind == "bulk":
        A = rng.normal(size=(4 * n, n)) + rng.normal(size=(4 * n, 20)) @ rng.normal(size=(20, n)) * 2
    else:
        A = rng.normal(size=(600, n)) + rng.normal(size=(600, 20)) @ rng.normal(size=(20, n)) * 2
    Gm = A.T @ A
    d_G = ctx.to_device(Gm)
    d_vals, d_vecs = ctx.empty((s,)), ctx.empty((s, n))
    it, rs = C.c_int32(0), C.c_double(0.0)
    for rep in range(2):
        ctx.sync(); t0 = time.perf_counter()
        rc = lib.scm_eig_topk(ctx.handle, d_G.ptr, n, s, s, 0, 1e-10, 500, 1, d_vals.ptr, d_vecs.ptr, C.byref(it), C.byref(rs))
        ctx.sync(); dt = time.perf_counter() - t0
    w = np.linalg.eigvalsh(Gm)[::-1]
    vals = d_vals.to_host()
    print(f"n={n} s={s} {kind}: rc={rc} sweeps={it.value} resid={rs.value:.2e} {dt*1e3:.1f} ms  max rel err vals {np.abs(vals-w[:s]).max()/w[0]:.2e}", flush=True)
