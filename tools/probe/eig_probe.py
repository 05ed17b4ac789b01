"""K4 microbenchmark: scm_eig_topk on the Gram of a planted-factor matrix (n = 2000, 20 factors), timed with CUDA events.
SCM_DEBUG_EIG=1 prints Rayleigh-Ritz steps / products / Jacobi sweeps.  Run under ncu for the per-kernel list."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import scmerge_b200 as sm  # noqa: E402

n, m, k, T = 2000, 200_000, 20, 20
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
A = torch.randn(k, n, generator=g, device=dev, dtype=torch.float64) * torch.linspace(3.0, 1.5, k, device=dev, dtype=torch.float64)[:, None]
W = torch.randn(m, k, generator=g, device=dev, dtype=torch.float64)
Y = W @ A + 0.5 * torch.randn(m, n, generator=g, device=dev, dtype=torch.float64)
Y -= Y.mean(dim=0)
G = (Y.T @ Y).contiguous()
del Y, W
ctx = sm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
lib = sm.lib()
s = int(sys.argv[1]) if len(sys.argv) > 1 else 50
d_vals, d_vecs = ctx.empty((s,)), ctx.empty((s, n))
it, rs = C.c_int32(0), C.c_double(0.0)
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.scm_eig_topk(ctx.handle, C.c_void_p(G.data_ptr()), n, s, k, 0, 1e-10, 500, 1, d_vals.ptr, d_vecs.ptr, C.byref(it), C.byref(rs))
    e1.record(); torch.cuda.synchronize()
    print(f"rep {rep}: rc {rc} rr_steps {it.value} resid {rs.value:.2e} {e0.elapsed_time(e1):.3f} ms  info {ctx.fit_info()}")
w = torch.linalg.eigvalsh(G).flip(0)[:s].cpu().numpy()
print("max rel eigenvalue error", np.abs(d_vals.to_host() - w).max() / w[0])
