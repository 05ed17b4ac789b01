"""A/B timing of scm_gram / scm_adjust across several builds of the library IN ONE PROCESS on one box
(box-to-box variation is larger than the effects being measured).

    python tools/probe/gram_ab.py build/libscm_prev.so scmerge_b200/libscmerge_b200.so [cells]

Every library gets its own context; the same device matrix is passed to each; launches alternate between the
libraries.  Prints the best / median kernel time per library (CUDA events around the call) and whether the Grams are
bit-identical to the first library's.
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from scmerge_b200 import _lib as L  # noqa: E402  (signatures only)

paths = [a for a in sys.argv[1:] if a.endswith(".so")]
rest = [a for a in sys.argv[1:] if not a.endswith(".so")]
m = int(rest[0]) if rest else 1_000_000
n = int(rest[1]) if len(rest) > 1 else 2000
torch.cuda.set_device(0)
stream = torch.cuda.current_stream().cuda_stream

libs = []
for p in paths:
    h = C.CDLL(os.path.abspath(p))
    for name, args in L.SIGNATURES.items():
        if hasattr(h, name):
            fn = getattr(h, name)
            fn.argtypes = args
            fn.restype = L._RESTYPES.get(name, C.c_int)
    ctx = C.c_void_p()
    assert h.scm_ctx_create(0, C.c_void_p(stream), C.byref(ctx)) == 0, h.scm_last_error()
    libs.append((p, h, ctx))

g = torch.Generator(device="cuda").manual_seed(1)
Y = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g) + 3.0
shift = Y.mean(dim=0).contiguous()
grams = [torch.empty(n, n, dtype=torch.float64, device="cuda") for _ in libs]
out = torch.empty_like(Y)


def timed(fn, h=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, (rc, h.scm_last_error() if h is not None else None)
    return e0.elapsed_time(e1)


times = {p: [] for p, _, _ in libs}
for it in range(7):
    for (p, h, ctx), G in zip(libs, grams):
        t = timed(lambda: h.scm_gram(ctx, Y.data_ptr(), m, n, n, shift.data_ptr(), G.data_ptr()), h)
        if it >= 2:
            times[p].append(t)
for (p, _, _), G in zip(libs, grams):
    ts = sorted(times[p])
    same = bool(torch.equal(G, grams[0]))
    rel = float((G - grams[0]).abs().max() / grams[0].abs().max())
    print(f"gram   {p:45s} best {ts[0]:8.3f} ms  median {ts[len(ts) // 2]:8.3f} ms  bit-identical to first: {same} (max rel diff {rel:.2e})")

# adjust: coefficient object from a random orthonormal alpha
k = 20
alpha = torch.linalg.qr(torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g))[0].T.contiguous()
ctl = torch.zeros(n, dtype=torch.uint8, device="cuda")
ctl[:500] = 1
outs = []
times = {p: [] for p, _, _ in libs}
for (p, h, ctx) in libs:
    coef = C.c_void_p()
    assert h.scm_coef_create(ctx, alpha.data_ptr(), n, k, ctl.data_ptr(), None, None, None, C.byref(coef)) == 0
    for it in range(7):
        t = timed(lambda: h.scm_adjust(ctx, Y.data_ptr(), m, n, n, coef, None, 0, out.data_ptr(), n, None), h)
        if it >= 2:
            times[p].append(t)
    outs.append(out[:4096].clone())
    h.scm_coef_destroy(ctx, coef)
for (p, _, _), o in zip(libs, outs):
    ts = sorted(times[p])
    print(f"adjust {p:45s} best {ts[0]:8.3f} ms  median {ts[len(ts) // 2]:8.3f} ms  bit-identical to first: {bool(torch.equal(o, outs[0]))}")
