import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scmerge_b200 as sm
from scmerge_b200 import ruv
from oracle import ruv_oracle as orc
warnings.simplefilter("always")
P = orc.planted_factors(120_000, 1000, 300, 10, n_types=12, n_batch=4, seed=123)
ref = orc.fastRUVIII_reduced(P["Y"], P["types"], P["ctl"], 10)
ctx = sm.default_context()
keys, G = ruv.replicate_keys(P["types"])
# device-resident path
dY = sm.DeviceMatrix.from_host(ctx, P["Y"])
fit = ruv._fit(ctx, dY, keys, G, 50, 10, True)
print("device fit iters", fit.iters, "resid", fit.resid, "angle", orc.principal_angles(fit.fullalpha[:10], ref["fullalpha"][:10]).max())
print("sigma dev", fit.sigma[:12]); print("sigma ref", np.linalg.norm(ref["fullalpha"], axis=1)[:12])
out, _ = ruv._adjust(ctx, dY, fit.fullalpha, 10, P["ctl"])
print("device-resident newY err", orc.rel_frobenius(out.to_host(), ref["newY"]))
o2 = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10, return_info=True)
print("host pipeline newY err", orc.rel_frobenius(o2["newY"], ref["newY"]), "angle", orc.principal_angles(o2["fullalpha"][:10], ref["fullalpha"][:10]).max())
perm = np.random.default_rng(1).permutation(P["Y"].shape[0])
o3 = sm.fastRUVIII(P["Y"][perm], P["types"][perm], P["ctl"], k=10, return_info=True)
print("perm host pipeline newY err", orc.rel_frobenius(o3["newY"], ref["newY"][perm]), "angle", orc.principal_angles(o3["fullalpha"][:10], ref["fullalpha"][:10]).max())
# adjust with the oracle's alpha: isolates K5/K6
o4, _ = ruv._adjust(ctx, dY, ref["fullalpha"], 10, P["ctl"])
print("adjust with oracle alpha err", orc.rel_frobenius(o4.to_host(), ref["newY"]))
