"""Measure cuBLAS DGEMM (fp64) peak on this GPU, same method as MEASURED_PEAKS.json's bf16 entry:
torch.matmul float64 8192^3, best of 10 (burst) and back to back for 4 s (sustained)."""
import json, sys, time, torch
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda")
b = torch.randn(N, N, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * N**3 / (best * 1e-3) * 1e-12
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); n = 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5):
        torch.matmul(a, b, out=c); n += 1
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = 2 * N**3 * n / (e0.elapsed_time(e1) * 1e-3) * 1e-12
# HBM copy, f64
x = torch.empty(1 << 30, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
bestc = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    bestc = min(bestc, e0.elapsed_time(e1))
out = {"fp64_dgemm_tflops_burst": burst, "fp64_dgemm_tflops_sustained": sus,
       "hbm_copy_gbs_f64_8GiB": 2 * x.numel() * 8 / (bestc * 1e-3) * 1e-9,
       "gpu": torch.cuda.get_device_name(0), "how": "torch.matmul float64 8192^3 best-of-10 / 4 s loop; y.copy_(x) 8 GiB f64"}
print(json.dumps(out))
json.dump(out, open(sys.argv[1], "w"), indent=1) if len(sys.argv) > 1 else None
