#!/usr/bin/env python
"""bench.py -- RUV-III fit-and-adjust throughput (cells/s) on B200, the metric of BASELINE.json
("RUV-III cells/sec (1M x 2k genes, k=20) at 1/2/4/8 B200").

One "step" = one full fastRUVIII pass (replicate sums + shifted copy -> Gram -> top-k eigenpairs -> coefficients -> fused
adjust) over one synthetic cells x genes matrix.  Headline workload at EVERY N: `fastRUVIII, 1M cells x 2000 genes IN TOTAL,
500 control genes, k = 20, 20 replicate groups` -- the cells are split over the N ranks (STRONG scaling: the configuration the
metric names), group sums and the 2000 x 2000 Gram are all-reduced by the library's own NCCL communicator.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells M] [--scaling strong|weak] ...
    torchrun --nproc-per-node N bench.py --gpus N ...        (N > 1)

Prints ONE JSON line (rank 0):
  value        device-resident cells/s (inputs in HBM), max over ranks, CUDA events
  e2e          the same call with pinned HOST buffers (H2D of Y and D2H of newY inside the timed region); e2e_pageable: with
               ordinary (pageable) numpy arrays, which is what an R caller's REAL() vectors are
  roofline     the dominant kernel (Gram SYRK, FP64 tensor pipe) timed with CUDA events inside the timed steps
  parity       BASELINE config 2 (50k x 2000, c = 500, k = 20) sharded over the same N ranks through the same communicator,
               checked against the CPU oracle outside the timed region; a failure makes the run exit non-zero
  extra        weak (1M cells per GPU, N > 1), c4 (BASELINE config 4: 4M x 3000 over N >= 2), k_sweep (config 5: k = 5/10/20/50),
               scmerge2 (config 3: sparse 1M cells) -- builder-side context, each with its own stage times
  cpu_baseline the NumPy oracle (port of the reference's R path) on a bounded sample on this box's host cores (N = 1)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = None
_BLAS_DESC = None  # BLAS / OpenMP pools of the CPU arms (threadpoolctl), reported in cpu_baseline


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--cells", type=int, default=1_000_000, help="cells IN TOTAL (strong scaling, default) or per GPU (--scaling weak)")
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    p.add_argument("--genes", type=int, default=2000)
    p.add_argument("--ctl", type=int, default=500)
    p.add_argument("--k", type=int, default=20)
    p.add_argument("--types", type=int, default=20)
    p.add_argument("--batches", type=int, default=4)
    p.add_argument("--cpu-cells", type=int, default=50_000,
                   help="cells in the bounded CPU sample (default 50000 = BASELINE.json configs[1]; the exact path is linear in cells)")
    p.add_argument("--workload", default="fastruviii", choices=["fastruviii", "scmerge2"],
                   help="fastruviii = the BASELINE.json metric workload (default); scmerge2 = configs[2] as the headline of this run")
    p.add_argument("--density", type=float, default=0.1, help="scmerge2: fraction of non-zero entries of the sparse expression matrix")
    p.add_argument("--dense-input", action="store_true", help="scmerge2 workload: dense device matrix instead of the sparse one")
    p.add_argument("--extras", default="all", help="comma list of weak,c4,k_sweep,scmerge2,pageable,cpu200k,out_of_core,kmeans,adjust_only | all | none")
    p.add_argument("--ooc-cells", type=int, default=2_000_000, help="out_of_core extra: cells of the host matrix (x 3000 genes; 4000000 = BASELINE configs[3] on ONE GPU, needs ~100 GB of host RAM)")
    p.add_argument("--transport", default="nccl", choices=["nccl", "hook"], help="cross-rank sums: the library's NCCL communicator or the torch hook")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-parity", action="store_true")
    a = p.parse_args()
    ex = a.extras.lower()
    a.extra_set = set() if ex == "none" else ({"weak", "c4", "k_sweep", "scmerge2", "pageable", "cpu200k", "out_of_core", "kmeans", "adjust_only"} if ex == "all" else set(ex.split(",")))
    return a


def workload_name(a, cells_total, world):
    if a.workload == "scmerge2":
        inp = "dense input" if a.dense_input else f"sparse (dgCMatrix) input, density {a.density}"
        return (f"scMerge2 core synthetic {cells_total} cells x {a.genes} genes, 8 batches, {a.types} cell types, k_pseudoBulk=30, "
                f"ctl=all genes, ruvK={a.k}, {inp}")
    return (f"fastRUVIII synthetic {cells_total} cells x {a.genes} genes ({a.scaling} scaling: {cells_total // world} cells per GPU), "
            f"{a.ctl} ctl genes, k={a.k}, {a.types} replicate groups, exact mode")


# ---- clocks sampler (B200_PROFILING.md recipe) ------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def summary(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ---- synthetic data: generator G1 (planted unwanted factors), generated on the device ----------------------
def make_data(a, ctx, rank, torch, m, n=None, c=None, scm2=False, kplant=None):
    """m local cells x n genes on the device (SURVEY.md 8d generator G1; gene-level parameters shared by all ranks).
    kplant: number of planted unwanted factors (default: the ruvK of the run)."""
    import scmerge_b200 as sm
    n = n or a.genes
    c = c or a.ctl
    k, T, Bt = (kplant or a.k), a.types, (8 if scm2 else a.batches)
    dev = torch.device("cuda", ctx.device)
    gs = torch.Generator(device=dev)
    gs.manual_seed(20241)  # shared across ranks: gene-level parameters
    theta = torch.randn(T, n, generator=gs, device=dev, dtype=torch.float64)
    theta[:, :c] = 0
    scale = torch.linspace(3.0, 1.5, k, device=dev, dtype=torch.float64)
    A = torch.randn(k, n, generator=gs, device=dev, dtype=torch.float64) * scale[:, None]
    offs = torch.randn(n, generator=gs, device=dev, dtype=torch.float64) * 2.0
    gc = torch.Generator(device=dev)
    gc.manual_seed(777 + rank)  # per-rank: cell-level draws
    types = torch.randint(0, T, (m,), generator=gc, device=dev)
    batch = torch.randint(0, Bt, (m,), generator=gc, device=dev)
    dY = sm.DeviceMatrix(ctx, m, n)
    Yt = torch.as_tensor(dY.buf, device=dev).view(m, dY.ld)
    nb = min(Bt - 1, k)  # one-hot columns for all batches but the last (full column rank after centring)
    # scmerge2 workload: the unwanted factors are (batch x cell type) effects, so they survive pseudobulk averaging
    Rbt = torch.randn(Bt, T, k, generator=gs, device=dev, dtype=torch.float64)
    for r0 in range(0, m, 100_000):
        r1 = min(m, r0 + 100_000)
        b = batch[r0:r1]
        if scm2:
            Wb = Rbt[b, types[r0:r1]]
        else:
            Wb = torch.zeros(r1 - r0, k, device=dev, dtype=torch.float64)
            hit = torch.nonzero(b < nb).squeeze(1)
            Wb[hit, b[hit]] = 1.0
            if k > nb:
                Wb[:, nb:] = torch.randn(r1 - r0, k - nb, generator=gc, device=dev, dtype=torch.float64)
        Yt[r0:r1, :n] = offs + theta[types[r0:r1]] + Wb @ A + 0.5 * torch.randn(r1 - r0, n, generator=gc, device=dev,
                                                                               dtype=torch.float64)
    torch.cuda.synchronize()
    ctl = np.zeros(n, dtype=bool)
    ctl[:c] = True
    return dY, types.to(torch.int32).cpu().numpy(), ctl, batch.to(torch.int32).cpu().numpy()


def sparsify(a, ctx, dY, rank, torch):
    """scmerge2 workload: Bernoulli dropout mask on the dense generator output (single-cell matrices are ~90 % zeros; the
    low-rank structure survives in expectation), then the three dgCMatrix slots of the cells-compressed matrix in
    page-locked host memory.  Returns a namespace with format / shape / indptr / indices / data (what scMerge2_core takes)."""
    from types import SimpleNamespace

    import scmerge_b200 as sm
    m, n = dY.m, dY.n
    dev = torch.device("cuda", ctx.device)
    g = torch.Generator(device=dev)
    g.manual_seed(4242 + rank)
    Yt = torch.as_tensor(dY.buf, device=dev).view(m, dY.ld)
    counts = torch.empty(m, dtype=torch.int64, device=dev)
    step = 50_000
    for r0 in range(0, m, step):
        r1 = min(m, r0 + step)
        blk = Yt[r0:r1, :n]
        blk *= (torch.rand(r1 - r0, n, generator=g, device=dev) < a.density)
        counts[r0:r1] = (blk != 0).sum(dim=1)
    indptr = np.zeros(m + 1, dtype=np.int64)
    indptr[1:] = torch.cumsum(counts, 0).cpu().numpy()
    nnz = int(indptr[-1])
    idx, val = sm.pinned_empty((nnz,), np.int32), sm.pinned_empty((nnz,), np.float64)
    for r0 in range(0, m, step):
        r1 = min(m, r0 + step)
        blk = Yt[r0:r1, :n]
        nzr, nzc = torch.nonzero(blk, as_tuple=True)  # row-major order: sorted by cell, then gene
        p0, p1 = int(indptr[r0]), int(indptr[r1])
        idx[p0:p1] = nzc.to(torch.int32).cpu().numpy()
        val[p0:p1] = blk[nzr, nzc].cpu().numpy()
    torch.cuda.synchronize()
    return SimpleNamespace(format="csr", shape=(m, n), indptr=indptr, indices=idx, data=val, nnz=nnz)


def fp64_peak(torch, dev):
    """cuBLAS DGEMM 8192^3 best of 5 (MEASURED_PEAKS.json has no fp64 entry; same method as its bf16 one)."""
    N = 8192
    x = torch.randn(N, N, dtype=torch.float64, device=dev)
    y = torch.randn(N, N, dtype=torch.float64, device=dev)
    z = torch.empty_like(x)
    torch.matmul(x, y, out=z)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(x, y, out=z); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del x, y, z
    torch.cuda.empty_cache()
    return 2 * N ** 3 / (best * 1e-3) * 1e-12


def pcie_probe(sm, ctx, pinned):
    """Measured pinned-host <-> device copy bandwidth (GB/s) of this box: the bound of the end-to-end number."""
    import ctypes as C
    nb = min(pinned.nbytes, 2 << 30)
    d = ctx.empty((nb // 8,))
    lib = sm.lib()
    hp = pinned.ctypes.data_as(C.c_void_p)
    out = {}
    for name, fn in (("h2d_gbs", lambda: lib.scm_h2d(ctx.handle, d.ptr, hp, nb)), ("d2h_gbs", lambda: lib.scm_d2h(ctx.handle, hp, d.ptr, nb))):
        fn()
        ctx.sync()
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        out[name] = nb / (time.perf_counter() - t0) * 1e-9
    d.free()
    return out


# ---- CPU arms ---------------------------------------------------------------------------------------------
def cpu_sample(a, cells):
    from oracle import ruv_oracle as orc
    nb = 8 if a.workload == "scmerge2" else a.batches
    P = orc.planted_factors(cells, a.genes, a.ctl, a.k, n_types=a.types, n_batch=nb, seed=20241)
    if a.workload == "scmerge2":
        from scmerge_b200.scmerge2 import pseudobulk_keys  # host-side index construction only (no device, no arithmetic)
        keys, _npb, names, _t, _b = pseudobulk_keys(P["batch"], P["types"], 30, seed=1)
        P["which"] = np.array([int(nm.rsplit("_", 1)[1]) for nm in names])[keys]
        P["Yt"] = np.ascontiguousarray(P["Y"].T)
    return orc, P


def cpu_time_once(a, orc, P):
    t0 = time.perf_counter()
    if a.workload == "scmerge2":
        orc.scmerge2_core(P["Yt"], P["batch"], P["types"], P["which"], ruvK=a.k, k_fold=30, exact=False,
                          rng=np.random.default_rng(1))
    else:
        orc.fastRUVIII(P["Y"], P["types"], P["ctl"], a.k)
    return time.perf_counter() - t0


def cpu_what(a):
    if a.workload == "scmerge2":
        return ("NumPy/LAPACK restatement of scMerge2's arithmetic core (R/scMerge2.R:193-309: cosineNorm, pseudobulk means, "
                "pseudoRUVIII with rsvd, adjust of every cell); R is not installable here")
    return "NumPy/LAPACK literal restatement of R/fastRUVIII.R:41-129 (dense residop + thin dgesdd + adjust); R is not installable here"


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms are meant to use every host core, so the BLAS / OpenMP
    pools are re-opened to the full core count at run time.  Returns the number of threads actually in use."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        info = [p for p in threadpool_info() if p.get("user_api") in ("blas", "openmp")]
        global _BLAS_DESC
        _BLAS_DESC = "; ".join(f"{p.get('internal_api')} {p.get('version') or ''} x{p.get('num_threads')}".strip() for p in info) or None
        got = [p.get("num_threads", 1) for p in info]
        return max(got) if got else n
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", n))


def sample_workload(a):
    """What the CPU arms actually run (they never run the 1M-cell workload: a thin dgesdd of 1M x 2000 needs > 48 GB of work
    space and ~3 minutes per pass)."""
    if a.workload == "scmerge2":
        return f"scMerge2 core synthetic {a.cpu_cells} cells x {a.genes} genes (bounded sample of the same generator/config), ruvK={a.k}"
    return (f"fastRUVIII synthetic {a.cpu_cells} cells x {a.genes} genes, {a.ctl} ctl genes, k={a.k}, {a.types} replicate groups, exact mode "
            f"(= BASELINE.json configs[1]; bounded sample of the GPU arm's 1M-cell workload, same generator)")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = use_all_host_threads()
    orc, P = cpu_sample(a, a.cpu_cells)
    for _ in range(min(a.warmup, 1)):
        cpu_time_once(a, orc, P)
    times = [cpu_time_once(a, orc, P) for _ in range(max(1, a.steps))]
    t = float(np.mean(times))
    v = a.cpu_cells / t
    sample = f"{a.cpu_cells} cells x {a.genes} genes per step: every timed step is one pass over this sample, not over 1M cells"
    out = {"impl": "reference", "metric": "RUV-III cells/sec", "value": v, "unit": "cells/s", "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": sample_workload(a), "cells_timed_per_step": a.cpu_cells,
                      "gpu_arm_workload": workload_name(a, a.cells if a.scaling == "strong" else a.cells * a.gpus, a.gpus),
                      "note": "cells/s of the exact path is flat in the number of cells (O(m n^2) work, O(m) cells): see the 200k point "
                              "in the GPU arm's cpu_baseline.linearity"},
           "cpu_baseline": {"value": v, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample,
                            "what": cpu_what(a), "blas": _BLAS_DESC},
           "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


# ---- roofline helpers -------------------------------------------------------------------------------------------
def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


TRAFFIC_KERNEL_FILES = ("common.cuh", "seg_colsum.cuh", "fit_core.cuh", "gram_tma.cuh", "gram_syrk.cuh", "adjust.cuh", "adjust_tma.cuh")


def kernel_source_hash():
    """sha1 over the sources of the kernels the traffic figure is about (K1 seg_partial_kernel, K3 gram_tma_kernel, K6
    adjust_tma_kernel, their launch code and common.cuh): an ncu traffic figure is only attached to the roofline when it was
    captured for THIS code.  Host glue (scm_api.cu) and the other kernels do not enter."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "scmerge_b200", "csrc")
    for f in TRAFFIC_KERNEL_FILES:
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:12]


def ncu_traffic(m, n):
    """dram bytes per launch from the committed `ncu --set full` capture -- only if it was taken for this shape AND these sources."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}, {"file": None}
    src = {"file": "profiles/ncu_traffic.json", "captured_for_sources": tj.get("kernel_source_hash"), "commit": tj.get("commit"),
           "current_sources": kernel_source_hash()}
    ok = tj.get("cells") == m and tj.get("genes") == n and tj.get("kernel_source_hash") == src["current_sources"]
    src["valid"] = bool(ok)
    return (tj.get("dram_bytes_per_launch", {}) if ok else {}), src


def hbm_sections(stages, m, n, hbm, fused_copy):
    sec = {}
    if stages.get("seg_colsum", 0) > 0:
        by = (16.0 if fused_copy else 8.0) * m * n  # the summing pass also WRITES the pre-shifted copy (8 m n) when fused
        gb = by / (stages["seg_colsum"] * 1e-3) * 1e-9
        sec["seg_colsum"] = {"bound": "hbm", "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": gb / hbm, "ms": stages["seg_colsum"],
                             "algorithmic": "8*m*n read + 8*m*n written (sums + pre-shifted copy in one pass)" if fused_copy else "8*m*n read"}
    if stages.get("preshift_copy", 0) > 0:
        gb = 16.0 * m * n / (stages["preshift_copy"] * 1e-3) * 1e-9
        sec["preshift_copy"] = {"bound": "hbm", "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": gb / hbm, "ms": stages["preshift_copy"]}
    if stages.get("adjust", 0) > 0:
        gb = 16.0 * m * n / (stages["adjust"] * 1e-3) * 1e-9
        sec["adjust_rank_k"] = {"bound": "hbm", "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": gb / hbm, "ms": stages["adjust"],
                                "algorithmic": "8*m*n read + 8*m*n written"}
    return sec


# ---- our arm ----------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    import scmerge_b200 as sm
    from scmerge_b200.dist import init_builtin_nccl, shard_rows, torch_allreduce_hook
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world == 1 and a.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None if os.environ.get("SCM_NO_NUMA_BIND") else sm.bind_host_to_gpu(local)  # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = sm.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    if world > 1:
        if a.transport == "nccl":
            init_builtin_nccl(ctx, rank, world)
        else:
            ctx.set_allreduce(torch_allreduce_hook(local), world, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxf(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumf(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(step, steps, warmup):
        """W warm-up + K timed steps between barriers; returns (max-over-ranks ms total, mean stage times, launches)."""
        for _ in range(warmup):
            step()
        barrier()
        l0 = ctx.launch_count()
        acc = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
            for kname, v in ctx.timers().items():
                acc[kname] = acc.get(kname, 0.0) + v
        e1.record()
        barrier()
        return maxf(e0.elapsed_time(e1)), {k2: v / steps for k2, v in acc.items()}, ctx.launch_count() - l0

    peaks = measured_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6.65 TB/s"
    fused_copy = os.environ.get("SCM_K1_FUSED_COPY", "1") != "0" and os.environ.get("SCM_GRAM_PRESHIFT", "1") != "0"
    failures = []

    # ---- parity block: BASELINE config 2 through the same ranks / communicator, against the CPU oracle (untimed) ----------
    parity = None
    if not a.no_parity and a.workload == "fastruviii":
        from oracle import ruv_oracle as orc
        pm, pn, pc, pk = 50_000, 2000, 500, 20
        P = orc.planted_factors(pm, pn, pc, pk, n_types=20, n_batch=4, seed=20242)  # same seed on every rank: identical data
        lo, hi = shard_rows(pm, world, rank)
        out = sm.fastRUVIII(np.ascontiguousarray(P["Y"][lo:hi]), P["types"][lo:hi], P["ctl"], k=pk, return_info=True, ctx=ctx)
        ref_new = torch.empty((pm, pn), dtype=torch.float64, device=dev)
        ref_fa = torch.empty((pk, pn), dtype=torch.float64, device=dev)
        if rank == 0:
            use_all_host_threads()
            t0 = time.perf_counter()
            ref = orc.fastRUVIII_reduced(P["Y"], P["types"], P["ctl"], pk)
            log(f"[parity] oracle (reduced form, rank 0) {time.perf_counter() - t0:.1f} s")
            ref_new.copy_(torch.from_numpy(ref["newY"]))
            ref_fa.copy_(torch.from_numpy(np.ascontiguousarray(ref["fullalpha"][:pk])))
        if world > 1:
            dist.broadcast(ref_new, 0)
            dist.broadcast(ref_fa, 0)
        mine = torch.from_numpy(out["newY"]).to(dev)
        num = sumf(float(((mine - ref_new[lo:hi]) ** 2).sum().item()))
        den = sumf(float((ref_new[lo:hi] ** 2).sum().item()))
        fa = np.array(out["fullalpha"])
        ang = float(orc.principal_angles(fa[:pk], ref_fa.cpu().numpy()).max())
        digest = int.from_bytes(hashlib.sha1(fa.tobytes()).digest()[:6], "big")
        same = True
        if world > 1:
            t = torch.tensor([digest], device=dev, dtype=torch.int64)
            lst = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            same = all(int(x.item()) == int(lst[0].item()) for x in lst)
        rel = (num / den) ** 0.5
        parity = {"config": f"BASELINE configs[1]: fastRUVIII {pm} x {pn}, {pc} ctl, k={pk}, cells sharded over {world} rank(s), "
                            f"transport {ctx.transport()}", "oracle": "oracle.ruv_oracle.fastRUVIII_reduced (R/fastRUVIII.R:41-111 in Gram form; "
                            "equal to the literal dgesdd form to 1e-12, tests/test_oracle.py)",
                  "rel_frobenius": rel, "max_angle": ang, "identical_fullalpha_across_ranks": same,
                  "tolerance": {"rel_frobenius": 1e-8, "max_angle": 1e-6}, "ok": bool(rel <= 1e-8 and ang <= 1e-6 and same)}
        if not parity["ok"]:
            failures.append("parity")
        del ref_new, ref_fa, mine, out, P
        ctx.release_scratch()
        torch.cuda.empty_cache()

    # ---- headline ------------------------------------------------------------------------------------------------------
    if a.scaling == "strong":
        lo, hi = shard_rows(a.cells, world, rank)
        m_local, cells_total = hi - lo, a.cells
    else:
        m_local, cells_total = a.cells, a.cells * world
    n = a.genes
    scm2 = a.workload == "scmerge2"
    dY, keys, ctl, batch_np = make_data(a, ctx, rank, torch, m_local, scm2=scm2)
    dOut = sm.DeviceMatrix(ctx, m_local, n)
    peak64 = fp64_peak(torch, dev) if rank == 0 else 0.0
    Xh = pb = None
    if scm2:
        import warnings
        warnings.simplefilter("ignore")
        if world > 1:
            raise SystemExit("the scmerge2 workload is a single-GPU line (replicas only): run it with --gpus 1")
        pb = sm.pseudobulk_keys(batch_np, keys, 30, seed=1)
        if a.dense_input:
            a.no_e2e = True
            src = dY
        else:
            # the matrix scMerge2 actually holds: sparse, compressed by cell (R/scMerge2.R:164-166)
            Xh = sparsify(a, ctx, dY, rank, torch)
            dY.free()
            src = sm.DeviceCSC(ctx, Xh.indptr, Xh.indices, Xh.data, m_local, n)

        def step():
            return sm.scMerge2_core(src, None, None, ctl=None, ruvK=a.k, pb=pb, return_bulk=False, ctx=ctx, out=dOut)
    else:
        def step():
            return sm.fastRUVIII(dY, keys, ctl, k=a.k, out=dOut, ctx=ctx)

    for _ in range(a.warmup):
        step()
    barrier()
    clocks = Clocks(local) if rank == 0 else None
    time.sleep(0.3)
    tw0 = time.time()
    ms_max, stages, launches = timed(step, a.steps, 0)
    tw1 = time.time()
    clk = clocks.summary(tw0, tw1) if clocks else None
    fit_info = ctx.fit_info()

    # ---- end to end: pinned host input, pinned host newY out, through the same public call -------------------------------
    e2e = e2e_pageable = None
    if not a.no_e2e:
        n_e2e = max(1, min(a.steps, 3))
        nbytes = m_local * n * 8
        Oh = sm.pinned_empty((m_local, n))
        if scm2:
            src.free()
            dOut.free()

            def call():
                sm.scMerge2_core(Xh, None, None, ctl=None, ruvK=a.k, pb=pb, return_bulk=False, ctx=ctx, out=Oh)
            h2d = Xh.nnz * 12 + (m_local + 1) * 8 + 4 * m_local
            d2h = nbytes + 8 * a.k * n
            api = "scmerge_b200.scMerge2_core(sparse cells-compressed host matrix (pinned), pb keys, ruvK, out=numpy pinned)"
        else:
            Yh = sm.pinned_empty((m_local, n))
            dY.to_host(Yh)
            dOut.free()
            dY.free()
            ctx.release_scratch()

            def call():
                sm.fastRUVIII(Yh, keys, ctl, k=a.k, out=Oh, ctx=ctx)
            h2d = nbytes + 4 * m_local
            d2h = nbytes + 8 * 50 * n
            api = "scmerge_b200.fastRUVIII(numpy pinned Y, keys, ctl, k, out=numpy pinned) -> scm_ruv3_host"
        call()  # warm-up
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(n_e2e):
            call()
        f1.record()
        barrier()
        t2 = maxf(f0.elapsed_time(f1))
        pcie = pcie_probe(sm, ctx, Oh)
        bound_s = h2d / (pcie["h2d_gbs"] * 1e9) + d2h / (pcie["d2h_gbs"] * 1e9)  # the fit needs every cell before the first adjusted byte exists
        e2e = {"value": cells_total * n_e2e / (t2 * 1e-3), "unit": "cells/s", "steps": n_e2e, "ms_per_step": t2 / n_e2e,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "bytes_are": "summed over the ranks",
               "api": api, "pcie": pcie, "host_cpus_bound_to": numa,
               "pcie_bound_ms_per_step": bound_s * 1e3, "pcie_bound_cells_per_s": cells_total / bound_s}
        if "pageable" in a.extra_set and not scm2 and world == 1:
            # what a caller WITHOUT page-locked buffers gets (an R session's REAL(Y) is ordinary pageable memory)
            try:
                Yp, Op = np.array(Yh), np.empty((m_local, n))
                sm.fastRUVIII(Yp, keys, ctl, k=a.k, out=Op, ctx=ctx)
                t0 = time.perf_counter()
                sm.fastRUVIII(Yp, keys, ctl, k=a.k, out=Op, ctx=ctx)
                tp = time.perf_counter() - t0
                e2e_pageable = {"value": m_local / tp, "unit": "cells/s", "ms_per_step": tp * 1e3, "steps": 1,
                                "api": "scmerge_b200.fastRUVIII(ordinary numpy Y, ..., out=ordinary numpy): same scm_ruv3_host call, pageable buffers"}
                del Yp, Op
            except Exception as exc:  # noqa: BLE001
                e2e_pageable = {"error": repr(exc)}
        del Oh
        if not scm2:
            del Yh
    else:
        dOut.free()
        if not scm2 or a.dense_input:
            dY.free()
        elif Xh is not None:
            src.free()
    ctx.release_scratch()
    torch.cuda.empty_cache()

    # ---- extras (builder-side context; never the headline) ----------------------------------------------------------------
    extra = {}

    def guarded(name, fn):
        if name not in a.extra_set:
            return
        try:
            t0 = time.perf_counter()
            extra[name] = fn()
            if isinstance(extra[name], dict):
                extra[name]["wall_s_incl_setup"] = round(time.perf_counter() - t0, 1)
        except Exception as exc:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            extra[name] = {"error": repr(exc)}
        ctx.release_scratch()
        torch.cuda.empty_cache()
        barrier()

    def fast_line(mloc, ngenes, nctl, kk, steps=3, warmup=2, inplace=False):
        """One fastRUVIII line on mloc local cells: returns ms/step (max over ranks), stages, sections."""
        dYx, kx, cx, _ = make_data(a, ctx, rank, torch, mloc, n=ngenes, c=nctl)
        dOx = dYx if inplace else sm.DeviceMatrix(ctx, mloc, ngenes)
        ms, st, _ = timed(lambda: sm.fastRUVIII(dYx, kx, cx, k=kk, out=dOx, ctx=ctx), steps, warmup)
        info = ctx.fit_info()
        dYx.free()
        if not inplace:
            dOx.free()
        return ms / steps, st, info

    if a.workload == "fastruviii":
        def x_weak():
            ms, st, _ = fast_line(1_000_000, n, a.ctl, a.k)
            return {"workload": f"fastRUVIII 1M cells PER GPU x {n} genes (weak scaling, {world} GPU(s))", "ms_per_step": ms,
                    "value": 1_000_000 * world / (ms * 1e-3), "unit": "cells/s", "stages_ms": st}
        if world > 1:
            guarded("weak", x_weak)

        def x_c4():
            tot, g4, c4 = 4_000_000, 3000, 500
            lo4, hi4 = shard_rows(tot, world, rank)
            ms, st, _ = fast_line(hi4 - lo4, g4, c4, 20, steps=2, warmup=1, inplace=True)
            syrk = st.get("gram_syrk_kernel", 0.0)
            fl = (hi4 - lo4) * g4 * (g4 + 1.0)
            return {"workload": f"BASELINE configs[3]: fastRUVIII {tot} cells x {g4} genes, {c4} ctl, k=20, cell-sharded over {world} B200 "
                                f"({hi4 - lo4} cells per GPU), NCCL Gram all-reduce ({g4 * g4 * 8 / 1e6:.0f} MB), adjust in place",
                    "ms_per_step": ms, "value": tot / (ms * 1e-3), "unit": "cells/s", "stages_ms": st,
                    "gram_tflops_per_gpu": fl / (syrk * 1e-3) * 1e-12 if syrk > 0 else None}
        if world >= 2:
            guarded("c4", x_c4)

        def x_ksweep():
            res = {}
            lo5, hi5 = shard_rows(1_000_000, world, rank)
            # 50 planted factors: every k of the sweep asks for a subspace the data defines (with 20 planted factors k = 50 would
            # ask for 30 directions of the flat noise bulk, which no iteration -- and no user -- can separate)
            dYx, kx, cx, _ = make_data(a, ctx, rank, torch, hi5 - lo5, kplant=50)
            dOx = sm.DeviceMatrix(ctx, hi5 - lo5, n)
            fl_per_k = lambda kk: 2.0 * (hi5 - lo5) * kk * (a.ctl + n)  # noqa: E731  flop of the two products of the rank-k update
            for kk in (5, 10, 20, 50):
                ms, st, _ = timed(lambda: sm.fastRUVIII(dYx, kx, cx, k=kk, out=dOx, ctx=ctx), 3, 2)
                info = ctx.fit_info()
                adj = st.get("adjust", 0.0)
                res[f"k={kk}"] = {"ms_per_step": ms / 3, "cells_per_s": 1_000_000 / (ms / 3 * 1e-3), "eig_ms": st.get("eig"), "coef_ms": st.get("coef"),
                                  "adjust_ms": adj, "adjust_hbm_frac": (16.0 * (hi5 - lo5) * n / (adj * 1e-3) * 1e-9 / hbm) if adj > 0 else None,
                                  "adjust_dmma_tflops": fl_per_k(kk) / (adj * 1e-3) * 1e-12 if adj > 0 else None,
                                  "gram_ms": st.get("gram"), "seg_colsum_ms": st.get("seg_colsum"), "rr_steps": info["iters"],
                                  "conv_rank": info["conv_rank"]}
            dYx.free()
            dOx.free()
            return {"workload": f"BASELINE configs[4]: ruvK sweep at 1M cells x {n} genes over {world} GPU(s), 50 planted factors, s = 50 rows of fullalpha",
                    "note": "the rank-k update is HBM-bound up to k ~ 30 (2 k (c + n) flop per 16 n bytes) and FP64-tensor-bound above: "
                            "read adjust_hbm_frac for k <= 20 and adjust_dmma_tflops (against the fp64 peak of the roofline block) for k = 50", **res}
        guarded("k_sweep", x_ksweep)

        def x_scm2():
            import warnings
            warnings.simplefilter("ignore")
            dYx, kx, _cx, bx = make_data(a, ctx, rank, torch, 1_000_000, scm2=True)
            pbx = sm.pseudobulk_keys(bx, kx, 30, seed=1)
            Xs = sparsify(a, ctx, dYx, rank, torch)
            dYx.free()
            srcx = sm.DeviceCSC(ctx, Xs.indptr, Xs.indices, Xs.data, 1_000_000, n)
            dOx = sm.DeviceMatrix(ctx, 1_000_000, n)
            ms, st, _ = timed(lambda: sm.scMerge2_core(srcx, None, None, ctl=None, ruvK=a.k, pb=pbx, return_bulk=False, ctx=ctx, out=dOx), 3, 2)
            fa_rows = int(min(pbx[1] - a.types, n))  # scMerge2's default RandomParam: svd_k = min(P - ncol(M), sum(ctl))
            # the same call with max(ruvK, 50) converged rows of fullalpha instead of all of them (adjusted cells identical)
            import scmerge_b200.ruv as ruv_mod
            ruv_mod.FULLALPHA_ROWS = "truncate"
            try:
                ms_t, st_t, _ = timed(lambda: sm.scMerge2_core(srcx, None, None, ctl=None, ruvK=a.k, pb=pbx, return_bulk=False, ctx=ctx, out=dOx), 3, 2)
            finally:
                ruv_mod.FULLALPHA_ROWS = "reference"
            adj = st.get("adjust", 0.0)
            by = 12.0 * Xs.nnz + 8.0 * 1_000_000 * n
            srcx.free()
            dOx.free()
            return {"workload": f"BASELINE configs[2]: scMerge2 core, sparse 1M cells x {n} genes (density {a.density}), 8 batches, {a.types} cell types, "
                                f"{pbx[1]} pseudobulks, ruvK={a.k}, device-resident sparse input", "ms_per_step": ms / 3,
                    "value": 1_000_000 / (ms / 3 * 1e-3), "unit": "cells/s", "stages_ms": st, "fullalpha_rows": fa_rows,
                    "fullalpha_rows_note": "all rows the reference's rule asks for (whole spectrum of the Gram, eig_full.cuh)",
                    "truncated_rows": {"fullalpha_rows": 50, "ms_per_step": ms_t / 3, "value": 1_000_000 / (ms_t / 3 * 1e-3), "stages_ms": st_t,
                                       "note": "scmerge_b200.ruv.FULLALPHA_ROWS = 'truncate': subspace iteration, same adjusted cells"},
                    "adjust_hbm_frac": by / (adj * 1e-3) * 1e-9 / hbm if adj > 0 else None, "adjust_algorithmic": "12*nnz read + 8*m*n written"}
        if world == 1:
            guarded("scmerge2", x_scm2)

        def x_ooc():
            """BASELINE configs[3]'s shape on ONE GPU through the host call: the matrix lives in (pageable) host memory, is
            streamed twice through three chunk buffers and adjusted in place on the host."""
            mo, go, co = a.ooc_cells, 3000, 500
            need = mo * go * 8
            avail = 0
            for ln in open("/proc/meminfo"):
                if ln.startswith("MemAvailable"):
                    avail = int(ln.split()[1]) * 1024
            if avail < 1.6 * need + (8 << 30):
                return {"skipped": f"host has {avail / 2**30:.0f} GiB available, the {need / 2**30:.0f} GiB matrix needs head-room"}
            Yh = np.empty((mo, go))
            keys_o = np.empty(mo, dtype=np.int32)
            ctl_o = None
            step_rows = 200_000
            for r0 in range(0, mo, step_rows):  # generated on the device block by block (same generator, block index as the "rank")
                r1 = min(mo, r0 + step_rows)
                dYb, kb, ctl_o, _ = make_data(a, ctx, 1000 + r0 // step_rows, torch, r1 - r0, n=go, c=co)
                dYb.to_host(Yh[r0:r1])
                keys_o[r0:r1] = kb
                dYb.free()
            ctx.release_scratch()
            os.environ["SCM_HOST_STREAM"] = "1"
            try:
                t0 = time.perf_counter()
                res = sm.fastRUVIII(Yh, keys_o, ctl_o, k=20, return_info=True, ctx=ctx, out=Yh)
                t1 = time.perf_counter() - t0
                t0 = time.perf_counter()
                sm.fastRUVIII(Yh, keys_o, ctl_o, k=20, ctx=ctx, out=Yh, fullalpha=res["fullalpha"])  # second call: warm bounce buffers
                t2 = time.perf_counter() - t0
            finally:
                os.environ.pop("SCM_HOST_STREAM", None)
            return {"workload": f"fastRUVIII {mo} cells x {go} genes ({need / 1e9:.0f} GB) held in PAGEABLE host memory, one B200, out-of-core form "
                                f"(3 chunk buffers on the device, matrix streamed twice, result written in place on the host)",
                    "fit_and_adjust_s": t1, "cells_per_s": mo / t1, "pcie_bytes": 3 * need, "effective_gb_s": 3 * need / t1 * 1e-9,
                    "adjust_only_s": t2, "adjust_only_effective_gb_s": 2 * need / t2 * 1e-9, "fit": ctx.fit_info(),
                    "conv_rank": int(res["fullalpha"].conv_rank), "device_bytes_for_cells": 3 * 512 * 2**20}
        if world == 1:
            guarded("out_of_core", x_ooc)

        def x_adjust_only():
            """The second half of the reference's fit / adjust checkpoint through the public call with HOST buffers (pinned):
            fastRUVIII(Y, fullalpha = stored) -- nothing has to be known about the cells first, so the host driver streams every
            chunk up, adjusts it and sends it home in one full-duplex pass."""
            mo = 1_000_000
            dYx, kx, cx, _ = make_data(a, ctx, rank, torch, mo, n=n, c=a.ctl)
            fit = sm.fastRUVIII(dYx, kx, cx, k=a.k, return_info=True, ctx=ctx)
            fit["newY"].free()
            Yh, Oh = sm.pinned_empty((mo, n)), sm.pinned_empty((mo, n))
            dYx.to_host(Yh)
            dYx.free()
            ctx.release_scratch()
            sm.fastRUVIII(Yh, kx, cx, k=a.k, fullalpha=fit["fullalpha"], ctx=ctx, out=Oh)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                sm.fastRUVIII(Yh, kx, cx, k=a.k, fullalpha=fit["fullalpha"], ctx=ctx, out=Oh)
                ts.append(time.perf_counter() - t0)
            t = min(ts)
            return {"workload": f"fastRUVIII(Y, fullalpha = stored) {mo} cells x {n} genes, pinned host in / out, one pass, uploads and downloads concurrent",
                    "ms": 1e3 * t, "cells_per_s": mo / t, "h2d_bytes": mo * n * 8, "d2h_bytes": mo * n * 8,
                    "pcie_gb_s_each_way": mo * n * 8 / t * 1e-9}
        if world == 1:
            guarded("adjust_only", x_adjust_only)

        def x_kmeans():
            """identifyCluster's kmeans (R/scReplicate.R:394-395): 1000 starts on the first 10 PCs of one batch, on the device."""
            from scmerge_b200 import replicate as rp
            rng = np.random.default_rng(0)
            res = {}
            for mb, K in ((50_000, 3), (200_000, 8)):
                cent = rng.normal(size=(K, 10)) * 2.0
                x = cent[rng.integers(0, K, mb)] + rng.normal(size=(mb, 10))
                dX = sm.DeviceMatrix.from_host(ctx, x)
                rp.kmeans(dX, K, nstart=10, ctx=ctx)
                ctx.sync()
                t0 = time.perf_counter()
                r = rp.kmeans(dX, K, nstart=1000, ctx=ctx)
                ctx.sync()
                res[f"{mb}x10_K{K}_nstart1000"] = {"ms": 1e3 * (time.perf_counter() - t0), "tot_withinss": r["tot_withinss"], "lloyd_iterations_of_winner": r["iter"],
                                                   "hartigan_wong_transfers": r["transfers"]}
                dX.free()
            return res
        if world == 1:
            guarded("kmeans", x_kmeans)

    if rank == 0:
        m = m_local
        syrk_ms = stages.get("gram_syrk_kernel", 0.0)
        flops = m * n * (n + 1.0)
        ach = flops / (syrk_ms * 1e-3) * 1e-12 if syrk_ms > 0 else 0.0
        sec = hbm_sections(stages, m, n, hbm, fused_copy)
        traffic, traffic_src = ncu_traffic(cells_total // world if a.scaling == "strong" else a.cells, n)
        preshifted = os.environ.get("SCM_GRAM_PRESHIFT", "1") != "0"
        gram_roof = {"kernel": "gram_tma_kernel<SHIFT=false> (FP64 DMMA SYRK, TMA-staged, on the pre-shifted copy of the cells)" if preshifted
                     else "gram_tma_kernel<SHIFT=true> (FP64 DMMA SYRK, TMA-staged)", "bound": "tensor", "achieved": ach, "peak": peak64,
                     "unit": "TFLOP/s", "frac": ach / peak64 if peak64 else None, "traffic": traffic.get("gram_tma_kernel"),
                     "traffic_source": traffic_src,
                     "algorithmic": f"m_local*n*(n+1) flop per launch, m_local = {m} cells on this GPU", "ms_per_launch": syrk_ms,
                     "peak_source": "cuBLAS DGEMM 8192^3 fp64 measured in this run (MEASURED_PEAKS.json has no fp64 entry)"}
        if scm2:
            # the fit runs on the 4800 x n pseudobulk matrix: the dominant kernel of this workload is the rank-k update
            adj = sec.get("adjust_rank_k", {})
            sparse = Xh is not None
            if sparse and adj.get("ms"):  # sparse-native rank-k update: 12 bytes per non-zero read + the dense result written once
                by = 12.0 * Xh.nnz + 8.0 * m * n
                adj = dict(adj, achieved=by / (adj["ms"] * 1e-3) * 1e-9, frac=by / (adj["ms"] * 1e-3) * 1e-9 / hbm)
                sec["adjust_rank_k"] = adj
            roof = {"kernel": "csc_project_kernel + csc_adjust_red_kernel (rank-k update on sparse rows)" if sparse else "adjust_tma_kernel (fused rank-k update)",
                    "bound": "hbm", "achieved": adj.get("achieved"), "peak": hbm,
                    "unit": "GB/s", "frac": adj.get("frac"), "traffic": traffic.get("csc_adjust_red_kernel" if sparse else "adjust_tma_kernel"),
                    "traffic_source": traffic_src,
                    "algorithmic": "12*nnz read + 8*m*n written per launch" if sparse else "8*m*n read + 8*m*n written per launch",
                    "ms_per_launch": adj.get("ms"), "peak_source": hbm_src}
            sec.pop("seg_colsum", None)  # K1 here is the pseudobulk pass; it is timed inside scMerge2_core, not by the fit timers
        else:
            roof = gram_roof
        for kname, kern in (("seg_colsum", "seg_partial_kernel"), ("adjust_rank_k", "adjust_tma_kernel"), ("preshift_copy", "shift_copy_kernel")):
            if kname in sec:
                sec[kname]["traffic"] = traffic.get(kern)
        out = {"metric": "RUV-III cells/sec", "value": cells_total * a.steps / (ms_max * 1e-3), "unit": "cells/s",
               "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
               "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic (planted-factor generator G1, on device)",
               "config": {"workload": workload_name(a, cells_total, world), "global_cells": cells_total, "cells_per_gpu": m_local,
                          "l2": f"inputs ({8e-9 * m * n:.1f} GB/GPU) larger than the 126 MB L2",
                          "parallelism": f"cells sharded over {world} GPU(s); 2 all-reduces per fit ([group sums|counts|flag], [{n}x{n} Gram]) "
                                         f"through {ctx.transport() if world > 1 else 'no'} transport"},
               "clocks": clk, "gpu_launches": launches, "e2e": e2e, "roofline": roof, "parity": parity,
               "hbm_kernels": sec, "hbm_peak_source": hbm_src, "stages_ms": stages, "fit": fit_info, "extra": extra}
        if e2e_pageable is not None:
            out["e2e_pageable"] = e2e_pageable
        if not a.no_cpu and world == 1:
            threads = use_all_host_threads()
            orc, P = cpu_sample(a, a.cpu_cells)
            tcpu = cpu_time_once(a, orc, P)
            cb = {"value": a.cpu_cells / tcpu, "unit": "cells/s", "cores": threads, "kind": "port",
                  "sample": f"{a.cpu_cells} cells x {a.genes} genes, one pass of the NumPy oracle (bounded sample of the 1M-cell workload)",
                  "what": cpu_what(a), "blas": _BLAS_DESC, "seconds": tcpu}
            if "cpu200k" in a.extra_set and a.workload == "fastruviii":
                # a second point to back "cells/s is flat in m" (the extrapolation the ratio rests on)
                try:
                    del P
                    orc, P2 = cpu_sample(a, 200_000)
                    t2 = cpu_time_once(a, orc, P2)
                    cb["linearity"] = {"cells": [a.cpu_cells, 200_000], "seconds": [tcpu, t2],
                                       "cells_per_s": [a.cpu_cells / tcpu, 200_000 / t2]}
                    del P2
                except Exception as exc:  # noqa: BLE001
                    cb["linearity"] = {"error": repr(exc)}
            out["cpu_baseline"] = cb
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if failures:
        log("FAILED: " + ", ".join(failures))
        sys.exit(3)


def main():
    a = parse()
    # Only the JSON line may reach stdout: libraries (NCCL's version banner, ...) write to fd 1 from native code,
    # so fd 1 is pointed at stderr for the whole run and the line is written to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
