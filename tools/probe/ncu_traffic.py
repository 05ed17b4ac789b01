"""Raw-page CSV of an `ncu --set full` capture -> ncu_traffic.json: for every kernel the launch with the longest duration (the
full-size one; the same kernels also run on the small G x n / sample problems), dram bytes read + written, duration and the pipe
utilisation figures DESIGN.md quotes.  Stamped with the hash of the kernel sources (bench.kernel_source_hash)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import TRAFFIC_KERNEL_FILES, kernel_source_hash  # noqa: E402

WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_cycles_pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "%": 1.0}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units = rows[hi], rows[hi + 1]
    kn = head.index("Kernel Name")
    cols = {name: head.index(name) for name in WANT if name in head}
    best = {}
    for r in rows[hi + 2:]:
        if len(r) <= kn:
            continue
        name = r[kn]
        for short in ("gram_tma_kernel", "adjust_tma_kernel", "seg_partial_kernel", "csc_adjust_red_kernel", "csc_adjust_kernel", "shift_copy_kernel"):
            if short in name:
                vals = {}
                for full, key in WANT.items():
                    if full in cols and r[cols[full]] not in ("", "n/a"):
                        vals[key] = float(r[cols[full]].replace(",", "")) * UNIT.get(units[cols[full]], 1.0)
                if "duration" in vals and (short not in best or vals["duration"] > best[short]["duration"]):
                    vals["kernel"] = name[:100]
                    best[short] = vals
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except Exception:
        commit = None
    out = {"cells": 1_000_000, "genes": 2000, "kernel_source_hash": kernel_source_hash(), "hash_files": list(TRAFFIC_KERNEL_FILES), "commit": commit or None,
           "source": f"{os.path.basename(src)}: ncu --set full --clock-control none, bench.py --steps 1 --warmup 1 (1M x 2000, one B200); "
                     "per kernel the longest launch; dram bytes = dram__bytes_read.sum + dram__bytes_write.sum",
           "dram_bytes_per_launch": {k: v.get("dram_read", 0.0) + v.get("dram_write", 0.0) for k, v in best.items()},
           "algorithmic_bytes": {"gram_tma_kernel": 16e9, "adjust_tma_kernel": 32e9, "seg_partial_kernel": 32e9},
           "detail": best}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out["dram_bytes_per_launch"]), {k: round(v["duration"] * 1e3, 3) for k, v in best.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
