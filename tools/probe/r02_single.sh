#!/bin/bash
# One-GPU round-2 check: full GPU test suite, bench line, ncu launch list of a bench step (per-launch durations).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo bench rc=$?
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_c.json"))
print("value", d["value"], "ms", d["ms_per_step"], "stages", d["stages_ms"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pageable", d.get("e2e_pageable"))
print("parity", d["parity"]["ok"], d["parity"]["rel_frobenius"])
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_bench_1M.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --extras none > gpurun_out/ncu_bench.log 2>&1; echo ncu rc=$?
