#!/bin/bash
# DRAM traffic per launch of the hot kernels from ONE `ncu --set full` capture of a bench step (run under gpurun, one GPU), written
# to profiles/ncu_traffic.json stamped with the hash of the kernel sources it was taken for (bench.py only attaches a traffic
# figure to its roofline when that hash matches the tree it runs from).
#   bash tools/probe/ncu_traffic.sh [tag]     ->  gpurun_out/<tag>_full.ncu-rep, gpurun_out/<tag>_full_raw.csv, gpurun_out/ncu_traffic.json
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --extras none"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:gram_tma_kernel|adjust_tma_kernel|seg_partial_kernel" -c 10 \
    -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1 || { echo "ncu failed"; tail -5 gpurun_out/${TAG}_ncu.log; exit 1; }
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv
python tools/probe/ncu_traffic.py gpurun_out/${TAG}_full_raw.csv gpurun_out/ncu_traffic.json
