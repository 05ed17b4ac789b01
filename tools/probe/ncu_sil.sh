#!/bin/bash
set -e
CMD="python tools/probe/select_timing.py"
CELLS=100000 $CMD > gpurun_out/ncu_plain5.log 2> gpurun_out/ncu_plain5.err
CELLS=100000 ncu --set full --clock-control none --import-source on -k "regex:sil_pairsum_kernel" -c 1 -o gpurun_out/prof_r01h_sil -f $CMD > gpurun_out/ncu_h.log 2>&1
echo done
