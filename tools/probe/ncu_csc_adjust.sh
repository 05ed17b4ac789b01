#!/bin/bash
set -e
CMD="python bench.py --workload scmerge2 --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain3.log 2> gpurun_out/ncu_plain3.err
ncu --set full --clock-control none --import-source on -k "regex:csc_adjust_kernel" -s 1 -c 1 -o gpurun_out/prof_r01e_csc -f $CMD > gpurun_out/ncu_d.log 2>&1
echo done
