#!/bin/bash
set -e
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain4.log 2> gpurun_out/ncu_plain4.err
ncu --set full --clock-control none --import-source on -k "regex:gram_tma_kernel" -s 2 -c 1 -o gpurun_out/prof_r01g -f $CMD > gpurun_out/ncu_g.log 2>&1
echo done
