#!/bin/bash
# ncu evidence for the bench workload (run under gpurun, one GPU):
#   1. plain run (must exit 0), 2. launch list with per-launch device time, 3. one `--set full` capture of the three hot kernels.
set -e
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2> gpurun_out/ncu_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01f.csv $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:gram_tma_kernel|adjust_tma_kernel|seg_partial_kernel" -s 3 -c 3 \
    -o gpurun_out/prof_r01f -f $CMD > gpurun_out/ncu_b.log 2>&1
CMD2="python bench.py --workload scmerge2 --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD2 > gpurun_out/ncu_plain2.log 2> gpurun_out/ncu_plain2.err
ncu --set full --clock-control none --import-source on -k "regex:csc_adjust_kernel|csc_seg_partial_kernel|csc_rownorm_kernel" -s 3 -c 3 -o gpurun_out/prof_r01f_csc -f $CMD2 > gpurun_out/ncu_c.log 2>&1
echo done
