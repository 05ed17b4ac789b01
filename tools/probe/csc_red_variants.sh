#!/bin/bash
# pass B alone (SCM_CSC_DEBUG=1: no projection) for builds with other cells-per-warp / slab widths / warps per CTA
P="python tools/probe/csc_adjust_probe.py --reps 3"
for lib in build/libscm_mt*.so scmerge_b200/libscmerge_b200.so; do
  SCM_LIB_PATH=$lib SCM_CSC_DEBUG=1 $P 2>&1 | grep "best"
done
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k "regex:csc_adjust_red_kernel|csc_project_kernel" -s 4 -c 4 -o gpurun_out/r02_csc_red2 -f \
  python tools/probe/csc_adjust_probe.py --reps 1 > gpurun_out/r02_csc_red2_ncu.log 2>&1
ncu -i gpurun_out/r02_csc_red2.ncu-rep --page raw --csv > gpurun_out/r02_csc_red2_raw.csv 2>/dev/null
echo ncu done
fi
