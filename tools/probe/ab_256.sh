#!/bin/bash
# 32-byte global accesses (LDG/STG.E.256) against pairs of 16-byte ones: stage times of the fastRUVIII step and the sparse update
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity --extras none"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['stages_ms']
print('%s: %.3f ms/step  ' % (sys.argv[1], d['ms_per_step']) + '  '.join('%s %.3f' % (k, v) for k, v in st.items()))" "$1"; }
for r in 1 2; do
  SCM_NO_256=1 $B 2>/dev/null | show "16-byte pairs"
  SCM_NO_256=0 $B 2>/dev/null | show "32-byte     "
done
SCM_NO_256=1 $B --k 50 2>/dev/null | show "k=50 16-byte pairs"
SCM_NO_256=0 $B --k 50 2>/dev/null | show "k=50 32-byte     "
P="python tools/probe/csc_adjust_probe.py --reps 3"
SCM_NO_256=1 SCM_CSC_DEBUG=1 $P 2>&1 | grep best
SCM_NO_256=0 SCM_CSC_DEBUG=1 $P 2>&1 | grep best
