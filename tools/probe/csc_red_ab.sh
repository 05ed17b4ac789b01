#!/bin/bash
# A/B of the two pass-B kernels of the sparse rank-k update (same box, same data)
P="python tools/probe/csc_adjust_probe.py"
SCM_CSC_RED=0 $P --save /tmp/csc_ref.npy
SCM_CSC_RED=1 $P --check /tmp/csc_ref.npy
SCM_CSC_RED=0 SCM_CSC_DEBUG=1 $P --reps 3
SCM_CSC_RED=1 SCM_CSC_DEBUG=1 $P --reps 3
SCM_CSC_RED=1 $P --k 8 --reps 3
SCM_CSC_RED=0 $P --k 8 --reps 3
