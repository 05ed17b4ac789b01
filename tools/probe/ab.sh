#!/bin/bash
# A/B of library builds on one box: tools/probe/ab.sh libA.so libB.so ...   (two rounds, alternating)
for round in $(seq 1 ${ROUNDS:-2}); do
  for lib in "$@"; do
    printf "%-32s " "$(basename $lib)"
    python tools/probe/bench_with_lib.py $lib --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python tools/probe/bench_with_lib.py --show
  done
done
