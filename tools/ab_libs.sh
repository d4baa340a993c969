#!/bin/bash
# same-box A/B of two builds of the library on shipped shapes: tools/ab_libs.sh "cfg2 cfg1" iters old.so new.so
# (iters >= 300: the sustained, power-capped regime the bench measures; 3: a burst)
shapes=$1; iters=$2; shift 2
for w in $shapes; do
  for rep in 1 2; do
    for lib in "$@"; do
      printf "%s %-22s " "$w" "$(basename $lib)"
      KWS_B200_LIB=$PWD/$lib timeout 200 python tools/prof_fused.py $w --pairs-k 592 --iters $iters 2>&1 | tail -1
    done
  done
done
