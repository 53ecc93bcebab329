#!/bin/bash
# A/B of two builds of the library on one box: ab/A.so (a saved build) against the in-tree one, alternating.
for rep in 1 2; do
  for v in A B; do
    if [ $v = A ]; then export SCGRHC_LIB=$PWD/ab/A.so; else unset SCGRHC_LIB; fi
    echo "== $v"; timeout 300 python tools/measure_planar.py 2>&1 | head -${1:-2}
  done
done
