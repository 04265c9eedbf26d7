#!/bin/bash
# A/B of two builds of the library on the SAME box: tools/ab_gemm.sh <old.so> [gemm_bench flags]   (diagnostic)
L=image-captioning-with-different-decoders_b200/libicd_b200.so
OLD=$1; shift
cp $L /tmp/new.so
for rep in 1 2; do
  for v in old new; do
    if [ $v = old ]; then cp $OLD $L; else cp /tmp/new.so $L; fi
    echo "== $v (rep $rep)"; python tools/gemm_bench.py "$@" 2>&1 | grep "M="
  done
done
cp /tmp/new.so $L
