#!/bin/bash
# A/B of two builds of the library on the SAME box: tools/ab_gemm.sh <old.so>   (diagnostic)
L=image-captioning-with-different-decoders_b200/libicd_b200.so
cp $L /tmp/new.so
for rep in 1 2; do
  for v in old new; do
    if [ $v = old ]; then cp $1 $L; else cp /tmp/new.so $L; fi
    echo "== $v (rep $rep)"; python tools/gemm_bench.py --dh --graph 2>&1 | tail -4
  done
done
cp /tmp/new.so $L
