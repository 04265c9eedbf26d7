#!/bin/bash
# diagnostic: the vocabulary-layer contraction under tile / pairing / cluster variants, then with the epilogue's stores / staging
# compiled out (libraries built with -DICD_GEMM_EPI_DEBUG=1|2, passed as arguments)
L=image-captioning-with-different-decoders_b200/libicd_b200.so
python tools/gemm_bench.py --fc --graph 2>&1 | grep "M="
cp $L /tmp/new.so
for dbg in "$@"; do
  cp $dbg $L; touch image-captioning-with-different-decoders_b200/build/stamp.txt
  echo "== $dbg"; ICD_SKIP_BUILD=1 python tools/gemm_bench.py --graph 2>&1 | grep "M=" | head -6
done
cp /tmp/new.so $L
