#!/bin/bash
# diagnostic: run one command under several builds of the library on the same box: tools/ab_libs.sh "<command>" lib1 lib2 ...
L=image-captioning-with-different-decoders_b200/libicd_b200.so
CMD=$1; shift
cp $L /tmp/cur.so
for lib in /tmp/cur.so "$@"; do
  [ $lib != /tmp/cur.so ] && cp $lib $L
  echo "== $lib"; bash -c "$CMD"
done
cp /tmp/cur.so $L
