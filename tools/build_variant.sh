#!/bin/bash
# tools/build_variant.sh NAME "<extra nvcc flags>" [source.cu ...]: an A/B build of the library with
# the named sources (default filtfilt.cu) recompiled with extra flags -> spikesort_b200/build/lib_vNAME.so
set -e
NAME=$1; FLAGS=$2; shift 2 || true
SRCS=${@:-filtfilt.cu}
B=spikesort_b200/build
OBJS=""
for o in $B/*.o; do
  keep=1
  for s in $SRCS; do [ "$(basename $o)" = "${s%.cu}.o" ] && keep=0; done
  [ $keep = 1 ] && [[ "$o" != *"_v"* ]] && OBJS="$OBJS $o"
done
for s in $SRCS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ispikesort_b200/csrc $FLAGS -c spikesort_b200/csrc/$s -o $B/${s%.cu}_v$NAME.o
  OBJS="$OBJS $B/${s%.cu}_v$NAME.o"
done
nvcc -shared -o $B/lib_v$NAME.so $OBJS
echo built $B/lib_v$NAME.so
