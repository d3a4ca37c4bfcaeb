#!/bin/bash
# Dev tool: A/B build of one translation unit with extra -D flags -> nerf_lidar_b200/libnlb200_<tag>.so
# usage: tools/build_variant.sh <tag> <file.cu> [-DFOO ...]   (load it with NLB_LIB=...)
set -e
cd "$(dirname "$0")/../nerf_lidar_b200"
tag=$1; src=$2; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$src -o build/${src%.cu}_$tag.o
objs=""
for o in build/*.o; do
  b=$(basename $o .o)
  case $b in
    ${src%.cu}) ;;
    *_*) [ "$b" = "${src%.cu}_$tag" ] && objs="$objs $o" || { [ -f csrc/$b.cu ] && objs="$objs $o"; } ;;
    *) objs="$objs $o" ;;
  esac
done
nvcc -shared -o libnlb200_$tag.so $objs -gencode arch=compute_100a,code=sm_100a -lcuda
echo built libnlb200_$tag.so from:$objs
