#!/bin/bash
# Build a variant of the library with extra -D flags for ONE source file, into variants/<name>.so:
#   bash tools/build_variant.sh <name> <source.cu> [-DFLAG ...]      (run `python -m camera_linearity_b200.build` first)
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p variants /tmp/variant_$name
obj=/tmp/variant_$name/$(basename ${src%.cu}).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Icamera_linearity_b200/csrc "$@" -c $src -o $obj
objs=""
for o in camera_linearity_b200/build/*.o; do
  if [ "$(basename $o)" = "$(basename $obj)" ]; then objs="$objs $obj"; else objs="$objs $o"; fi
done
nvcc -shared -o variants/$name.so $objs -gencode arch=compute_100a,code=sm_100a
echo variants/$name.so
