#!/bin/bash
# Same-box comparison of several library builds kept under variants/*.so:  bash tools/ab_variants.sh <script> [args]
cd "$(dirname "$0")/.."
script=$1; shift
cp camera_linearity_b200/libcamlin_b200.so /tmp/keep.so
for rep in 1 2; do
  for v in variants/*.so; do
    cp $v camera_linearity_b200/libcamlin_b200.so
    echo -n "$(basename $v): "; python $script "$@" | tail -1
  done
done
cp /tmp/keep.so camera_linearity_b200/libcamlin_b200.so
