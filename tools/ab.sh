#!/bin/bash
# Same-box A/B of two library builds (boxes differ by ~3 %, so smaller deltas are only visible when both builds
# alternate on one GPU):
#   build the baseline, cp camera_linearity_b200/libcamlin_b200.so camera_linearity_b200/libcamlin_old.so,
#   build the candidate, then   gpurun -- 'bash tools/ab.sh [script] [args]'      (default script: tools/ab_merge.py)
cd "$(dirname "$0")/.."
script=${1:-tools/ab_merge.py}; shift
cp camera_linearity_b200/libcamlin_b200.so /tmp/new.so
for rep in 1 2 3; do
  for v in old new; do
    if [ $v = old ]; then cp camera_linearity_b200/libcamlin_old.so camera_linearity_b200/libcamlin_b200.so; else cp /tmp/new.so camera_linearity_b200/libcamlin_b200.so; fi
    echo -n "$v: "; python $script "$@" | tail -1
  done
done
cp /tmp/new.so camera_linearity_b200/libcamlin_b200.so
