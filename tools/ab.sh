#!/bin/bash
# Same-box A/B of two library builds on the cfg2 merge (boxes differ by ~3 %, so deltas smaller than that are
# only visible when both builds alternate on one GPU):
#   build the baseline, cp camera_linearity_b200/libcamlin_b200.so camera_linearity_b200/libcamlin_old.so,
#   build the candidate, then   gpurun -- 'bash tools/ab.sh'
cd "$(dirname "$0")/.."
cp camera_linearity_b200/libcamlin_b200.so /tmp/new.so
for rep in 1 2; do
  for v in old new; do
    if [ $v = old ]; then cp camera_linearity_b200/libcamlin_old.so camera_linearity_b200/libcamlin_b200.so; else cp /tmp/new.so camera_linearity_b200/libcamlin_b200.so; fi
    echo -n "$v: "; python tools/run_merge.py 0.05 8 1 1 | tail -1
  done
done
cp /tmp/new.so camera_linearity_b200/libcamlin_b200.so
