set -e
SRC="tools/microbench/scan_bench.cu camera_linearity_b200/csrc/hdr_merge.cu camera_linearity_b200/csrc/hdr_merge_staged.cu camera_linearity_b200/csrc/linearize.cu"
FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DWITH_LIB -Iinclude -Icamera_linearity_b200/csrc"
nvcc $FL -o /tmp/sb0 $SRC
/tmp/sb0 | grep -E "library"
