#!/bin/bash
# Ablation builds of libflan_b200.so (timing experiments only; results are wrong by construction).
#   tools/abl_build.sh NAME -DPV_ABL_X [-DPV_ABL_Y ...]   ->  flan_b200/lib/abl/NAME/libflan_b200.so
set -e
name=$1; shift
mkdir -p flan_b200/lib/abl/$name
nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 \
  -Xcompiler -fPIC,-ffp-contract=off "$@" -shared -o flan_b200/lib/abl/$name/libflan_b200.so \
  flan_b200/csrc/pv_kernels.cu flan_b200/csrc/pv_capi.cu
