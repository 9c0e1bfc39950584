#!/bin/bash
# Ablation builds of libflan_b200.so (timing experiments only: results are wrong by construction; select with FLAN_B200_LIB=<path>).
#   tools/abl_build.sh NAME -DPV_ABL_X [-DPV_ABL_Y ...]   ->  flan_b200/lib/abl/NAME/libflan_b200.so
# Only pv_kernels.cu is recompiled; every other object is taken from the regular build (flan_b200/lib/obj).
set -e
name=$1; shift
mkdir -p flan_b200/lib/abl/$name
nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 \
  -Xcompiler -fPIC,-ffp-contract=off "$@" -c -o flan_b200/lib/abl/$name/pv_kernels.o flan_b200/csrc/pv_kernels.cu
objs=$(ls flan_b200/lib/obj/*.o | grep -v pv_kernels.cu.o)
nvcc -shared -o flan_b200/lib/abl/$name/libflan_b200.so flan_b200/lib/abl/$name/pv_kernels.o $objs
echo flan_b200/lib/abl/$name/libflan_b200.so
