#!/bin/bash
# Experiment builds of libflan_b200.so (timing experiments; select with FLAN_B200_LIB=<path>).
#   tools/abl_build.sh NAME -DPV_X [-DPV_Y ...]   ->  flan_b200/lib/abl/NAME/libflan_b200.so
# pv_modify / pv_io objects are taken from the regular build (flan_b200/lib/obj).
set -e
name=$1; shift
mkdir -p flan_b200/lib/abl/$name
nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 \
  -Xcompiler -fPIC,-ffp-contract=off "$@" -shared -o flan_b200/lib/abl/$name/libflan_b200.so \
  flan_b200/csrc/pv_kernels.cu flan_b200/csrc/pv_capi.cu flan_b200/lib/obj/pv_modify.cu.o flan_b200/lib/obj/pv_io.cu.o
