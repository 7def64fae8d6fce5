#!/bin/bash
# Tuning builds: tools/build_variant.sh NAME [-DFLAG ...] -> densepoints_b200/_variants/NAME.so
# (load with DENSEPOINTS_CUDA_LIB=...; never the shipped library)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
name=$1; shift
mkdir -p "$ROOT/densepoints_b200/_variants"
cd "$ROOT/densepoints_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -Xptxas -v -o "$ROOT/densepoints_b200/_variants/$name.so" densepoints_cuda.cu > "/tmp/build_variant_$name.log" 2>&1 \
  || { grep -i "error" "/tmp/build_variant_$name.log" | head -5; echo "BUILD FAILED <- $name"; exit 1; }
cat "/tmp/build_variant_$name.log" \
  | grep -A2 "dp_refine_group_kernelI10DpGroupCfgILi4ELi13E\|dp_refine_lane_kernelILi7E" | grep "spill\|Used" | tr '\n' ' '
echo " <- $name"
