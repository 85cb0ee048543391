#!/bin/bash
# build_variant.sh NAME [-D switches]: an experimental build of the CUDA library as build/variants/NAME.so
# (run with HIMUT_B200_LIB=build/variants/NAME.so; build/ travels to the GPU box, it is not in the history)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC "$@" \
  -Xptxas -v -o build/variants/$name.so himut_b200/csrc/himut_b200.cu 2> build/variants/$name.ptxas.log
grep -A1 "k_call_scanILb0" build/variants/$name.ptxas.log | grep -o "Used [0-9]* registers.*" | head -1
grep -A2 "k_call_scanILb0" build/variants/$name.ptxas.log | grep -o "[0-9]* bytes spill stores" | head -1
