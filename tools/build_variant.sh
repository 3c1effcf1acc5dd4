#!/bin/bash
# Build a kernel variant of the tensor-core translation unit into variants/libb200node_<name>.so
# (git-ignored; selected at run time with B200NODE_LIB=...).  usage: tools/build_variant.sh name [-DFOO ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../diffeqflux.jl_b200/csrc"
mkdir -p ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c tc_kernels.cu -o ../../variants/tc_$name.o
nvcc -shared -o ../../variants/libb200node_$name.so ../b200node.o ../../variants/tc_$name.o
echo built variants/libb200node_$name.so
