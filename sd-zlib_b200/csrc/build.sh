#!/bin/sh
# Builds libsdzcuda.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC,-O2,-Wall -shared -cudart static \
      -o libsdzcuda.so sdzcuda.cu -lpthread "$@"
