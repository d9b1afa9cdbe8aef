#!/bin/sh
# Builds experiment variants of libsdzcuda.so (compile-time knobs of inflate_kernel.cuh) into
# sd-zlib_b200/csrc/variants/<name>.so; tools/bench_variants.py times them side by side on one box.
# usage: tools/build_variants.sh name1:"-DFLAG=1 -DOTHER=2" name2:"..." ...
set -e
cd "$(dirname "$0")/../sd-zlib_b200/csrc"
mkdir -p variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    [ "$flags" = "$spec" ] && flags=""
    ( $NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2,-Wall -shared -cudart static \
        -o variants/$name.so sdzcuda.cu -lpthread $flags && echo "built $name [$flags]" ) &
done
wait
