#!/bin/bash
# One gpurun call for compile-time variants of libsdzcuda.so (tools/build_variants.sh): the GPU parity suite against every
# variant named on the command line (SDZ_LIB), then the side-by-side timing of tools/bench_variants.py on the same box.
# usage: tools/gpu_variants_check.sh TAG test_variant[,test_variant...] bench_variant ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; TESTED=$2; shift 2
for v in ${TESTED//,/ }; do
    SDZ_LIB=$PWD/sd-zlib_b200/csrc/variants/$v.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_$v.log 2>&1
    echo "pytest $v rc=$?"; tail -4 gpurun_out/${TAG}_pytest_$v.log
done
timeout 900 python tools/bench_variants.py "$@" 2>&1 | tee gpurun_out/${TAG}_sweep.log
